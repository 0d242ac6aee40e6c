#!/bin/bash
# round 2, GPU call S (2 GPUs): bench.py as the driver launches it at N=2, lammps_dropin leg included (16 M atoms through
# ONE lmp_cuda process on both GPUs after the ranks have gone)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
( time timeout 1500 $T bench.py --gpus 2 --steps 100 --warmup 10 > $O/r2s_bench_g2.json 2> $O/r2s_bench_g2.err ) 2> $O/r2s_time.txt; echo "bench g2 rc=$?" > $O/r2s_steps.log
( time timeout 900 $T bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/r2s_bench_ref_g2.json 2> $O/r2s_bench_ref_g2.err ) 2>> $O/r2s_time.txt; echo "bench ref g2 rc=$?" >> $O/r2s_steps.log
cat $O/r2s_steps.log $O/r2s_time.txt; cut -c1-1500 $O/r2s_bench_g2.json; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2s_bench_g2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["parity_check"]["ok"], d["lammps_dropin"], d["e2e"]["value"])
print(open("gpurun_out/r2s_bench_ref_g2.json").read()[:600])
PY
tail -5 $O/r2s_bench_g2.err
