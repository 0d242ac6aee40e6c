#!/bin/bash
# round 2, GPU call T (4 GPUs): 4-brick parity tests and bench.py as the driver launches it at N=4 (2 x 2 x 1 bricks,
# lammps_dropin leg: 32 M atoms through ONE lmp_cuda process)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -q -p no:cacheprovider > $O/r2t_pytest_4gpu.log 2>&1; echo "pytest rc=$?" > $O/r2t_steps.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521"
( time timeout 1500 $T bench.py --gpus 4 --steps 100 --warmup 10 > $O/r2t_bench_g4.json 2> $O/r2t_bench_g4.err ) 2> $O/r2t_time.txt; echo "bench g4 rc=$?" >> $O/r2t_steps.log
cat $O/r2t_steps.log $O/r2t_time.txt; tail -2 $O/r2t_pytest_4gpu.log; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2t_bench_g4.json").read().strip().splitlines()[-1])
k=d["kernels"]
print(d["value"], d["ms_per_step"], d["parity_check"]["ok"], d["parity_check"]["max_err"], d["lammps_dropin"], d["e2e"]["value"], {a:round(b["ms"],1) for a,b in k.items()})
PY
tail -3 $O/r2t_bench_g4.err
