#!/bin/bash
# round 2, GPU call R (2 GPUs): multi-GPU tests with the lazy atom style / persistent schedule, bench at N=2 default vs
# one CTA per chunk
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_lammps_dropin.py -k "bricks or two_gpus" -q -p no:cacheprovider --maxfail=10 > $O/r2r_pytest_2gpu.log 2>&1; echo "pytest rc=$?" > $O/r2r_steps.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
Q="--gpus 2 --no-cpu --no-lammps --steps 50"
timeout 900 $T bench.py $Q > $O/r2r_bench_g2_warp.json 2> $O/r2r_bench_g2_warp.err; echo "bench g2 warp rc=$?" >> $O/r2r_steps.log
SPHBVF_PAIR_SCHED=grid timeout 900 $T bench.py $Q --no-parity --no-e2e > $O/r2r_bench_g2_grid.json 2> $O/r2r_bench_g2_grid.err; echo "bench g2 grid rc=$?" >> $O/r2r_steps.log
cat $O/r2r_steps.log; tail -3 $O/r2r_pytest_2gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2r_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(f, "%.4g atom-steps/s"%d["value"], "%.3f ms/step"%d["ms_per_step"], "pair %.3f ms"%(k["pair"]["ms"]/d["steps"]), "halo %.2f"%k["pack_halo"]["ms"], "rebuild %.2f"%k["neighbor_rebuild"]["ms"], d.get("parity_check") and d["parity_check"]["ok"])
    except Exception as e: print(f, "failed", e)
PY
