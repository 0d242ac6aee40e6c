#!/bin/bash
# round 2, GPU call X (2 GPUs): early halo (started by the fused integrator, faces first) against the overlapped halo
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_lammps_dropin.py -k "bricks or two_gpus" -q -p no:cacheprovider --maxfail=10 > $O/r2x_pytest_2gpu.log 2>&1; echo "pytest rc=$?" > $O/r2x_steps.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523"
Q="--gpus 2 --no-cpu --no-lammps --no-e2e --steps 100"
timeout 900 $T bench.py $Q > $O/r2x_bench_g2_early.json 2> $O/r2x_bench_g2_early.err; echo "bench g2 early rc=$?" >> $O/r2x_steps.log
SPHBVF_HALO=overlap timeout 900 $T bench.py $Q --no-parity > $O/r2x_bench_g2_overlap.json 2> $O/r2x_bench_g2_overlap.err; echo "bench g2 overlap rc=$?" >> $O/r2x_steps.log
cat $O/r2x_steps.log; tail -3 $O/r2x_pytest_2gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2x_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(f, "%.4g atom-steps/s"%d["value"], "%.3f ms/step"%d["ms_per_step"], "pair %.3f ms"%(k["pair"]["ms"]/d["steps"]), "fused %.3f"%(k["final_initial_pack_fused"]["ms"]/d["steps"]), "halo %.2f"%k["pack_halo"]["ms"], "rebuild %.2f"%k["neighbor_rebuild"]["ms"], d.get("parity_check") and (d["parity_check"]["ok"], d["parity_check"]["max_err"]), d["gpu_launches"])
    except Exception as e: print(f, "failed", e)
PY
tail -3 $O/r2x_bench_g2_early.err
