#!/bin/bash
# round 2, GPU call Q: Z-order (Morton) tile order under the persistent schedule: A/B against row order at 8 M atoms,
# parity at 64^3 / 100^3 / 8 M, --set full capture
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_midsize_oracle.py tests/test_fullsize_properties.py tests/test_gpu_parity.py -q -p no:cacheprovider --maxfail=10 > $O/r2q_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2q_steps.log
Q="--no-cpu --no-e2e --no-lammps --no-parity --steps 50"
timeout 600 python bench.py $Q > $O/r2q_bench_morton.json 2> $O/r2q_bench_morton.err; echo "bench morton rc=$?" >> $O/r2q_steps.log
SPHBVF_TILE_ORDER=row timeout 600 python bench.py $Q > $O/r2q_bench_row.json 2> /dev/null; echo "bench row rc=$?" >> $O/r2q_steps.log
SPHBVF_PAIR_SCHED=grid timeout 600 python bench.py $Q > $O/r2q_bench_grid.json 2> /dev/null; echo "bench grid rc=$?" >> $O/r2q_steps.log
B="--no-cpu --no-e2e --no-lammps --no-parity --steps 20 --warmup 3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel --launch-skip 6 --launch-count 1 -f -o /tmp/r2q_pair python bench.py $B > $O/r2q_ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/r2q_steps.log
ncu -i /tmp/r2q_pair.ncu-rep --page raw --csv > $O/r2q_pair.raw.csv 2>/dev/null
ncu -i /tmp/r2q_pair.ncu-rep --page source --print-source sass --csv > $O/r2q_pair.src.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2q_launches.csv python bench.py $B > $O/r2q_ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $O/r2q_steps.log
cat $O/r2q_steps.log; grep -E "passed|failed" $O/r2q_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $O/r2q_pytest.log | head
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2q_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["kernels"]
        print(f, "%.4g atom-steps/s"%d["value"], "%.3f ms/step"%d["ms_per_step"], "pair %.3f ms"%(k["pair"]["ms"]/d["steps"]), "rebuild %.2f ms/10"%(k["neighbor_rebuild"]["ms"]/5), "fused %.3f"%(k["final_initial_pack_fused"]["ms"]/max(1,k["final_initial_pack_fused"]["launches"])))
    except Exception as e: print(f, "failed", e)
PY
