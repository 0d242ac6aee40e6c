#!/bin/bash
# round 2, GPU call I: float-classifying tile list builder, A/B against the all-FP64 sweep, parity, ncu
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu --no-e2e --no-lammps"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tile_vs_gather.py tests/test_gpu_edge_cases.py tests/test_midsize_oracle.py -q -p no:cacheprovider --maxfail=12 > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2i_steps.log
timeout 600 python bench.py --steps 50 $B > $O/r2i_bench_t32.json 2> $O/r2i_bench_t32.err; echo "bench t32 rc=$?" >> $O/r2i_steps.log
SPHBVF_LIST_BUILD=tile64 timeout 600 python bench.py --steps 50 $B --no-parity > $O/r2i_bench_t64.json 2> $O/r2i_bench_t64.err; echo "bench t64 rc=$?" >> $O/r2i_steps.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_list_tile32 --launch-skip 1 --launch-count 1 -f -o $O/r2i_build32 python bench.py --steps 12 --warmup 3 $B --no-parity > $O/r2i_ncu1.log 2>&1; echo "ncu rc=$?" >> $O/r2i_steps.log
cat $O/r2i_steps.log; tail -3 $O/r2i_pytest.log; grep -E "^FAILED|^ERROR" $O/r2i_pytest.log | head
python - <<'PY'
import json
for n in ("t32","t64"):
    try:
        b=json.load(open("gpurun_out/r2i_bench_%s.json"%n)); k=b["kernels"]["neighbor_rebuild"]; print(n, "%.4g"%b["value"], "ms/step %.3f"%b["ms_per_step"], "pair %.3f ms"%b["roofline"]["pair_ms_per_step"], "rebuild ms total %.2f launches %d"%(k["ms"],k["launches"]), "parity", (b.get("parity_check") or {}).get("ok"))
    except Exception as e: print(n, "ERR", e)
PY
