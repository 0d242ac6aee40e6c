#!/bin/bash
# round 2, GPU call G: persistent SM-local schedule of the gather kernel, A/B + parity + L1 hit rate
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu --no-e2e --no-lammps"
timeout 600 python bench.py --steps 50 $B > $O/r2g_bench_smid.json 2> $O/r2g_bench_smid.err; echo "bench smid rc=$?" > $O/r2g_steps.log
SPHBVF_PAIR_SCHED=grid timeout 600 python bench.py --steps 50 $B --no-parity > $O/r2g_bench_grid.json 2> $O/r2g_bench_grid.err; echo "bench grid rc=$?" >> $O/r2g_steps.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tile_vs_gather.py tests/test_gpu_edge_cases.py tests/test_midsize_oracle.py -q -p no:cacheprovider --maxfail=12 > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_steps.log
M="l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum"
timeout 600 ncu --metrics $M --clock-control none -k regex:pair_kernel --launch-skip 4 --launch-count 2 --csv --log-file $O/r2g_ncu_smid.csv python bench.py --steps 4 --warmup 3 $B --no-parity > /dev/null 2>&1; echo "ncu smid rc=$?" >> $O/r2g_steps.log
cat $O/r2g_steps.log; tail -3 $O/r2g_pytest.log
python - <<'PY'
import json
for n in ("smid","grid"):
    try:
        b=json.load(open("gpurun_out/r2g_bench_%s.json"%n)); print(n, "%.4g"%b["value"], "ms/step %.3f"%b["ms_per_step"], "pair %.3f ms"%b["roofline"]["pair_ms_per_step"], "parity", (b.get("parity_check") or {}).get("ok"))
    except Exception as e: print(n, "ERR", e)
PY
grep -E "hit_rate|time_duration|fp64|issue_active|wavefronts" $O/r2g_ncu_smid.csv | cut -d, -f5,13- | head -16
