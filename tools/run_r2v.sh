#!/bin/bash
# round 2, GPU call V: final state of the round: smoke, full GPU suite, bench (both arms) as the driver runs them
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > $O/r2v_smoke.log 2>&1; echo "smoke rc=$?" > $O/r2v_steps.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=15 > $O/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2v_steps.log
( time timeout 900 python bench.py > $O/r2v_bench_g1.json 2> $O/r2v_bench_g1.err ) 2> $O/r2v_time.txt; echo "bench rc=$?" >> $O/r2v_steps.log
( time timeout 900 python bench.py --impl reference > $O/r2v_bench_ref_g1.json 2> $O/r2v_bench_ref_g1.err ) 2>> $O/r2v_time.txt; echo "bench ref rc=$?" >> $O/r2v_steps.log
cat $O/r2v_steps.log $O/r2v_smoke.log $O/r2v_time.txt; grep -E "passed|failed" $O/r2v_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $O/r2v_pytest.log | head
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2v_bench_g1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"], d["lammps_dropin"]["value"], d["parity_check"]["ok"], d["roofline"]["frac"], d["roofline"]["fp64_frac"], d["clocks"])
r=json.loads(open("gpurun_out/r2v_bench_ref_g1.json").read().strip().splitlines()[-1])
print(r["value"], r["steps"], r["cpu_baseline"]["cores"], r["ms_per_step"])
PY
