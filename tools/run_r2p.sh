#!/bin/bash
# round 2, GPU call P: warp-granular persistent schedule as the default: full GPU suite, bench record (all legs), tail
# fraction A/B, ncu launch list + --set full capture of the pair kernel as shipped, per-config throughput
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=15 > $O/r2p_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2p_steps.log
timeout 900 python bench.py > $O/r2p_bench_g1.json 2> $O/r2p_bench_g1.err; echo "bench rc=$?" >> $O/r2p_steps.log
Q="--no-cpu --no-e2e --no-lammps --no-parity --steps 50"
for t in 0.05 0.2 0.35; do
  SPHBVF_PAIR_TAIL=$t timeout 600 python bench.py $Q > $O/r2p_bench_tail$t.json 2> /dev/null; echo "bench tail $t rc=$?" >> $O/r2p_steps.log
done
B="--no-cpu --no-e2e --no-lammps --no-parity --steps 20 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2p_launches.csv python bench.py $B > $O/r2p_ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $O/r2p_steps.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel --launch-skip 6 --launch-count 1 -f -o /tmp/r2p_pair python bench.py $B > $O/r2p_ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/r2p_steps.log
ncu -i /tmp/r2p_pair.ncu-rep --page raw --csv > $O/r2p_pair.raw.csv 2>/dev/null
ncu -i /tmp/r2p_pair.ncu-rep --page details > $O/r2p_pair.details.txt 2>/dev/null
ncu -i /tmp/r2p_pair.ncu-rep --page source --print-source sass --csv > $O/r2p_pair.src.csv 2>/dev/null
timeout 1500 python tools/config_bench.py --scales 1,8 --steps 200 --warm 40 --out $O/r2p_config_bench.json > $O/r2p_config_bench.txt 2>&1; echo "config rc=$?" >> $O/r2p_steps.log
cat $O/r2p_steps.log; grep -E "passed|failed" $O/r2p_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $O/r2p_pytest.log | head -20
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4g atom-steps/s"%d["value"], "%.3f ms/step"%d["ms_per_step"], "pair %.3f ms"%(d["kernels"]["pair"]["ms"]/d["steps"]), d.get("e2e") and "%.3g"%d["e2e"]["value"], d.get("lammps_dropin") and d["lammps_dropin"].get("value"))
    except Exception as e: print(f, "failed", e)
PY
cat $O/r2p_config_bench.txt
