#!/bin/bash
# round 2, GPU call O: full GPU suite with the split-lane kernel / lazy atom style / quantified Press, small decks A/B,
# warp-granular persistent schedule A/B at 8 M atoms, 64 M atoms through lmp_cuda on ONE GPU
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider --maxfail=15 > $O/r2o_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2o_steps.log
timeout 300 python tools/small_deck_bench.py > $O/r2o_small_split.txt 2>&1; echo "small split rc=$?" >> $O/r2o_steps.log
SPHBVF_PAIR_LANES=1 timeout 300 python tools/small_deck_bench.py > $O/r2o_small_lanes1.txt 2>&1; echo "small lanes1 rc=$?" >> $O/r2o_steps.log
Q="--no-cpu --no-e2e --no-lammps --no-parity --steps 50"
timeout 600 python bench.py $Q > $O/r2o_bench_grid.json 2> $O/r2o_bench_grid.err; echo "bench grid rc=$?" >> $O/r2o_steps.log
SPHBVF_PAIR_SCHED=warp timeout 600 python bench.py $Q > $O/r2o_bench_warp.json 2> $O/r2o_bench_warp.err; echo "bench warp rc=$?" >> $O/r2o_steps.log
SPHBVF_PAIR_SCHED=smid timeout 600 python bench.py $Q > $O/r2o_bench_smid.json 2> $O/r2o_bench_smid.err; echo "bench smid rc=$?" >> $O/r2o_steps.log
timeout 900 python tools/lmp_cuda_bench.py 400 1 20 > $O/r2o_lmp_cuda_64M_1gpu.txt 2>&1; echo "lmp 64M rc=$?" >> $O/r2o_steps.log
cat $O/r2o_steps.log; grep -E "passed|failed" $O/r2o_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $O/r2o_pytest.log | head -20; grep "thermo Press" $O/r2o_pytest.log
cat $O/r2o_small_split.txt $O/r2o_small_lanes1.txt
python - <<'PY'
import json
for n in ("grid","warp","smid"):
    try:
        d=json.loads(open("gpurun_out/r2o_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "%.4g atom-steps/s"%d["value"], "%.3f ms/step"%d["ms_per_step"], "pair %.3f ms"%(d["kernels"]["pair"]["ms"]/50))
    except Exception as e: print(n, "failed", e)
PY
tail -4 $O/r2o_lmp_cuda_64M_1gpu.txt | cut -c1-600
