#!/bin/bash
# round 2, GPU call C: full GPU suite (gather default, tile form selectable) + bench
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > $O/r2c_smoke.log 2>&1; echo "smoke rc=$?" > $O/r2c_steps.log
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=25 -s > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_steps.log
timeout 900 python bench.py > $O/r2c_bench.json 2> $O/r2c_bench.err; echo "bench rc=$?" >> $O/r2c_steps.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2c_bench_ref.json 2> $O/r2c_bench_ref.err; echo "bench ref rc=$?" >> $O/r2c_steps.log
SPHBVF_PAIR=tile timeout 600 python bench.py --steps 30 --no-cpu --no-e2e --no-lammps > $O/r2c_bench_tile.json 2> $O/r2c_bench_tile.err; echo "bench tile rc=$?" >> $O/r2c_steps.log
cat $O/r2c_steps.log; grep -E "passed|failed" $O/r2c_pytest.log | tail -3; grep -E "^FAILED|^ERROR" $O/r2c_pytest.log | head -20
