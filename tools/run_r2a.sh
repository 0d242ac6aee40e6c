#!/bin/bash
# round 2, GPU call A: first run of the tile-staged pair kernel
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > $O/r2a_gpu.txt 2>&1
echo "== smoke" > $O/r2a_steps.log
timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > $O/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2a_steps.log
echo "== tile vs gather + parity (fast fixtures)" >> $O/r2a_steps.log
timeout 900 python -m pytest tests/test_tile_vs_gather.py tests/test_gpu_parity.py -q -p no:cacheprovider --maxfail=12 > $O/r2a_pytest_fast.log 2>&1; echo "pytest fast rc=$?" >> $O/r2a_steps.log
echo "== bench tile" >> $O/r2a_steps.log
timeout 900 python bench.py --steps 50 --warmup 10 > $O/r2a_bench_tile.json 2> $O/r2a_bench_tile.err; echo "bench tile rc=$?" >> $O/r2a_steps.log
echo "== bench gather" >> $O/r2a_steps.log
SPHBVF_PAIR=gather timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu --no-e2e --no-lammps --no-parity > $O/r2a_bench_gather.json 2> $O/r2a_bench_gather.err; echo "bench gather rc=$?" >> $O/r2a_steps.log
echo "== rest of pytest" >> $O/r2a_steps.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=20 --deselect tests/test_tile_vs_gather.py --deselect tests/test_gpu_parity.py -s > $O/r2a_pytest_rest.log 2>&1; echo "pytest rest rc=$?" >> $O/r2a_steps.log
tail -5 $O/r2a_pytest_fast.log; tail -5 $O/r2a_pytest_rest.log; cat $O/r2a_steps.log
