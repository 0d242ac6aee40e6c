#!/bin/bash
# round 2, GPU call F: full suite after the output-path work, ncu captures of the shipped kernels, host info
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
(free -g; nproc; lscpu | grep "Model name") > $O/r2f_host.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --maxfail=25 -s > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2f_steps.log
B="--no-cpu --no-e2e --no-lammps --no-parity"
timeout 600 python bench.py --steps 50 $B > $O/r2f_bench.json 2> $O/r2f_bench.err; echo "bench rc=$?" >> $O/r2f_steps.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel --launch-skip 4 --launch-count 1 -f -o $O/r2f_pair_gather python bench.py --steps 3 --warmup 3 $B > $O/r2f_ncu1.log 2>&1; echo "ncu pair rc=$?" >> $O/r2f_steps.log
timeout 900 ncu --set full --clock-control none -k regex:"integrate_kernel|build_list_tile_kernel|gather_state_kernel" --launch-skip 3 --launch-count 12 -f -o $O/r2f_others python bench.py --steps 12 --warmup 3 $B > $O/r2f_ncu2.log 2>&1; echo "ncu others rc=$?" >> $O/r2f_steps.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches.csv python bench.py --steps 30 --warmup 3 $B > $O/r2f_ncu3.log 2>&1; echo "ncu launches rc=$?" >> $O/r2f_steps.log
timeout 1500 python tools/lmp_cuda_bench.py 252 1 20 > $O/r2f_lmp252.log 2>&1; echo "lmp 252 rc=$?" >> $O/r2f_steps.log
cat $O/r2f_steps.log; grep -E "passed|failed" $O/r2f_pytest.log | tail -2; grep -E "^FAILED|^ERROR" $O/r2f_pytest.log | head; tail -2 $O/r2f_lmp252.log; cat $O/r2f_host.txt
