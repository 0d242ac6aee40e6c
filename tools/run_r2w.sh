#!/bin/bash
# round 2, GPU call W: --set full captures of the shipped list builder (8 M atoms) and of the split-lane pair kernel (42 k atoms)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
B="--no-cpu --no-e2e --no-lammps --no-parity --steps 12 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_list_tile32 --launch-skip 1 --launch-count 1 -f -o /tmp/r2w_list python bench.py $B > $O/r2w_ncu_list.log 2>&1; echo "ncu list rc=$?" > $O/r2w_steps.log
ncu -i /tmp/r2w_list.ncu-rep --page raw --csv > $O/r2w_list.raw.csv 2>/dev/null
ncu -i /tmp/r2w_list.ncu-rep --page source --print-source sass --csv > $O/r2w_list.src.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_split --launch-skip 300 --launch-count 1 -f -o /tmp/r2w_split python tools/small_deck_bench.py > $O/r2w_ncu_split.log 2>&1; echo "ncu split rc=$?" >> $O/r2w_steps.log
ncu -i /tmp/r2w_split.ncu-rep --page raw --csv > $O/r2w_split.raw.csv 2>/dev/null
ncu -i /tmp/r2w_split.ncu-rep --page source --print-source sass --csv > $O/r2w_split.src.csv 2>/dev/null
cat $O/r2w_steps.log; ls -la $O/r2w_*
