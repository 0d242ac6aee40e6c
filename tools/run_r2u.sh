#!/bin/bash
# round 2, GPU call U: late B/C record loads (PAIR_BC_LATE) at 12 and 16 warps per SM against the shipped kernel
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
STEPS=50 BENCH_ARGS="--no-lammps --no-parity" bash tools/ab_bench.sh base late192 late256 late128x4 2>&1 | tee gpurun_out/r2u_ab.txt
