#!/bin/bash
# round 2, GPU call N: atom_style ssa_tsdpd/atomic/cuda under -sf cuda (every deck test now picks it), late fetch from the
# parked contexts, small decks with the tile form, 64 M atoms through lmp_cuda on ONE GPU (lean host mirrors)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
free -g | head -2 > $O/r2n_host.txt; nproc >> $O/r2n_host.txt
timeout 1200 python -m pytest tests/test_lammps_dropin.py tests/test_shipped_decks.py tests/test_atom_style_cuda.py -q -p no:cacheprovider --maxfail=10 > $O/r2n_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2n_steps.log
SPHBVF_PAIR=tile timeout 300 python tools/small_deck_bench.py > $O/r2n_small_tile.txt 2>&1; echo "small tile rc=$?" >> $O/r2n_steps.log
MEM=$(free -g | awk '/^Mem:/{print $7}')
if [ "$MEM" -ge 90 ]; then
  timeout 1200 python tools/lmp_cuda_bench.py 400 1 20 > $O/r2n_lmp_cuda_64M_1gpu.txt 2>&1; echo "lmp 64M rc=$?" >> $O/r2n_steps.log
else
  echo "lmp 64M skipped: $MEM GB available" >> $O/r2n_steps.log
fi
cat $O/r2n_steps.log $O/r2n_host.txt; tail -4 $O/r2n_pytest.log; grep -E "^FAILED|^ERROR" $O/r2n_pytest.log | head; cat $O/r2n_small_tile.txt; tail -5 $O/r2n_lmp_cuda_64M_1gpu.txt
