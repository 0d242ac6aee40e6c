#!/bin/bash
# round 2, GPU call B: tile kernel variants A/B + ncu captures
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
B="--steps 30 --warmup 10 --no-cpu --no-e2e --no-lammps --no-parity"
SPHBVF_VERBOSE=1 timeout 600 python bench.py $B > $O/r2b_tile.json 2> $O/r2b_tile.err; echo "tile rc=$?" > $O/r2b_steps.log
for v in t192 t128; do
  SPHBVF_LIB=$PWD/sph-bvf_b200/variants/libsphbvf_$v.so timeout 600 python bench.py $B > $O/r2b_$v.json 2> $O/r2b_$v.err; echo "$v rc=$?" >> $O/r2b_steps.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_tile_kernel --launch-skip 4 --launch-count 1 -f -o $O/r2b_pair_tile python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-lammps --no-parity > $O/r2b_ncu1.log 2>&1; echo "ncu pair rc=$?" >> $O/r2b_steps.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_list_tile_kernel --launch-skip 1 --launch-count 1 -f -o $O/r2b_build_tile python bench.py --steps 12 --warmup 3 --no-cpu --no-e2e --no-lammps --no-parity > $O/r2b_ncu2.log 2>&1; echo "ncu build rc=$?" >> $O/r2b_steps.log
cat $O/r2b_steps.log; grep sphbvf $O/r2b_tile.err | head -3
python - <<'PY'
import json
for n in ("tile","t192","t128"):
    try:
        b=json.load(open("gpurun_out/r2b_%s.json"%n)); print(n, "%.4g"%b["value"], "pair %.3f ms"%b["roofline"]["pair_ms_per_step"], "rebuild %.3f ms/step"%(b["kernels"]["neighbor_rebuild"]["ms"]/b["steps"]))
    except Exception as e: print(n, "ERR", e)
PY
