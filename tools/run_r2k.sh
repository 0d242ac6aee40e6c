#!/bin/bash
# round 2, GPU call K: ncu of the pair instantiations of configs 2-4 (reports exported to text on the box)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=$GRAFT_REPO_ROOT/gpurun_out
python - <<'PY'
import sys, os
sys.path.insert(0, "tools")
import config_bench as cb
for name in ("natconv", "fsi", "cellpol"):
    fn, rv, _ = cb.DECKS[name]
    os.makedirs("/tmp/ncu_" + name, exist_ok=True)
    open("/tmp/ncu_%s/in.lmp" % name, "w").write(cb.edit(open(os.path.join(cb.DECKDIR, fn)).read(), rv, 8, 12, 3))
PY
for d in natconv fsi cellpol; do
  (cd /tmp/ncu_$d && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel --launch-skip 8 --launch-count 1 -f -o /tmp/pair_$d $GRAFT_REPO_ROOT/sph-bvf_b200/lammps/_build/lmp_cuda -in in.lmp -log none -echo none -sf cuda > $O/r2k_ncu_$d.log 2>&1); echo "ncu $d rc=$?" >> $O/r2k_steps.log
  ncu -i /tmp/pair_$d.ncu-rep --page details > $O/r2k_pair_$d.details.txt 2>&1
  ncu -i /tmp/pair_$d.ncu-rep --page raw --csv > $O/r2k_pair_$d.raw.csv 2>&1
  ncu -i /tmp/pair_$d.ncu-rep --page source --print-source sass,cuda --csv > $O/r2k_pair_$d.src.csv 2>&1
done
cat $O/r2k_steps.log; ls -la $O
