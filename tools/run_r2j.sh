#!/bin/bash
# round 2, GPU call J: per-config throughput (BASELINE configs 1-4) through lmp_cuda -sf cuda + ncu of the pair instantiations
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python tools/config_bench.py --scales 1,8 --steps 200 --warm 40 --out $O/r2j_config_bench.json > $O/r2j_config_bench.txt 2>&1; echo "config rc=$?" > $O/r2j_steps.log
timeout 600 python tools/config_bench.py --exe lmp_serial --scales 1 --steps 20 --warm 5 --out $O/r2j_config_ref.json > $O/r2j_config_ref.txt 2>&1; echo "ref rc=$?" >> $O/r2j_steps.log
# ncu: one pair launch of each scaled deck
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
  (cd /tmp/ncu_$d && timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel --launch-skip 8 --launch-count 1 -f -o $GRAFT_REPO_ROOT/$O/r2j_pair_$d $GRAFT_REPO_ROOT/sph-bvf_b200/lammps/_build/lmp_cuda -in in.lmp -log none -echo none -sf cuda > $GRAFT_REPO_ROOT/$O/r2j_ncu_$d.log 2>&1); echo "ncu $d rc=$?" >> $O/r2j_steps.log
done
cat $O/r2j_steps.log; cat $O/r2j_config_bench.txt; cat $O/r2j_config_ref.txt
