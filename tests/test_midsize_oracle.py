"""CUDA path against the plain-C oracle at sizes between the golden fixtures (<= 2 744 atoms in 3D) and the
benchmark (8 M atoms): the synthetic 3D cavity lattice of BASELINE.json configs[4] / SURVEY.md 8(d) at
64^3 = 262 144 and 100^3 = 1 000 000 atoms.  At these sizes the cell grid holds thousands of full interior
4x4x4-cell tiles (the 14^3 fixture has less than one), tile-major cell numbers and packed list entries pass
2^16, and every tile shape of the list builder and of the tile-staged pair kernel occurs -- what a
tile-boundary or index-width bug needs in order to show.

The oracle (pinned to the reference by tests/test_oracle_vs_reference.py) needs ~10 us per atom-step on one
core, so the horizon is short: 64^3 runs the deck as benchmarked (11 steps, the step-10 rebuild included);
100^3 uses `neigh_modify delay 2` so that a rebuild falls inside 4 steps (the lid moves skin/2 per step).
Bar: pair sets bit-exact, every consumed field within 1e-10 of its max-norm, same number of rebuilds.
"""
import os
import sys

import numpy as np
import pytest

from common import TOL
from conftest import ROOT, load_package

pytestmark = pytest.mark.gpu
sys.path.insert(0, ROOT)

FIELDS = ("x", "v", "vest", "rho", "rhoI", "f", "drho", "phi", "number_density", "nw", "ddv", "rhoAux2")


def pair_keys(p):
    """unordered tag pairs -> sorted int64 keys (min << 32 | max)"""
    p = np.asarray(p, dtype=np.int64).reshape(-1, 2)
    k = (np.minimum(p[:, 0], p[:, 1]) << 32) | np.maximum(p[:, 0], p[:, 1])
    k.sort()
    return k


@pytest.mark.parametrize("n,delay,nsteps", [(64, 10, 11), (100, 2, 4)])
def test_lattice_matches_oracle(n, delay, nsteps):
    import bench
    from oracle_api import Oracle
    if n > int(os.environ.get("SPHBVF_MIDSIZE_MAX_N", "100")):
        pytest.skip("SPHBVF_MIDSIZE_MAX_N caps the lattice edge")
    pkg = load_package()
    meta = bench.cavity_meta(n)
    meta["delay"] = delay
    a = bench.cavity_atoms(meta, meta["boxlo"], meta["boxhi"])
    eng, orc = pkg.Engine(meta), Oracle(meta)
    for e in (eng, orc):
        e.set_atoms(a["tag"], a["type"], a["mask"], a["solid"], a["fixed"], a["x"], a["v"], a["rho"], a["e"])
        e.set_run_length(10 ** 6)
        e.setup()
    # step 0: the list of the setup build and the first pair pass
    assert np.array_equal(pair_keys(eng.pairs()), pair_keys(orc.pairs())), "pair set after setup"
    for e in (eng, orc):
        e.run(nsteps)
    assert eng.nbuilds == orc.nbuilds >= 1, (eng.nbuilds, orc.nbuilds)
    keep = ~((a["solid"] == 1) & (a["fixed"] == 1))   # force on fixed walls: never consumed, orientation dependent
    worst = {}
    for f in FIELDS:
        ref, got = orc.get(f), eng.get(f)
        if f == "f":
            ref, got = ref[keep], got[keep]
        assert np.array_equal(np.isfinite(got), np.isfinite(ref)), f
        err = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))
        worst[f] = err
        assert err <= TOL, "n=%d field %s: err %.3e" % (n, f, err)
    assert np.array_equal(pair_keys(eng.pairs()), pair_keys(orc.pairs())), "pair set after the rebuild"
    print("midsize n=%d: %d atoms, %d steps, %d rebuilds, worst field errors %s" % (
        n, n ** 3, nsteps, eng.nbuilds, {k: "%.1e" % v for k, v in worst.items()}))
    eng.close()
    orc.close()


@pytest.mark.parametrize("sched", ["smid", "warp"])
def test_persistent_schedules_are_bitwise_the_default(sched, monkeypatch):
    """SPHBVF_PAIR_SCHED=smid / warp (persistent CTAs pulling 192-atom / 32-atom chunks from SM-local queues) only
    change WHICH thread evaluates an atom, never the order of its neighbour sum: every pair-pass output must be
    bit-identical to the one-CTA-per-chunk launch.  64^3 = 1366 chunks, past the 4 x SMs threshold below which the
    persistent launch is not used."""
    import bench
    pkg = load_package()
    meta = bench.cavity_meta(64)
    a = bench.cavity_atoms(meta, meta["boxlo"], meta["boxhi"])
    out = {}
    for mode in ("grid", sched):
        monkeypatch.setenv("SPHBVF_PAIR_SCHED", mode)
        eng = pkg.Engine(meta)
        eng.set_atoms(a["tag"], a["type"], a["mask"], a["solid"], a["fixed"], a["x"], a["v"], a["rho"], a["e"])
        eng.set_run_length(10 ** 6)
        eng.setup()
        eng.run(21)     # two rebuilds and the Shepard-filter step 20 (the FILTER instantiation)
        out[mode] = {f: eng.get(f) for f in FIELDS}
        assert eng.nbuilds >= 2
        eng.close()
    for f in FIELDS:
        assert np.array_equal(out["grid"][f], out[sched][f]), f
