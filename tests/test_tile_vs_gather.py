"""The two traversals of the pair pass against each other.  They share one arithmetic body (PairAcc::visit in
csrc/kernels_pair.cu) and the same pair sets, so they may differ only by summation order (eight partial sums per
atom combined by a butterfly in the tile form, one running sum in the gather form) and by the last bit of rho_j in
the solid-viscosity term: every output must agree to 1e-12 of its max-norm, after one pass and after a trajectory
through rebuilds, for every variant / species / solids combination the fixtures reach, including the virial pass
and a Shepard-filter step."""
import numpy as np
import pytest

from common import feed_atoms, fixture_names, load_fixture
from conftest import load_package
from refsnap import canonical_pairs

pytestmark = pytest.mark.gpu


def run_mode(pkg, meta, z, mode, nsteps, monkeypatch):
    monkeypatch.setenv("SPHBVF_PAIR", mode)
    eng = pkg.Engine(meta)
    feed_atoms(eng, z)
    eng.set_run_length(max(nsteps, 1))
    eng.setup()
    out = {"mode0": eng.pair_mode()}
    out["step0"] = {f: eng.get(f) for f in meta["fields"]}
    out["pairs0"] = canonical_pairs(eng.pairs())
    if nsteps:
        eng.run(nsteps)
    out["stepN"] = {f: eng.get(f) for f in meta["fields"]}
    out["pairsN"] = canonical_pairs(eng.pairs())
    out["nbuilds"] = eng.nbuilds
    # the ghost part of the virial needs a pair pass on the current state
    eng.pair_compute()
    out["virial"] = eng.virial()
    eng.close()
    return out


@pytest.mark.parametrize("name", fixture_names())
def test_tile_form_equals_gather_form(name, monkeypatch):
    pkg = load_package()
    meta, z = load_fixture(name)
    nsteps = 22   # one Shepard-filter step (20) and, in most fixtures, a rebuild
    a = run_mode(pkg, meta, z, "tile", nsteps, monkeypatch)
    b = run_mode(pkg, meta, z, "gather", nsteps, monkeypatch)
    assert a["mode0"] == "tile" and b["mode0"] == "gather"
    assert a["nbuilds"] == b["nbuilds"]
    assert np.array_equal(a["pairs0"], b["pairs0"]) and np.array_equal(a["pairsN"], b["pairsN"])
    worst = {}
    for key, tol in (("step0", 1e-12), ("stepN", 1e-10)):   # 22 steps amplify rounding differences a little
        for f in meta["fields"]:
            x, y = a[key][f], b[key][f]
            fin = np.isfinite(y)
            assert np.array_equal(np.isfinite(x), fin), (name, key, f)
            scale = max(np.abs(y[fin]).max() if fin.any() else 0.0, 1e-300)
            err = float(np.abs(x[fin] - y[fin]).max() / scale) if fin.any() else 0.0
            worst[(key, f)] = err
            assert err <= tol, (name, key, f, err)
    va, vb = a["virial"], b["virial"]
    assert np.abs(va - vb).max() <= 1e-9 * max(np.abs(vb).max(), 1e-300), (va, vb)
    print("tile vs gather %s: worst %s" % (name, max(worst.items(), key=lambda kv: kv[1])))
