"""The reference's own example decks AT SHIPPED SIZE through the drop-in boundary: examples/ssa-tsdpd/
lid_driven_cavity/Re1000_N200 (42 436 atoms), natural_convection/Ra_10E4 (42 436, one species, buoyancy + Dirichlet
forcing; as shipped with e = 1e-6 and with e = 0), fsi/fsi.lmp (14 036, mechanics style, elastic beam, sponge
buffers, periodic in x) and cell_polarization/case_1 (10 292, fsi style with a full list, reaction-diffusion in a
deformable ring) -- BASELINE.json configs 1-4.

tests/golden/make_decks.py (run by __graft_entry__.build() where /root/reference exists) writes the decks to
oracle/_ref/decks/ with the only edits a user without the VTK library needs: `dump vtk` -> `dump custom` of
the same columns (+ positions and forces), a bounded `run`.  Each deck is run through the UNMODIFIED reference
(oracle/_ref/lmp_serial) and through `lmp_cuda -sf cuda` (the /cuda style classes over the C ABI), the deck itself
untouched between the two.  Bar: every dumped column within 1e-10 of its max-norm over the run; the as-shipped
natural-convection deck carries the stochastic stress (e = 1e-6, clock()-seeded upstream, ~1e-8 of the forces) and
is compared at 1e-5.
"""
import os

import numpy as np
import pytest

from test_lammps_dropin import CUDA, REF, ROOT, read_dumps, read_thermo, run_deck

pytestmark = pytest.mark.gpu
DECKDIR = os.path.join(ROOT, "oracle", "_ref", "decks")
CASES = {"cavity_Re1000_N200": 1e-10, "natconv_Ra1e4_e0": 1e-10, "natconv_Ra1e4": 1e-5, "fsi": 1e-10,
         "cell_polarization_case1": 1e-10}


@pytest.mark.parametrize("name", sorted(CASES))
def test_shipped_deck_at_shipped_size(name, monkeypatch):
    path = os.path.join(DECKDIR, name + ".lmp")
    if not (os.path.exists(REF) and os.path.exists(CUDA) and os.path.exists(path)):
        pytest.skip("lmp_serial / lmp_cuda / oracle/_ref/decks not built (python -c 'import __graft_entry__ as g; g.build()')")
    deck = open(path).read()
    wd_ref, out_ref = run_deck(REF, deck, [])
    monkeypatch.setenv("SPHBVF_VERBOSE", "1")   # the engine prints its download statistics: proof that it ran
    wd_cuda, out_cuda = run_deck(CUDA, deck, ["-sf", "cuda"])
    assert "sphbvf:" in out_cuda, "the /cuda styles were not selected:\n" + out_cuda[-1500:]
    ref, got = read_dumps(wd_ref), read_dumps(wd_cuda)
    assert sorted(ref) == sorted(got) and len(ref) >= 3, (sorted(ref), sorted(got))
    cols = ref[min(ref)][0]
    first, last = ref[min(ref)][1], ref[max(ref)][1]
    natoms = len(first)
    solid_col = [k for k, c in enumerate(cols) if c in ("c_solidtagatom", "c_solid_tag")][0]
    xk, yk = cols.index("x"), cols.index("y")
    # forces on FIXED solids are never consumed and orientation dependent in the reference:
    # fixed = solid-tagged and not moved over the run
    fixed = (first[:, solid_col] == 1) & (first[:, xk] == last[:, xk]) & (first[:, yk] == last[:, yk])
    scale = {c: max(np.abs(ref[s][1][:, k]).max() for s in ref) for k, c in enumerate(cols)}
    tol = CASES[name]
    worst = {}
    for s in sorted(ref):
        a, b = ref[s][1], got[s][1]
        assert a.shape == b.shape, (name, s, a.shape, b.shape)
        assert np.array_equal(a[:, 0], b[:, 0]) and np.array_equal(a[:, 1], b[:, 1])
        for k, c in enumerate(cols[2:], start=2):
            x, y = a[:, k], b[:, k]
            if c in ("fx", "fy"):
                x, y = x[~fixed], y[~fixed]
            fin = np.isfinite(x)
            assert np.array_equal(fin, np.isfinite(y)), (name, s, c)
            err = float(np.abs(x[fin] - y[fin]).max() / max(scale[c], 1e-300)) if fin.any() else 0.0
            worst[c] = max(worst.get(c, 0.0), err)
    bad = {c: e for c, e in worst.items() if e > tol}
    print("shipped deck %s: %d atoms (%d fixed), dumps at %s, worst column errors %s" % (
        name, natoms, int(fixed.sum()), sorted(ref), {c: "%.1e" % e for c, e in worst.items()}))
    assert not bad, "%s: columns beyond %g: %s (all: %s)" % (name, tol, bad, worst)
    # thermo: same steps printed, temperature at print precision
    ta, tb = read_thermo(out_ref), read_thermo(out_cuda)
    if ta.size and ta.shape == tb.shape:
        assert np.array_equal(ta[:, 0], tb[:, 0])
