"""The lazily fused integrator (final_integrate(n) + initial_integrate(n+1) + pack in one kernel, see
include/sphbvf.h at sphbvf_final_integrate) must be BIT-identical to the separate kernels: same
fields after the same number of steps, through rebuilds, Shepard-filter steps (every 20), fixes in
post_integrate / end_of_step (which must block the pack or the fusion), elastic solids and species."""
import os

import numpy as np
import pytest

from common import feed_atoms, load_fixture
from conftest import load_package

pytestmark = pytest.mark.gpu

FIELDS = ["x", "v", "vest", "rho", "rhoI", "f", "drho", "phi", "nw", "number_density", "ddv"]


def run(name, nsteps, fuse, chunks):
    pkg = load_package()
    meta, z = load_fixture(name)
    os.environ["SPHBVF_NO_FUSE"] = "0" if fuse else "1"
    try:
        eng = pkg.Engine(meta)
    finally:
        os.environ.pop("SPHBVF_NO_FUSE", None)
    feed_atoms(eng, z)
    eng.set_run_length(nsteps)
    eng.setup()
    done = 0
    for c in chunks:           # sphbvf_run flushes a pending final_integrate at its end: vary the chunking
        eng.run(c)
        done += c
    eng.run(nsteps - done)
    out = {f: eng.get(f) for f in FIELDS}
    if meta["S"]:
        out["C"] = eng.get("C")
    out["dev"] = eng.get("dev")
    fam = [eng.kernel_ms(k)[1] for k in range(7)]
    eng.close()
    return out, fam


@pytest.mark.parametrize("name", ["cavity_n20", "synth3d_n14", "cavity_mech_n20", "fsi_nx20", "natconv_n40",
                                  "yeast_nx40", "solid3d_fsi_n10", "solid3d_mech_n10"])
def test_fused_equals_unfused_bitwise(name):
    nsteps = 24
    a, fam_a = run(name, nsteps, True, [7, 1])
    b, fam_b = run(name, nsteps, False, [24])
    assert fam_a[6] > 0, "the fused kernel never ran"
    assert fam_b[6] == 0
    for f in a:
        assert np.array_equal(a[f], b[f], equal_nan=True), (name, f, float(np.nanmax(np.abs(a[f] - b[f]))))


def test_pieces_fuse_and_downloads_flush():
    """Through the fine-grained entry points (what the /cuda host classes call): a download between two
    steps sees the finished step (the deferred final_integrate is flushed), and stepping on still matches."""
    pkg = load_package()
    meta, z = load_fixture("cavity_n20")
    res = []
    for fuse in (True, False):
        os.environ["SPHBVF_NO_FUSE"] = "0" if fuse else "1"
        try:
            eng = pkg.Engine(meta)
        finally:
            os.environ.pop("SPHBVF_NO_FUSE", None)
        feed_atoms(eng, z)
        eng.set_run_length(30)
        eng.setup()
        snaps = []
        for s in range(15):
            eng.step_pieces()
            if s in (3, 4, 11):
                snaps.append((eng.get("v"), eng.get("rho"), eng.get("phi")))
        snaps.append((eng.get("x"), eng.get("v"), eng.get("rho")))
        res.append(snaps)
        eng.close()
    for sa, sb in zip(*res):
        for u, w in zip(sa, sb):
            assert np.array_equal(u, w)
