"""sphbvf_ke_tensor (device sums of m v_a v_b, the accumulation of ComputeTemp) against numpy on the
downloaded velocities, and its run-to-run reproducibility (fixed summation order)."""
import numpy as np
import pytest

from common import feed_atoms, load_fixture
from conftest import load_package

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cavity_n50", "synth3d_n14", "fsi_nx20"])
def test_ke_tensor_matches_numpy(name):
    pkg = load_package()
    meta, z = load_fixture(name)
    eng = pkg.Engine(meta)
    feed_atoms(eng, z)
    eng.set_run_length(30)
    eng.setup()
    eng.run(13)
    v = eng.get("v")
    typ = z["init_type"]
    mass = np.array([0.0] + [t["mass"] for t in meta["types"]])[typ]
    mask = z["init_mask"]
    for groupbit in (1, 2):
        sel = (mask & groupbit) != 0
        want = np.array([(mass * v[:, a] * v[:, b])[sel].sum() for a, b in ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))])
        got = eng.ke_tensor(groupbit)
        again = eng.ke_tensor(groupbit)
        assert np.array_equal(got, again)
        scale = max(np.abs(want).max(), 1e-300)
        assert np.abs(got - want).max() <= 1e-12 * scale, (groupbit, got, want)
    eng.close()
