"""Error behaviour of the C ABI on the GPU: the conditions the reference reports through
error->one/all come back as status codes + messages, never as a crash or a silent wrong answer
(SURVEY.md 8b "Error convention")."""
import numpy as np
import pytest

from common import feed_atoms, load_fixture
from conftest import load_package

pytestmark = pytest.mark.gpu


def engine(pkg, name="cavity_n20"):
    meta, z = load_fixture(name)
    eng = pkg.Engine(meta)
    feed_atoms(eng, z)
    return eng, meta, z


def test_calls_out_of_order():
    pkg = load_package()
    meta, z = load_fixture("cavity_n20")
    eng = pkg.Engine(meta)
    with pytest.raises(pkg.SphbvfError, match="set_atoms"):
        eng.setup()
    feed_atoms(eng, z)
    with pytest.raises(pkg.SphbvfError, match="before setup"):
        eng.run(1)
    eng.close()


def test_missing_pair_coefficients():
    pkg = load_package()
    meta, z = load_fixture("cavity_n20")
    meta = dict(meta, pairs=[p for p in meta["pairs"] if not (p["i"] == 1 and p["j"] == 2)])
    eng = pkg.Engine(meta)
    feed_atoms(eng, z)
    with pytest.raises(pkg.SphbvfError, match="coeffs"):      # pair_...:1034-1036
        eng.setup()
    eng.close()


def test_non_finite_position_is_reported():
    pkg = load_package()
    eng, meta, z = engine(pkg)
    eng.setup()
    x = eng.get("x")
    x[5, 0] = np.nan
    eng.put("x", x)
    with pytest.raises(pkg.SphbvfError, match="Non-numeric"):  # nbin.cpp:120
        eng.build_neighbors()
    eng.close()


def test_lost_atom_is_reported():
    pkg = load_package()
    eng, meta, z = engine(pkg)
    eng.setup()
    x = eng.get("x")
    x[7, 1] = meta["boxhi"][1] + 10.0        # far outside a fixed boundary
    eng.put("x", x)
    with pytest.raises(pkg.SphbvfError, match="Lost atoms"):   # thermo.cpp:436-450
        eng.build_neighbors()
    eng.close()


def test_bad_arguments():
    pkg = load_package()
    meta, z = load_fixture("cavity_n20")
    with pytest.raises(pkg.SphbvfError):
        pkg.Engine(dict(meta, ntypes=9, types=meta["types"] * 5))
    eng = pkg.Engine(meta)
    t = z["init_type"].copy()
    t[3] = 7
    with pytest.raises(pkg.SphbvfError, match="type"):
        eng.set_atoms(z["init_tag"], t, z["init_mask"], z["init_solid_tag"], z["init_fixed_tag"], z["init_x"], z["init_v"],
                      z["init_rho"], z["init_e"], z["init_C"], z["init_dev"])
    with pytest.raises(pkg.SphbvfError, match="unknown field"):
        eng._ck(pkg.lib().sphbvf_download(eng.h, 99, None))
    eng.close()


def test_neighbour_capacity_grows_transparently():
    """A compressed blob has far more neighbours than the density-based first guess: the library
    must regrow its list storage (the reference raises 'Neighbor list overflow' only beyond 2000)."""
    pkg = load_package()
    meta, z = load_fixture("synth3d_n14")
    x = z["init_x"].copy()
    c = x.mean(axis=0)
    x = c + (x - c) * np.where((np.abs(x - c) < 0.2).all(axis=1), 0.55, 1.0)[:, None]
    eng = pkg.Engine(meta)
    eng.set_atoms(z["init_tag"], z["init_type"], z["init_mask"], z["init_solid_tag"], z["init_fixed_tag"], x, z["init_v"],
                  z["init_rho"], z["init_e"], z["init_C"], z["init_dev"])
    eng.setup()
    from oracle_api import Oracle
    orc = Oracle(meta)
    orc.set_atoms(z["init_tag"], z["init_type"], z["init_mask"], z["init_solid_tag"], z["init_fixed_tag"], x, z["init_v"],
                  z["init_rho"], z["init_e"], z["init_C"], z["init_dev"])
    orc.setup()
    from refsnap import canonical_pairs
    assert np.array_equal(canonical_pairs(eng.pairs()), canonical_pairs(orc.pairs()))
    a, b = eng.get("number_density"), orc.get("number_density")
    assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max()
    eng.close()


def test_uploaded_e_and_type_reach_the_kernel_selection():
    """set_atoms(e = 0) followed by upload(e != 0) must switch the stochastic term on (the flags that select the pair
    kernel come from the device state, not from what set_atoms saw), and an uploaded type change must reach the
    packed neighbour-list entries (forced rebuild): both runs must equal a run that was given the data up front."""
    from common import feed_atoms, load_fixture
    pkg = load_package()
    meta, z = load_fixture("mixedh2d_n30")
    n = meta["natoms"]
    e_new = np.full(n, 1e-6)
    type_new = z["init_type"].copy()
    type_new[::7] = 1 + (type_new[::7] % meta["ntypes"])

    def arrays(e, typ):
        return (z["init_tag"], typ, z["init_mask"], z["init_solid_tag"], z["init_fixed_tag"], z["init_x"], z["init_v"],
                z["init_rho"], e, z["init_C"], z["init_dev"])

    ref = pkg.Engine(meta)
    ref.set_atoms(*arrays(e_new, type_new))
    ref.set_random(1.380649e-23, 7)
    ref.set_run_length(6)
    ref.setup()
    ref.run(6)
    got = pkg.Engine(meta)
    got.set_atoms(*arrays(np.zeros(n), z["init_type"]))
    got.set_random(1.380649e-23, 7)
    got.put("e", e_new)
    got.put("type", type_new)
    got.set_run_length(6)
    got.setup()
    got.run(6)
    for f in ("x", "v", "f", "rho"):
        assert np.array_equal(ref.get(f), got.get(f)), f
    # after setup: upload during the run, the next step re-derives the flags and rebuilds
    late = pkg.Engine(meta)
    late.set_atoms(*arrays(np.zeros(n), z["init_type"]))
    late.set_random(1.380649e-23, 7)
    late.set_run_length(6)
    late.setup()
    nb0 = late.nbuilds
    late.put("e", e_new)
    late.put("type", type_new)
    late.run(1)
    assert late.nbuilds == nb0 + 1
    f_late = late.get("f")
    chk = pkg.Engine(meta)
    chk.set_atoms(*arrays(np.zeros(n), z["init_type"]))
    chk.set_random(1.380649e-23, 7)
    chk.set_run_length(6)
    chk.setup()
    chk.run(1)
    assert not np.array_equal(f_late, chk.get("f")), "the uploaded e / type had no effect"
    for e in (ref, got, late, chk):
        e.close()
