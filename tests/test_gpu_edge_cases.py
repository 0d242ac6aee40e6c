"""Edge cases of the hot path through the C ABI, against the plain-C oracle on the same inputs: an empty system, a
single atom, atoms with no neighbour at all next to a dense cluster that fills one cell (ragged neighbour counts),
and coincident atoms."""
import numpy as np
import pytest

from conftest import load_package
from oracle_api import Oracle
from refsnap import canonical_pairs

pytestmark = pytest.mark.gpu
TOL = 1e-10


def box_meta(dim, h=0.05):
    return dict(dim=dim, periodic=[0, 0, 0 if dim == 3 else 1], boxlo=[0.0, 0.0, 0.0], boxhi=[1.0, 1.0, 1.0 if dim == 3 else h],
                ntypes=2, S=0, variant=0, skin=0.1 * h, every=1, delay=2, check=1, dt=2e-6, integrate_groupbit=1,
                types=[dict(mass=1e-3, rho0=1.0, c0=10.0, G0=0.0)] * 2,
                pairs=[dict(i=i, j=j, eta=1e-2, h=h, cutc=h, kappa=[]) for i in (1, 2) for j in (1, 2) if j >= i], fixes=[])


def run_both(meta, x, v, typ=None, solid=None, nsteps=7):
    pkg = load_package()
    n = len(x)
    typ = np.ones(n, np.int32) if typ is None else typ
    solid = np.zeros(n, np.int32) if solid is None else solid
    out = []
    for cls in (pkg.Engine, Oracle):
        e = cls(meta)
        e.set_atoms(np.arange(1, n + 1, dtype=np.int32), typ, np.ones(n, np.int32), solid, solid.copy(), x, v,
                    np.ones(n), np.zeros(n))
        e.set_run_length(nsteps)
        e.setup()
        first = canonical_pairs(e.pairs())
        e.run(nsteps)
        out.append((canonical_pairs(e.pairs()), {f: e.get(f) for f in ("x", "v", "rho", "f", "drho", "phi", "number_density")}, first))
        e.close()
    return out


def compare(out):
    (pa, fa, ia), (pb, fb, ib) = out
    assert np.array_equal(pa, pb) and np.array_equal(ia, ib)
    for f in fa:
        a, b = fa[f], fb[f]
        assert a.shape == b.shape and np.array_equal(np.isfinite(a), np.isfinite(b)), f
        fin = np.isfinite(b)
        if fin.any():
            err = np.abs(a[fin] - b[fin]).max() / max(np.abs(b[fin]).max(), 1e-300)
            assert err <= TOL, (f, err)


@pytest.mark.parametrize("dim", [2, 3])
def test_empty_system(dim):
    pkg = load_package()
    e = pkg.Engine(box_meta(dim))
    z = np.zeros((0, 3))
    e.set_atoms(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32),
                np.zeros(0, np.int32), z, z, np.zeros(0), np.zeros(0))
    e.set_run_length(5)
    e.setup()
    e.run(5)
    assert e.nlocal == 0 and len(e.pairs()) == 0
    e.close()


@pytest.mark.parametrize("dim", [2, 3])
def test_single_atom(dim):
    x = np.array([[0.5, 0.5, 0.5 if dim == 3 else 0.0]])
    v = np.array([[0.3, -0.2, 0.1 if dim == 3 else 0.0]])
    out = run_both(box_meta(dim), x, v)
    compare(out)
    assert len(out[0][0]) == 0 and np.all(out[0][1]["f"] == 0.0)


@pytest.mark.parametrize("dim", [2, 3])
def test_ragged_neighbour_counts(dim):
    """isolated atoms (no neighbour), a tight cluster inside one cell (dozens of neighbours each), a wall patch and two
    coincident atoms (r = 0: the kernel and its derivative are finite there)"""
    rng = np.random.default_rng(20261018 + dim)
    h = 0.05
    iso = rng.uniform(0.05, 0.95, size=(40, 3)) * np.array([1, 1, 1.0 if dim == 3 else 0.0])
    cluster = np.array([0.52, 0.47, 0.5 if dim == 3 else 0.0]) + rng.uniform(-0.25, 0.25, size=(60, 3)) * h * np.array([1, 1, 1.0 if dim == 3 else 0.0])
    wall = np.array([[0.3 + 0.4 * h * i, 0.2 + 0.4 * h * j, 0.5 if dim == 3 else 0.0] for i in range(6) for j in range(3)])
    twin = np.array([[0.8, 0.8, 0.5 if dim == 3 else 0.0]] * 2)
    x = np.concatenate([iso, cluster, wall, twin])
    n = len(x)
    v = rng.normal(0, 0.05, size=(n, 3)) * np.array([1, 1, 1.0 if dim == 3 else 0.0])
    typ = np.ones(n, np.int32)
    solid = np.zeros(n, np.int32)
    sl = slice(len(iso) + len(cluster), len(iso) + len(cluster) + len(wall))
    typ[sl] = 2
    solid[sl] = 1
    v[sl] = 0.0
    meta = box_meta(dim, h)
    for t in meta["types"]:
        t["mass"] = (0.4 * h) ** dim          # density ~ 1 in the wall patch, ~ 30 in the cluster
    out = run_both(meta, x, v, typ, solid, nsteps=0)   # the state right after setup: list + one pair pass
    compare(out)
    counts = np.bincount(out[0][2].ravel(), minlength=n + 1)[1:]   # list of the setup build
    assert counts.min() == 0 and counts.max() >= 40
