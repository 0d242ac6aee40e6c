"""The stochastic stress term (e != 0, pair_ssa_tsdpd_bvf_transport_velocity.cpp:403-431).

The reference's random forces cannot be reproduced (Marsaglia stream seeded from clock(), drawn in
half-list order), so parity is statistical, against a numpy restatement of the reference's
construction of the random matrix, plus the properties that make the gather formulation legal:
  * the random forces of a pair are equal and opposite (sum over all fluid atoms = 0 to rounding);
  * same seed and timestep -> bit-identical forces; another seed or timestep -> different ones;
  * per-component variance of f_rand matches sum_j pref_ij^2 Var[(Wn del)_l] (Monte Carlo of the
    reference's symmetric-traceless construction), and the mean is zero.
"""
import numpy as np
import pytest

from conftest import load_package

pytestmark = pytest.mark.gpu
KB = 1.3806504e-23


def lattice_meta(dim, n):
    d = 1.0 / n
    h = 2.6 * d if dim == 3 else 2.5 * d
    per = [1, 1, 1]
    return dict(dim=dim, periodic=per, boxlo=[0, 0, 0], boxhi=[1, 1, 1 if dim == 3 else d], ntypes=1, S=0, variant=0,
                skin=0.01 * h, every=1, delay=10, check=1, dt=1e-4, integrate_groupbit=1,
                types=[dict(mass=d ** dim, rho0=1.0, c0=10.0, G0=0.0)],
                pairs=[dict(i=1, j=1, eta=1e-2, h=h, cutc=h, kappa=[])], fixes=[]), d, h


def make_atoms(dim, n, rng):
    d = 1.0 / n
    g = (np.arange(n) + 0.5) * d
    if dim == 3:
        x = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    else:
        xy = np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2)
        x = np.concatenate([xy, np.zeros((len(xy), 1))], axis=1)
    x[:, :dim] += (rng.random((len(x), dim)) - 0.5) * 0.2 * d
    return np.ascontiguousarray(x)


def forces(pkg, meta, x, e, seed, step=0, random=True):
    n = len(x)
    eng = pkg.Engine(meta)
    if random:
        eng.set_random(KB, seed)
    eng.set_atoms(np.arange(1, n + 1), np.ones(n, np.int32), np.ones(n, np.int32), np.zeros(n, np.int32),
                  np.zeros(n, np.int32), x, np.zeros((n, 3)), np.ones(n), np.full(n, e))
    eng.set_timestep(step)
    eng.setup()
    f = eng.get("f")
    eng.close()
    return f


def wn_variance(dim, nsamp=400000, seed=5):
    """Var of the entries of the reference's symmetric traceless matrix (d x d Gaussians)."""
    rng = np.random.default_rng(seed)
    g = np.zeros((nsamp, 3, 3))
    g[:, :dim, :dim] = rng.standard_normal((nsamp, dim, dim))
    w = 0.5 * (g + g.transpose(0, 2, 1))
    tr = (w[:, 0, 0] + w[:, 1, 1] + w[:, 2, 2]) / dim
    for k in range(3):
        w[:, k, k] -= tr
    return w.var(axis=0)


@pytest.mark.parametrize("dim,n", [(2, 48), (3, 16)])
def test_random_stress_statistics_and_conservation(dim, n):
    pkg = load_package()
    rng = np.random.default_rng(dim)
    meta, d, h = lattice_meta(dim, n)
    x = make_atoms(dim, n, rng)
    e = 1e-3 / KB * 1e-9      # large enough that the random force dominates rounding, small vs nothing else
    f_det = forces(pkg, meta, x, e, 1, random=False)
    f1 = forces(pkg, meta, x, e, 1234)
    f1b = forces(pkg, meta, x, e, 1234)
    f2 = forces(pkg, meta, x, e, 99)
    f3 = forces(pkg, meta, x, e, 1234, step=7)
    assert np.array_equal(f1, f1b)
    fr = f1 - f_det
    assert np.abs(fr).max() > 0 and not np.array_equal(f1, f2) and not np.array_equal(f1, f3)
    # equal and opposite pair forces: total random force vanishes to rounding
    assert np.abs(fr.sum(axis=0)).max() <= 1e-10 * np.abs(fr).sum(axis=0).max()
    # expected variance per atom and component from the pair geometry
    from scipy.spatial import cKDTree
    prd = np.array(meta["boxhi"]) - np.array(meta["boxlo"])
    box = prd.copy()
    if dim == 2:
        box[2] = 1e9
    tree = cKDTree(x % box if dim == 3 else np.concatenate([x[:, :2] % 1.0, x[:, 2:]], axis=1), boxsize=box)
    pairs = tree.query_pairs(h, output_type="ndarray")
    dl = x[pairs[:, 0]] - x[pairs[:, 1]]
    dl -= np.round(dl / prd) * prd * np.array([1, 1, 1 if dim == 3 else 0])
    r = np.sqrt((dl ** 2).sum(axis=1))
    keep = r < h
    pairs, dl, r = pairs[keep], dl[keep], r[keep]
    cwfd = (-25.066903536973515383 / h ** 7) if dim == 3 else (-19.098593171027440292 / h ** 6)
    wfd = cwfd * (h - r) ** 2
    V = d ** dim
    pref2 = -4 * KB * e * V * V * wfd / meta["dt"] / (r + 0.01 * h) ** 2
    var_w = wn_variance(dim)
    var = np.zeros((len(x), 3))
    for l in range(dim):
        contrib = pref2 * (var_w[l, 0] * dl[:, 0] ** 2 + var_w[l, 1] * dl[:, 1] ** 2 + var_w[l, 2] * dl[:, 2] ** 2)
        np.add.at(var[:, l], pairs[:, 0], contrib)
        np.add.at(var[:, l], pairs[:, 1], contrib)
    z = fr[:, :dim] / np.sqrt(var[:, :dim])
    assert abs(z.mean()) < 5.0 / np.sqrt(z.size)
    assert abs(z.var() - 1.0) < 0.05, z.var()
