"""Shared helpers for the parity tests: fixture loading and the comparison metric."""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def fixture_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


def feed_atoms(engine, z):
    engine.set_atoms(z["init_tag"], z["init_type"], z["init_mask"], z["init_solid_tag"], z["init_fixed_tag"],
                     z["init_x"], z["init_v"], z["init_rho"], z["init_e"], z["init_C"], z["init_dev"])


def norm_err(a, ref):
    """max |a-ref| / max(|ref|_inf, tiny): SURVEY.md A.9 -- elementwise relative error is
    meaningless where lattice symmetry cancels sums to rounding noise."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    bad = ~(np.isfinite(a) == np.isfinite(ref))
    if bad.any():
        return np.inf
    fin = np.isfinite(ref)
    if not fin.any():
        return 0.0
    scale = max(np.abs(ref[fin]).max(), 1e-300)
    return float(np.abs(a[fin] - ref[fin]).max() / scale)
