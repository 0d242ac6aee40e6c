"""Diagnostic (not a test): per-fixture, per-field worst normalised error of the CUDA library vs
the golden fixtures.  Usage on the GPU box: python tests/gpu_diag.py [fixture ...]"""
import sys

import numpy as np

from common import feed_atoms, fixture_names, load_fixture
from conftest import load_package
from refsnap import canonical_pairs

pkg = load_package()
for name in sys.argv[1:] or fixture_names():
    meta, z = load_fixture(name)
    try:
        eng = pkg.Engine(meta)
        feed_atoms(eng, z)
        eng.set_run_length(meta["nsteps"])
        eng.setup()
        step, worst, pairs_ok = 0, {}, []
        for s in meta["steps"]:
            if s > step:
                eng.run(s - step)
                step = s
            for f in meta["fields"]:
                ref = z["s%d_%s" % (s, f)]
                got = eng.get(f)
                scale = max(float(np.abs(z["s%d_%s" % (q, f)]).max()) if ref.size else 0.0 for q in meta["steps"])
                fin = np.isfinite(ref) & np.isfinite(got)
                e = np.abs(got[fin] - ref[fin]).max() / max(scale, 1e-300) if fin.any() else 0.0
                if not np.array_equal(np.isfinite(ref), np.isfinite(got)):
                    e = np.inf
                worst[f] = max(worst.get(f, 0.0), e)
            if s in meta["pair_steps"]:
                a, b = canonical_pairs(eng.pairs()), z["p%d" % s]
                pairs_ok.append((s, a.shape[0], b.shape[0], bool(a.shape == b.shape and (a == b).all())))
        print(name, "nghost", eng.nghost, "builds", eng.nbuilds, meta["build_steps"], "pairs", pairs_ok)
        print("   ", {k: "%.1e" % v for k, v in worst.items()})
        eng.close()
    except Exception as ex:  # noqa: BLE001
        print(name, "FAILED:", repr(ex))
