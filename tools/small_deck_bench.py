"""Per-step cost of a small 2D deck (BASELINE configs 1-2: lid-driven cavity N200 = 42 436 atoms) through the C ABI:
where a step is a handful of launches, the interesting number is microseconds per step, not bandwidth."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = bench.load_package()


def cavity2d(n):
    delta = 1.0 / (n - 6)
    h = 2.5 * delta
    lo, hi = -3 * delta, 1.0 + 3 * delta
    meta = dict(dim=2, periodic=[0, 0, 1], boxlo=[lo, lo, 0.0], boxhi=[hi, hi, delta], ntypes=2, S=0, variant=0,
                skin=0.01 * h, every=1, delay=10, check=1, dt=0.05 * h / 10.0, integrate_groupbit=1,
                types=[dict(mass=delta ** 2, rho0=1.0, c0=10.0, G0=0.0)] * 2,
                pairs=[dict(i=i, j=j, eta=1e-2, h=h, cutc=h, kappa=[]) for i in (1, 2) for j in (1, 2) if j >= i], fixes=[])
    ix, iy = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    x = np.stack([lo + (ix.ravel() + 0.5) * delta, lo + (iy.ravel() + 0.5) * delta, np.zeros(n * n)], axis=1)
    fluid = np.all((x[:, :2] > 0) & (x[:, :2] < 1), axis=1)
    v = np.zeros_like(x)
    v[(~fluid) & (x[:, 1] > 1), 0] = 1.0
    v[fluid, 0] = 0.1 * np.sin(np.pi * x[fluid, 0]) * np.cos(np.pi * x[fluid, 1])
    v[fluid, 1] = -0.1 * np.cos(np.pi * x[fluid, 0]) * np.sin(np.pi * x[fluid, 1])
    a = dict(tag=np.arange(1, n * n + 1, dtype=np.int32), type=np.where(fluid, 1, 2).astype(np.int32),
             mask=np.ones(n * n, np.int32), solid=(~fluid).astype(np.int32), fixed=(~fluid).astype(np.int32),
             x=x, v=v, rho=np.ones(n * n), e=np.zeros(n * n))
    return meta, a


for n in (56, 106, 206, 412):
    meta, a = cavity2d(n)
    eng = pkg.Engine(meta)
    eng.set_atoms(a["tag"], a["type"], a["mask"], a["solid"], a["fixed"], a["x"], a["v"], a["rho"], a["e"])
    eng.set_run_length(10 ** 9)
    eng.setup()
    eng.run(200)
    l0 = eng.launch_count
    t0 = time.perf_counter()
    eng.run(2000)
    dt = time.perf_counter() - t0
    print("cavity2d n=%d: %d atoms, %.1f us/step, %.3g atom-steps/s, %.1f launches/step, %d rebuilds" % (
        n, n * n, dt / 2000 * 1e6, n * n * 2000 / dt, (eng.launch_count - l0) / 2000.0, eng.nbuilds))
    # second pass with CUDA events around every kernel family (serialises nothing, but adds two event records per
    # family call): where the device time of a step goes
    eng.profiling(True)
    eng.run(400)
    eng.sync()
    fam = ["pair", "initial_integrate", "final_integrate", "neighbor_rebuild", "pack_halo", "fixes", "final_initial_pack_fused"]
    print("   device us/step:", ", ".join("%s %.2f (%d)" % (f, eng.kernel_ms(k)[0] * 1e3 / 400, eng.kernel_ms(k)[1]) for k, f in enumerate(fam)))
    eng.close()
