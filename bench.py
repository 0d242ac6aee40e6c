#!/usr/bin/env python
"""bench.py -- headline benchmark of the SPH-BVF timestep hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n LATTICE] [--impl ours|reference]

Workload (BASELINE.json configs[4]): synthetic 3D lid-driven cavity lattice,
n^3 simple-cubic particles (3-layer fixed BVF walls, moving lid, fluid jittered by 0.1 delta with
an analytic velocity/density perturbation so forces do not cancel), h = 2.6 delta, skin 0.01 h,
dt = 0.05 h / c0, transportVelocity pair + fix.  A "step" is ONE full timestep of the hot path
(initial_integrate -> neighbour decide/rebuild or halo -> fused density+force pair pass ->
final_integrate; final_integrate(n) and initial_integrate(n+1) run as one kernel when nothing is scheduled
between two steps) on device-resident state; rebuilds (every 10 steps with this skin) are inside
the timed region.  Weak scaling: n = 200 / 252 / 318 / 400 at 1 / 2 / 4 / 8 GPUs (8M atoms per GPU),
brick decomposition, one process per GPU (torchrun), NCCL halo inside libsphbvf.so.

Prints ONE JSON line (rank 0).  `value` = atoms x K / max-over-ranks device time (CUDA events on
the library's stream), inputs resident in HBM.  `e2e` = same metric through the C ABI with HOST
(pinned) buffers: per step the pair/integrator inputs are copied host->device and the results
device->host inside the timed region.  `roofline` is for the dominant kernel (the fused pair
kernel): algorithmic bytes 240 B/atom-step (124 + 116) over its CUDA-event time.
`cpu_baseline` = the UNMODIFIED reference (oracle/_ref/lmp_serial, built by oracle/Makefile from
/root/reference) on ALL host cores on a bounded sample (smaller n, same deck): the reference is MPI-parallel
only and the image has no MPI runtime, so one single-rank replica runs per core and the aggregate is reported
(an upper bound for one MPI job); if that binary did not travel, the plain-C oracle port is timed instead
(one core) and says so.

`lammps_dropin` (extra key) = the same deck as an UNMODIFIED LAMMPS input through `lmp_cuda -sf cuda` (the /cuda
style classes over the C ABI), LAMMPS' own Loop time: n=160 at N=1; at N>1 the weak-scaling lattice of the launch
(400^3 = 64 M atoms at N=8) driven by ONE LAMMPS process on all N GPUs, run by rank 0 after the timed ranks are gone.

--impl reference runs only that CPU arm (rank 0) and prints the same line shape.
"""
import argparse
import ctypes
import importlib.util
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
B_ALG_PAIR = 240.0      # algorithmic bytes / atom-step of the pair pass
B_ALG_STEP = 616.0      # whole step
F_ALG_PAIR_SURVEY = 1.2e4   # FP64 flop / atom-step, 3D bulk, SURVEY.md 8d's paper estimate (2.8e3 + 9.2e3)
# FP64 flop / atom-step the kernel really executes, re-derived as SURVEY 8(d) asks from ncu's
# smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on of one launch at 8.0 M atoms (see PROFILE below):
# (2 x dfma + dmul + dadd) / atoms.  Updated whenever a new --set full capture is committed under profiles/.
PROFILE = {"file": "profiles/r02p_pair_kernel_warp_schedule_n200.txt",
           "kernel": "pair_kernel<TV,0,1,UNIFORM,PERSIST> (gather form, persistent 192-thread CTAs whose warps draw 32-atom chunks, as shipped)",
           "flop_per_atom": 7.70e3, "fp64_inst_per_atom": 5.30e3, "dram_bytes_per_atom": 732.9,
           "l1_data_pipe_busy": 0.822, "fp64_pipe_busy": 0.491, "issue_active": 0.454, "warps_per_sm": 12.0}
PARITY_FIXTURES = ["synth3d_n14", "solid3d_mech_n10", "cavity_n50"]   # 3D non-periodic, 3D periodic with free solids, 2D
WEAK_N = {1: 200, 2: 252, 4: 318, 8: 400}


def load_package():
    name = "sphbvf_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "sph-bvf_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------
def cavity_meta(n):
    delta = 1.0 / (n - 6)
    h = 2.6 * delta
    lo, hi = -3 * delta, 1.0 + 3 * delta
    m = delta ** 3
    return dict(dim=3, periodic=[0, 0, 0], boxlo=[lo] * 3, boxhi=[hi] * 3, ntypes=2, S=0, variant=0,
                skin=0.01 * h, every=1, delay=10, check=1, dt=0.05 * h / 10.0, integrate_groupbit=1,
                types=[dict(mass=m, rho0=1.0, c0=10.0, G0=0.0)] * 2,
                pairs=[dict(i=i, j=j, eta=1e-2, h=h, cutc=h, kappa=[]) for i in (1, 2) for j in (1, 2) if j >= i],
                fixes=[], delta=delta, n=n)


def cavity_atoms(meta, sublo, subhi, seed=20261018):
    """Atoms of the global n^3 lattice that fall into the brick [sublo, subhi)."""
    n, delta = meta["n"], meta["delta"]
    lo = meta["boxlo"][0]
    ranges = []
    for k in range(3):
        c = lo + (np.arange(n) + 0.5) * delta
        idx = np.nonzero((c >= sublo[k]) & (c < subhi[k]))[0]
        ranges.append(idx)
    ix, iy, iz = np.meshgrid(ranges[0], ranges[1], ranges[2], indexing="ij")
    ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
    tag = (1 + ix + n * (iy + n * iz)).astype(np.int32)
    x = np.stack([lo + (ix + 0.5) * delta, lo + (iy + 0.5) * delta, lo + (iz + 0.5) * delta], axis=1)
    fluid = np.all((x > 0.0) & (x < 1.0), axis=1)
    typ = np.where(fluid, 1, 2).astype(np.int32)
    solid = (~fluid).astype(np.int32)
    lid = (~fluid) & (x[:, 1] > 1.0)
    v = np.zeros_like(x)
    v[lid, 0] = 1.0
    rho = np.ones(len(x))
    # perturbation keyed on the tag so every decomposition builds the same global state
    rng_u = (np.sin(tag.astype(np.float64)[:, None] * np.array([12.9898, 78.233, 37.719]) + seed % 1000) * 43758.5453)
    jit = (rng_u - np.floor(rng_u) - 0.5) * 0.2 * delta
    x[fluid] += jit[fluid]
    xf = x[fluid]
    v[fluid, 0] = 0.1 * np.sin(np.pi * xf[:, 0]) * np.cos(np.pi * xf[:, 1])
    v[fluid, 1] = -0.1 * np.cos(np.pi * xf[:, 0]) * np.sin(np.pi * xf[:, 1])
    rho[fluid] = 1.0 + 0.01 * np.sin(2 * np.pi * xf[:, 0]) * np.sin(2 * np.pi * xf[:, 1]) * np.sin(2 * np.pi * xf[:, 2])
    return dict(tag=tag, type=typ, mask=np.ones(len(x), np.int32), solid=solid, fixed=solid.copy(),
                x=np.ascontiguousarray(x), v=np.ascontiguousarray(v), rho=rho, e=np.zeros(len(x)))


REF_DECK = """
dimension 3
units si
atom_style ssa_tsdpd/atomic 0 0 0
boundary f f f
variable n equal {n}
variable delta equal 1.0/(v_n-6)
variable lo equal -3*v_delta
variable hi equal 1.0+3*v_delta
region domain block ${{lo}} ${{hi}} ${{lo}} ${{hi}} ${{lo}} ${{hi}} units box
create_box 2 domain
lattice sc ${{delta}} origin 0.5 0.5 0.5
create_atoms 2 box
region fluid_region block 0 1 0 1 0 1 units box
group fluid region fluid_region
set group fluid type 1
group solid subtract all fluid
region lid_region block ${{lo}} ${{hi}} 1 ${{hi}} ${{lo}} ${{hi}} units box
group lid region lid_region
variable m equal v_delta*v_delta*v_delta
mass * ${{m}}
set group all ssa_tsdpd/rho 1.0
set group all ssa_tsdpd/e 0.
set group solid ssa_tsdpd/solid_tag 1 fixed
variable h equal 2.6*v_delta
pair_style ssa_tsdpd/bvf/transportVelocity
pair_coeff * * 1.0 10.0 1e-2 ${{h}} ${{h}} 0.0
velocity lid set 1.0 0.0 0.0 units box
displace_atoms fluid random $(0.1*v_delta) $(0.1*v_delta) $(0.1*v_delta) 20261018 units box
variable ux atom 0.1*sin(PI*x)*cos(PI*y)
variable uy atom -0.1*cos(PI*x)*sin(PI*y)
velocity fluid set v_ux v_uy 0.0 units box
variable r atom 1.0+0.01*sin(2*PI*x)*sin(2*PI*y)*sin(2*PI*z)
set group fluid ssa_tsdpd/rho v_r
fix integration all ssa_tsdpd/bvf/transportVelocity
variable skin equal 0.01*v_h
neighbor ${{skin}} bin
variable dt equal 0.05*v_h/10.0
timestep ${{dt}}
thermo 1000
run {warm}
run {steps}
"""


def host_cores():
    try:
        return max(1, min(128, len(os.sched_getaffinity(0))))
    except AttributeError:
        return max(1, min(128, os.cpu_count() or 1))


def time_reference(n, steps, warm, procs=1):
    """atom-steps/s of the unmodified reference CPU path (Loop time of the 2nd run).  The reference is MPI-parallel
    only (no threaded variant of these styles) and the image has no MPI runtime, so "all host cores" is `procs`
    independent single-rank replicas of the same deck running at the same time, one per core; the aggregate is an
    upper bound for what one MPI job over the same cores could do (no halo exchange, no load imbalance)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
    if os.path.exists(exe):
        with tempfile.TemporaryDirectory() as wd:
            jobs = []
            for p in range(procs):
                sub = os.path.join(wd, "r%d" % p)
                os.makedirs(sub)
                with open(os.path.join(sub, "in.lmp"), "w") as fh:
                    fh.write(REF_DECK.format(n=n, steps=steps, warm=warm))
                jobs.append(subprocess.Popen([exe, "-in", "in.lmp", "-log", "none", "-echo", "none"], cwd=sub,
                                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
            total, atoms = 0.0, 0
            for j in jobs:
                so, se = j.communicate(timeout=3000)
                loops = re.findall(r"Loop time of ([0-9.eE+-]+) on (\d+) procs for (\d+) steps with (\d+) atoms", so)
                if j.returncode != 0 or not loops:
                    raise RuntimeError("lmp_serial failed: " + so[-400:] + se[-400:])
                t, _, st, atoms = loops[-1]
                total += int(atoms) * int(st) / float(t)
            return total, "reference", int(atoms)
    # the reference binary did not travel: time the plain-C restatement instead
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_api import Oracle
    meta = cavity_meta(n)
    a = cavity_atoms(meta, meta["boxlo"], meta["boxhi"])
    o = Oracle(meta)
    o.set_atoms(a["tag"], a["type"], a["mask"], a["solid"], a["fixed"], a["x"], a["v"], a["rho"], a["e"])
    o.setup()
    o.run(warm)
    t0 = time.perf_counter()
    o.run(steps)
    dt = time.perf_counter() - t0
    return len(a["tag"]) * steps / dt, "port", len(a["tag"])


def time_lmp_cuda(n, steps, warm, ngpu=1, timeout=1200):
    """atom-steps/s of REF_DECK run unmodified through lmp_cuda -sf cuda (Loop time of the 2nd run); None if not built"""
    exe = os.path.join(ROOT, "sph-bvf_b200", "lammps", "_build", "lmp_cuda")
    if not os.path.exists(exe):
        return None
    try:
        with tempfile.TemporaryDirectory() as wd:
            with open(os.path.join(wd, "in.lmp"), "w") as fh:
                fh.write(REF_DECK.format(n=n, steps=steps, warm=warm))
            env = dict(os.environ, SPHBVF_NGPU=str(ngpu))
            out = subprocess.run([exe, "-in", "in.lmp", "-log", "none", "-echo", "none", "-sf", "cuda"], cwd=wd, env=env,
                                 capture_output=True, text=True, timeout=timeout)
            loops = re.findall(r"Loop time of ([0-9.eE+-]+) on (\d+) procs for (\d+) steps with (\d+) atoms", out.stdout)
            if out.returncode != 0 or not loops:
                return {"value": None, "error": (out.stdout[-300:] + out.stderr[-300:])}
            t, _, st, atoms = loops[-1]
            return {"value": int(atoms) * int(st) / float(t), "unit": "atom-steps/s", "atoms": int(atoms), "steps": int(st),
                    "n_gpus": ngpu, "how": "unmodified LAMMPS deck through lmp_cuda -sf cuda, one LAMMPS process; LAMMPS' own "
                                           "Loop time (state device resident between the upload in setup and the download after the run)"}
    except Exception as ex:  # noqa: BLE001
        return {"value": None, "error": repr(ex)}


def parity_check(pkg, rank, world, dev, dist):
    """Golden fixtures of the unmodified reference across the `world` bricks of this launch (tests/brick_parity.py):
    fields within 1e-10 of the field's max-norm, pair sets bit-exact.  -> dict on rank 0, None elsewhere."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from brick_parity import run_fixture
    t0 = time.perf_counter()
    out = {"ranks": world, "tolerance": 1e-10, "fixtures": [], "max_err": 0.0, "pairs_equal": True, "ok": True}
    for name in PARITY_FIXTURES:
        errors, info, stats, pairs_equal = run_fixture(pkg, name, rank, world, dev, dist)
        if rank != 0:
            continue
        out["fixtures"].append({"name": name, "procgrid": list(info[2]), "steps": stats["steps"], "max_err": stats["max_err"],
                                "pairs": stats["pairs"], "pairs_equal": bool(pairs_equal), "rebuilds": info[5],
                                "errors": [list(map(str, e)) for e in errors[:4]]})
        out["max_err"] = max(out["max_err"], stats["max_err"])
        out["pairs_equal"] = out["pairs_equal"] and bool(pairs_equal)
        out["ok"] = out["ok"] and not errors
    flag = [out["ok"]]
    if world > 1:
        dist.broadcast_object_list(flag, src=0)
    if rank != 0:
        return {"ok": flag[0]} if not flag[0] else None
    out["seconds"] = time.perf_counter() - t0
    return out


def pair_kernel_name(eng):
    mode = eng.pair_mode() if hasattr(eng, "pair_mode") else "gather"
    return "pair_kernel (fused density/BVF + force pass), %s form" % mode


class ClockSampler(threading.Thread):
    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.stop_flag, self.rows, self.first, self.last = dev, False, [], 0, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def mark(self):
        self.first = len(self.rows)

    def stop(self):
        self.last = len(self.rows)
        self.stop_flag = True

    def summary(self):
        """median SM clock / throttle reasons over the timed region; if it was shorter than a sampling period, over the
        samples around it (the run is steady state from the warm-up on)"""
        rows = self.rows[self.first:self.last]
        window = "timed region"
        if len(rows) < 2:
            rows = self.rows[max(0, self.first - 3):(self.last or len(self.rows)) + 1]
            window = "timed region +- 0.3 s"
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in rows for k in range(4) if len(r) > 2 + k and r[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(rows), "samples_whole_run": len(self.rows), "window": window}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--n", type=int, default=0, help="lattice edge (default: weak-scaling table)")
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-n", type=int, default=40, help="lattice edge of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lammps", action="store_true", help="skip the lmp_cuda drop-in leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the golden-fixture parity check before the timed region")
    ap.add_argument("--lammps-n", type=int, default=160, help="lattice edge of the lmp_cuda drop-in leg")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.n or WEAK_N.get(max(world, args.gpus), 200)
    workload = "synthetic 3D cavity lattice %d^3 (%.1fM atoms), TV pair+fix, h=2.6delta, skin 0.01h, jittered" % (n, n ** 3 / 1e6)

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: EXACTLY the K timed and W warm-up steps that were asked for, on a lattice small enough for the
        # whole run to end within ~2.5 minutes (the reference does ~6.9e4 atom-steps/s per core; throughput is per
        # atom-step, so the sample size does not enter the metric)
        ref_steps, ref_warm = K, W
        ref_n = args.ref_n
        for cand in (args.ref_n, 32, 26, 20, 16):
            ref_n = min(cand, args.ref_n)
            if (K + W) * ref_n ** 3 / 6.5e4 <= 150.0:
                break
        cores = host_cores()
        val, kind, atoms = time_reference(ref_n, ref_steps, ref_warm, cores)
        if kind != "reference":
            cores = 1
        sample = ("%d independent single-rank replicas of the same deck at n=%d (%d atoms each), one per host core, %d timed "
                  "steps after %d warm-up: aggregate throughput, an upper bound for one MPI job on these cores (no MPI runtime "
                  "in the image, no threaded variant of these styles)" % (cores, ref_n, atoms, ref_steps, ref_warm))
        line = {"impl": "reference", "metric": "atom-steps/sec", "value": val, "unit": "atom-steps/s", "n_gpus": args.gpus,
                "steps": ref_steps, "warmup": ref_warm, "ms_per_step": 1e3 * atoms * cores / val, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload,
                           "cpu_sample": "%d replicas x %d^3 = %d atoms, %d timed steps after %d warm-up (throughput is per "
                                         "atom-step, so the sample size does not enter the metric)"
                                         % (cores, ref_n, atoms, ref_steps, ref_warm)},
                "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ONE JSON line on stdout: everything else that may write to file descriptor 1 (NCCL prints its version banner
    # there from C) goes to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    pkg = load_package()
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))

    # clocks are sampled from here on (warm-up included) so that the line carries more than one sample
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()

    # multi-rank correctness on record next to the throughput: golden fixtures of the unmodified reference run
    # across the SAME N bricks (migration, borders, halo over NCCL), fields reassembled by tag on rank 0
    parity = None
    if not args.no_parity:
        parity = parity_check(pkg, rank, world, dev, dist)
        if parity is not None and not parity["ok"]:
            if rank == 0:
                os.write(real_stdout, (json.dumps({"metric": "atom-steps/sec", "value": None, "n_gpus": world,
                                                   "error": "parity_check failed", "parity_check": parity}) + "\n").encode())
            sys.exit(3)

    meta = cavity_meta(n)
    prd = [meta["boxhi"][k] - meta["boxlo"][k] for k in range(3)]
    grid = (ctypes.c_int * 3)(1, 1, 1)
    pkg.lib().sphbvf_proc_grid(world, 3, (ctypes.c_double * 3)(*prd), ctypes.byref(grid))
    procgrid = tuple(grid)
    eng = pkg.Engine(meta, device=dev, procgrid=procgrid, rank=rank, nranks=world)
    sublo, subhi = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
    pkg.lib().sphbvf_brick_bounds(ctypes.byref(eng.cfg), rank, ctypes.byref(sublo), ctypes.byref(subhi))
    atoms = cavity_atoms(meta, list(sublo), list(subhi))
    nloc0 = len(atoms["tag"])
    if world > 1:
        ident = [pkg.capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        eng.comm_init(ident[0])
    eng.set_atoms(atoms["tag"], atoms["type"], atoms["mask"], atoms["solid"], atoms["fixed"], atoms["x"], atoms["v"],
                  atoms["rho"], atoms["e"])
    eng.set_run_length(10 ** 9)
    eng.setup()
    stream = torch.cuda.ExternalStream(pkg.lib().sphbvf_stream(eng.h), device=torch.device("cuda", dev))
    natoms = n ** 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn(steps)
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    world_done = False
    eng.run(W)
    if rank == 0:
        sampler.mark()          # the summary's median is over the samples taken from here to stop()
    launches0 = eng.launch_count
    ms = timed(eng.run, K)
    launches = eng.launch_count - launches0
    if rank == 0:
        sampler.stop()
    value = natoms * K / (ms * 1e-3)

    # per-kernel device times (second pass, CUDA events around each kernel family on the lib's stream)
    eng.profiling(True)
    eng.run(K)
    fam = {}
    for idx, nm in enumerate(["pair", "initial_integrate", "final_integrate", "neighbor_rebuild", "pack_halo", "fixes",
                              "final_initial_pack_fused"]):
        t, c = eng.kernel_ms(idx)
        fam[nm] = {"ms": t, "launches": c}
    eng.profiling(False)
    pair_ms = fam["pair"]["ms"] / max(K, 1)
    peaks, peak_kind = measured_peaks()
    achieved = B_ALG_PAIR * eng.nlocal / (pair_ms * 1e-3) / 1e9 if pair_ms > 0 else 0.0
    # FP64 vector peak: 64 FMA/clk/SM (ncu sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained
    # = 9472 inst/cycle on 148 SMs) x 2 flop x max SM clock
    fp64_peak = float(os.environ.get("SPHBVF_FP64_PEAK_TFLOPS", "0") or 0) or 148 * 64 * 2 * 1.965e9 / 1e12
    f_alg = PROFILE["flop_per_atom"]
    roof = {"bound": "hbm", "kernel": pair_kernel_name(eng), "achieved": achieved,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            # dram__bytes_read.sum + dram__bytes_write.sum per launch: NOT measured live (ncu cannot run inside a timed
            # bench); bytes/atom of the committed --set full capture x the atoms of this launch
            "traffic": PROFILE["dram_bytes_per_atom"] * eng.nlocal, "traffic_source": "profile: " + PROFILE["file"],
            "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6650 GB/s",
            "algorithmic_bytes_per_atom_step": B_ALG_PAIR, "pair_ms_per_step": pair_ms,
            # FP64 work: flop the kernel executes per atom-step by ncu's op_d{fma,mul,add} counters (fma = 2) and, beside
            # it, SURVEY 8(d)'s paper estimate
            "fp64_flop_per_atom_step": f_alg, "fp64_flop_source": "ncu op_dfma/dmul/dadd counters, " + PROFILE["file"],
            "fp64_tflops": f_alg * eng.nlocal / (pair_ms * 1e-3) / 1e12 if pair_ms > 0 else 0.0,
            "fp64_tflops_survey_estimate": F_ALG_PAIR_SURVEY * eng.nlocal / (pair_ms * 1e-3) / 1e12 if pair_ms > 0 else 0.0,
            "note": "the pair pass is FP64-pipe / on-chip-gather bound (AI ~ 30 flop/B, SURVEY.md 8d); HBM fraction reported "
                    "as the contract asks, fp64_frac is the meaningful roof",
            "ncu": {k: PROFILE[k] for k in ("l1_data_pipe_busy", "fp64_pipe_busy", "issue_active", "warps_per_sm")}}
    roof["ncu"]["source"] = PROFILE["file"] + " (" + PROFILE["kernel"] + ", one --set full capture, not live)"
    if fp64_peak:
        roof["fp64_peak_tflops"] = fp64_peak
        roof["fp64_frac"] = roof["fp64_tflops"] / fp64_peak

    # e2e: host-authoritative stepping through the C ABI with pinned host buffers, every rank: per
    # step the integrator/pair inputs go host->device and the step's results device->host (rows in
    # device order, which the download of the previous step established)
    e2e = None
    if not args.no_e2e:
        F = pkg.FIELDS
        ins = ["x", "v", "vest", "rho", "rhoI"]
        outs = ["x", "v", "vest", "rho", "rhoI", "f", "drho", "phi"]
        ncol = {"x": 3, "v": 3, "vest": 3, "f": 3}
        cap = eng.nlocal + eng.nlocal // 8 + 4096
        hbuf = {k: torch.empty((cap, ncol.get(k, 1)), dtype=torch.float64).pin_memory() for k in set(ins + outs)}
        for k in ins:
            eng.download_local_ptr(F[k], hbuf[k].data_ptr(), cap)
        moved = [0, 0]

        def e2e_steps(steps):
            for _ in range(steps):
                nl = eng.nlocal
                for k in ins:
                    eng.upload_local_ptr(F[k], hbuf[k].data_ptr(), nl)
                eng.step_pieces()
                nl = eng.nlocal
                for k in outs:
                    eng.download_local_ptr(F[k], hbuf[k].data_ptr(), cap)
                moved[0] += sum(8 * nl * ncol.get(k, 1) for k in ins)
                moved[1] += sum(8 * nl * ncol.get(k, 1) for k in outs)

        ke = max(3, min(K, 10))
        e2e_steps(2)
        moved[:] = [0, 0]
        ms_e = timed(e2e_steps, ke)
        tot = torch.tensor(moved, device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot)
        e2e = {"value": natoms * ke / (ms_e * 1e-3), "unit": "atom-steps/s", "h2d_bytes_per_step": int(tot[0].item() / ke),
               "d2h_bytes_per_step": int(tot[1].item() / ke), "steps": ke,
               "mode": "host-authoritative through the C ABI: upload x,v,vest,rho,rhoI -> one device step (hooks) -> "
                       "download x,v,vest,rho,rhoI,f,drho,phi; pinned host memory; every step, every rank"}

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        try:
            cores = host_cores()
            val, kind, ca = time_reference(args.ref_n, 3, 1, cores)
            if kind != "reference":
                cores = 1
            cpu = {"value": val, "unit": "atom-steps/s", "cores": cores, "kind": kind,
                   "sample": "%d independent single-rank replicas of the same deck at n=%d (%d atoms each), one per host core, "
                             "3 timed steps after 1: aggregate throughput (upper bound for one MPI job on these cores)"
                             % (cores, args.ref_n, ca)}
        except Exception as ex:  # noqa: BLE001
            cpu = {"value": None, "unit": "atom-steps/s", "cores": 1, "kind": "reference", "sample": "failed: %r" % (ex,)}

    # drop-in leg: the same deck as an unmodified LAMMPS input through `lmp_cuda -sf cuda` (one LAMMPS process, the
    # /cuda style classes of sph-bvf_b200/lammps over the C ABI): LAMMPS' own "Loop time" of the second `run`.
    # N = 1: n = 160 (4.1 M atoms).  N > 1: the weak-scaling lattice of this launch (8 M atoms per GPU, 64 M at N = 8)
    # driven by ONE LAMMPS process with a worker thread per GPU, after the ranks of this launch have released theirs.
    dropin = None
    if not args.no_lammps:
        if world == 1:
            dropin = time_lmp_cuda(args.lammps_n, 50, 10)
        else:
            eng.close()
            eng = None
            dist.barrier()
            dist.destroy_process_group()
            world_done = True
            if rank != 0:
                return
            dropin = time_lmp_cuda(n, 30, 10, ngpu=world, timeout=420)

    if rank == 0:
        time.sleep(0.3)
        line = {"metric": "atom-steps/sec", "value": value, "unit": "atom-steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload, "atoms": natoms, "atoms_per_gpu": natoms // world, "procgrid": list(procgrid),
                           "l2": "inputs larger than L2 (state >> 126 MB), no explicit flush",
                           "rebuild": "every 10 steps (delay 10, skin 0.01h), inside the timed region"},
                "gpu_launches": int(launches), "kernels": fam, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
                "parity_check": parity,
                "clocks": sampler.summary(), "lammps_dropin": dropin}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if eng is not None:
        eng.close()
    if world > 1 and not world_done:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
