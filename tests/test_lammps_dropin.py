"""Drop-in test of the LAMMPS-style /cuda classes (sph-bvf_b200/lammps/): the SAME input deck is run
through the UNMODIFIED reference (oracle/_ref/lmp_serial, plain styles, CPU) and through lmp_cuda
(the reference fork + the /cuda classes + libsphbvf.so) with `-sf cuda`, i.e. without touching the
deck, and the `dump custom` files are compared column by column.

The decks below are ours (written for this test, small enough for seconds of CPU time); they use
the package's own input-script API exactly as the decks in the reference's examples/ssa-tsdpd do:
atom_style ssa_tsdpd/atomic, set ssa_tsdpd/*, pair_style ssa_tsdpd/bvf/<variant>, fix ssa_tsdpd/bvf/<variant>,
fix ssa_tsdpd/{buoyancy,forcing,buffer}, fix setforce, compute ssa_tsdpd/*/atom, dump custom, thermo.
Bar: every dumped column within 1e-10 of its max-norm over the run.
"""
import glob
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
CUDA = os.path.join(ROOT, "sph-bvf_b200", "lammps", "_build", "lmp_cuda")
TOL = 1e-10
# thermo `Press` of the wall-bounded transportVelocity decks against the reference: see _check_deck
# measured on a B200 (gpurun_out/r2o_pytest.log): cavity2d 7.8e-4, cavity3d 6.0e-2, natconv2d 1.12e-1, react2d 1.3e-3
PRESS_ARTEFACT_BOUND = {"cavity2d": 1.6e-3, "cavity3d": 0.12, "natconv2d": 0.23, "react2d": 2.6e-3}

CAVITY2D = """
dimension 2
units si
atom_style ssa_tsdpd/atomic 0 0 0
boundary f f p
variable n equal 26
variable d equal 1.0/(v_n-6)
variable lo equal -3*v_d
variable hi equal 1.0+3*v_d
region box block ${lo} ${hi} ${lo} ${hi} 0 ${d} units box
create_box 2 box
lattice sq ${d} origin 0.5 0.5 0.0
create_atoms 2 box
region inner block 0 1 0 1 0 ${d} units box
group fluid region inner
set group fluid type 1
group solid subtract all fluid
region top block ${lo} ${hi} 1 ${hi} 0 ${d} units box
group lid region top
mass * $(v_d*v_d)
set group all ssa_tsdpd/rho 1.0
set group all ssa_tsdpd/e 0.
set group solid ssa_tsdpd/solid_tag 1 fixed
variable h equal 2.5*v_d
pair_style ssa_tsdpd/bvf/transportVelocity
pair_coeff * * 1.0 10.0 1e-2 ${h} ${h} 0.0
velocity lid set 1.0 0.0 0.0 units box
variable ux atom 0.8*sin(PI*x)*cos(PI*y)
variable uy atom -0.8*cos(PI*x)*sin(PI*y)
velocity fluid set v_ux v_uy 0.0 units box
fix integ all ssa_tsdpd/bvf/transportVelocity
fix hold lid setforce 0.0 0.0 0.0
neighbor $(0.01*v_h) bin
timestep 1e-4
compute crho all ssa_tsdpd/rho/atom
compute cphi all ssa_tsdpd/phi/atom
compute cp all ssa_tsdpd/p/atom
dump d all custom 7 dump.*.txt id type x y vx vy fx fy c_crho c_cphi c_cp
dump_modify d sort id format float %.17g
thermo 7
run 28
"""

CAVITY3D = """
dimension 3
units si
atom_style ssa_tsdpd/atomic 0 0 0
boundary f f f
variable n equal 14
variable d equal 1.0/(v_n-6)
variable lo equal -3*v_d
variable hi equal 1.0+3*v_d
region box block ${lo} ${hi} ${lo} ${hi} ${lo} ${hi} units box
create_box 2 box
lattice sc ${d} origin 0.5 0.5 0.5
create_atoms 2 box
region inner block 0 1 0 1 0 1 units box
group fluid region inner
set group fluid type 1
group solid subtract all fluid
region top block ${lo} ${hi} 1 ${hi} ${lo} ${hi} units box
group lid region top
mass * $(v_d*v_d*v_d)
set group all ssa_tsdpd/rho 1.0
set group all ssa_tsdpd/e 0.
set group solid ssa_tsdpd/solid_tag 1 fixed
variable h equal 2.6*v_d
pair_style ssa_tsdpd/bvf/transportVelocity
pair_coeff * * 1.0 10.0 1e-2 ${h} ${h} 0.0
velocity lid set 1.0 0.0 0.0 units box
displace_atoms fluid random $(0.1*v_d) $(0.1*v_d) $(0.1*v_d) 4711 units box
variable ux atom 0.8*sin(PI*x)*cos(PI*y)
variable uy atom -0.8*cos(PI*x)*sin(PI*y)
velocity fluid set v_ux v_uy 0.0 units box
variable r atom 1.0+0.01*sin(2*PI*x)*sin(2*PI*y)*sin(2*PI*z)
set group fluid ssa_tsdpd/rho v_r
fix integ all ssa_tsdpd/bvf/transportVelocity
neighbor $(0.01*v_h) bin
timestep $(0.05*v_h/10.0)
compute crho all ssa_tsdpd/rho/atom
compute cphi all ssa_tsdpd/phi/atom
dump d all custom 11 dump.*.txt id type x y z vx vy vz fx fy fz c_crho c_cphi
dump_modify d sort id format float %.17g
thermo 11
run 22
"""

# heated cavity: one species, Boussinesq buoyancy, Dirichlet walls through fix forcing, two runs
NATCONV2D = """
dimension 2
units si
atom_style ssa_tsdpd/atomic 1 0 0
boundary f f p
variable n equal 30
variable d equal 1.0/(v_n-6)
variable lo equal -3*v_d
variable hi equal 1.0+3*v_d
region box block ${lo} ${hi} ${lo} ${hi} 0 ${d} units box
create_box 2 box
lattice sq ${d} origin 0.5 0.5 0.0
create_atoms 2 box
region inner block 0 1 0 1 0 ${d} units box
group fluid region inner
set group fluid type 1
group solid subtract all fluid
mass * $(v_d*v_d)
set group all ssa_tsdpd/rho 1.0
set group all ssa_tsdpd/e 0.
set group all ssa_tsdpd/C 0 0.5
set group solid ssa_tsdpd/solid_tag 1 fixed
variable h equal 2.5*v_d
pair_style ssa_tsdpd/bvf/transportVelocity
pair_coeff * * 1.0 5.0 0.0266 ${h} ${h} 0.0 0.0375
fix integ all ssa_tsdpd/bvf/transportVelocity
fix buoy fluid ssa_tsdpd/buoyancy boussinesq/sdpd -1.0 1 0 0.5
fix hot all ssa_tsdpd/forcing tsdpd 3 0 rectangle $(-1.5*v_d) 0.5 $(1.5*v_d) 2.0 1.0
fix cold all ssa_tsdpd/forcing tsdpd 3 0 circle 1.0 0.5 $(4.2*v_d) 0.0
neighbor $(0.3*v_h) bin
timestep 2e-4
compute crho all ssa_tsdpd/rho/atom
compute cphi all ssa_tsdpd/phi/atom
compute cc all ssa_tsdpd/C/atom 0
dump d all custom 10 dump.*.txt id type x y vx vy fx fy c_crho c_cphi c_cc
dump_modify d sort id format float %.17g
thermo 10
run 20
run 20
"""

# channel with an elastic beam: mechanics style, periodic in x, free solid with shear modulus, sponge fixes
FSI2D = """
dimension 2
units si
atom_style ssa_tsdpd/atomic 0 0 0
boundary p f p
variable d equal 0.05
variable lo equal -3*v_d
variable hi equal 1.0+3*v_d
region box block 0 2.0 ${lo} ${hi} 0 ${d} units box
create_box 3 box
lattice sq ${d} origin 0.5 0.5 0.0
create_atoms 1 box
region inner block 0 2.0 0 1 0 ${d} units box
group chan region inner
group walls subtract all chan
set group walls type 3
region beam_region block 0.9 1.1 0 0.6 0 ${d} units box
group beam region beam_region
set group beam type 2
group fluid subtract chan beam
mass 1 $(v_d*v_d)
mass 2 $(v_d*v_d*8.0)
mass 3 $(v_d*v_d)
set group all ssa_tsdpd/rho 1.0
set group beam ssa_tsdpd/rho 8.0
set group all ssa_tsdpd/e 0.
set group beam ssa_tsdpd/solid_tag 1 free
set group walls ssa_tsdpd/solid_tag 1 fixed
variable h equal 3.0*v_d
pair_style ssa_tsdpd/bvf/mechanics
pair_coeff 1 1 1.0 1.0 1e-2 ${h} ${h} 0.0
pair_coeff 1 2 1.0 1.0 1e-2 ${h} ${h} 0.0
pair_coeff 1 3 1.0 1.0 1e-2 ${h} ${h} 0.0
pair_coeff 2 2 8.0 3.0 1e-2 ${h} ${h} 40.0
pair_coeff 2 3 8.0 3.0 1e-2 ${h} ${h} 40.0
pair_coeff 3 3 1.0 1.0 1e-2 ${h} ${h} 0.0
variable ux atom 0.1*y*(1.0-y)*4.0
velocity fluid set v_ux 0.0 0.0 units box
fix integ all ssa_tsdpd/bvf/mechanics
fix sponge_vx fluid ssa_tsdpd/buffer velocity x 1 0 0.2 0.5 0.2 0.5 0.1
fix sponge_vy fluid ssa_tsdpd/buffer velocity x 1 1 0.2 0.5 0.2 0.5 0.0
fix sponge_rho fluid ssa_tsdpd/buffer density x 1 0 0.2 0.5 0.2 0.5 1.0
neighbor $(0.3*v_h) bin
timestep 2e-3
compute crho all ssa_tsdpd/rho/atom
compute cphi all ssa_tsdpd/phi/atom
# stress = -Pnew + dev: upstream SUMS the assigned Pnew over periodic ghost images in its reverse
# communication (SURVEY.md D.7), so it is only meaningful away from periodic faces -> beam group
compute sxx beam ssa_tsdpd/stress/atom 0 0
compute sxy beam ssa_tsdpd/stress/atom 0 1
# run 0 before the dump is defined: upstream creates ghosts before setup_pre_force sets vest = v
# (verlet.cpp:118-132), so the very first setup with moving atoms at a periodic face computes forces
# from stale ghost velocities (SURVEY.md D.9)
run 0
dump d all custom 8 dump.*.txt id type x y vx vy fx fy c_crho c_cphi c_sxx c_sxy
dump_modify d sort id format float %.17g
thermo 8
run 24
"""

# ring-shaped elastic wall in a doubly periodic box: fsi style (full list), species softening the wall
RING2D = """
dimension 2
units si
atom_style ssa_tsdpd/atomic 1 0 0
boundary p p p
variable d equal 0.05
region box block 0 2.0 0 2.0 0 ${d} units box
create_box 3 box
lattice sq ${d} origin 0.5 0.5 0.0
create_atoms 1 box
region outer sphere 1.0 1.0 0.0 0.6 units box
region hole sphere 1.0 1.0 0.0 0.4 units box
group disc region outer
group core region hole
group ring subtract disc core
set group ring type 2
set group core type 3
mass 1 $(v_d*v_d)
mass 2 $(v_d*v_d*2.0)
mass 3 $(v_d*v_d)
set group all ssa_tsdpd/rho 1.0
set group ring ssa_tsdpd/rho 2.0
set group all ssa_tsdpd/e 0.
set group all ssa_tsdpd/C 0 0.0
set group ring ssa_tsdpd/solid_tag 1 free
variable h equal 3.0*v_d
pair_style ssa_tsdpd/bvf/fsi
pair_coeff 1 1 1.0 2.0 1e-2 ${h} ${h} 0.0 1e-3
pair_coeff 1 2 1.0 2.0 1e-2 ${h} ${h} 0.0 1e-3
pair_coeff 1 3 1.0 2.0 1e-2 ${h} ${h} 0.0 1e-3
pair_coeff 2 2 2.0 4.0 1e-2 ${h} ${h} 30.0 1e-3
pair_coeff 2 3 2.0 4.0 1e-2 ${h} ${h} 30.0 1e-3
pair_coeff 3 3 1.0 2.0 1e-2 ${h} ${h} 0.0 1e-3
variable ur atom 0.05*(x-1.0)
variable vr atom 0.05*(y-1.0)
velocity core set v_ur v_vr 0.0 units box
fix integ all ssa_tsdpd/bvf/fsi
fix src ring ssa_tsdpd/forcing tsdpd 2 0 rectangle 1.0 0.5 0.3 0.12 1.0
neighbor $(0.3*v_h) bin
timestep 1e-3
compute crho all ssa_tsdpd/rho/atom
compute cphi all ssa_tsdpd/phi/atom
compute cc all ssa_tsdpd/C/atom 0
compute syy ring ssa_tsdpd/stress/atom 1 1
run 0
dump d all custom 6 dump.*.txt id type x y vx vy fx fy c_crho c_cphi c_cc c_syy
dump_modify d sort id format float %.17g
thermo 6
run 24
"""

# three species with mass-action reactions A + B -> C, C -> A and a constant source of B, and a CFL-
# controlled timestep: fix ssa_tsdpd/chem_rxn_mass_action and fix dt/adaptive (no shipped deck uses
# them)
REACT2D = """
dimension 2
units si
atom_style ssa_tsdpd/atomic 3 0 0
boundary f f p
variable n equal 26
variable d equal 1.0/(v_n-6)
variable lo equal -3*v_d
variable hi equal 1.0+3*v_d
region box block ${lo} ${hi} ${lo} ${hi} 0 ${d} units box
create_box 2 box
lattice sq ${d} origin 0.5 0.5 0.0
create_atoms 2 box
region inner block 0 1 0 1 0 ${d} units box
group fluid region inner
set group fluid type 1
group solid subtract all fluid
mass * $(v_d*v_d)
set group all ssa_tsdpd/rho 1.0
set group all ssa_tsdpd/e 0.
variable ca atom 0.5+0.5*sin(PI*x)
variable cb atom 0.5+0.5*cos(PI*y)
set group all ssa_tsdpd/C 0 v_ca
set group all ssa_tsdpd/C 1 v_cb
set group all ssa_tsdpd/C 2 0.1
set group solid ssa_tsdpd/solid_tag 1 fixed
variable h equal 2.5*v_d
pair_style ssa_tsdpd/bvf/transportVelocity
pair_coeff * * 1.0 5.0 1e-2 ${h} ${h} 0.0 0.02 0.01 0.005
variable ux atom 0.6*sin(PI*x)*cos(PI*y)
variable uy atom -0.6*cos(PI*x)*sin(PI*y)
velocity fluid set v_ux v_uy 0.0 units box
fix integ all ssa_tsdpd/bvf/transportVelocity
fix rxn1 fluid ssa_tsdpd/chem_rxn_mass_action 0.5 2 0 1 1 2
fix rxn2 fluid ssa_tsdpd/chem_rxn_mass_action 0.7 1 2 1 0
fix src fluid ssa_tsdpd/chem_rxn_mass_action 0.05 0 1 1
fix cfl fluid dt/adaptive 1 1e-5 1e-3 0.005 ${d}
neighbor $(0.3*v_h) bin
timestep 2e-4
compute crho all ssa_tsdpd/rho/atom
compute c0 all ssa_tsdpd/C/atom 0
compute c1 all ssa_tsdpd/C/atom 1
compute c2 all ssa_tsdpd/C/atom 2
dump d all custom 8 dump.*.txt id type x y vx vy fx fy c_crho c_c0 c_c1 c_c2
dump_modify d sort id format float %.17g
thermo_style custom step temp epair emol etotal press dt
thermo 8
run 32
"""

DECKS = {"react2d": REACT2D, "cavity2d": CAVITY2D, "cavity3d": CAVITY3D, "natconv2d": NATCONV2D, "fsi2d": FSI2D, "ring2d": RING2D}


def read_dumps(wd):
    out = {}
    for path in sorted(glob.glob(os.path.join(wd, "dump.*.txt"))):
        lines = open(path).read().splitlines()
        step = int(lines[1])
        k = next(i for i, ln in enumerate(lines) if ln.startswith("ITEM: ATOMS"))
        cols = lines[k].split()[2:]
        data = np.array([[float(v) for v in ln.split()] for ln in lines[k + 1:]])
        out[step] = (cols, data)
    return out


def read_thermo(stdout):
    """rows of the thermo table(s): Step Temp E_pair E_mol TotEng Press"""
    rows, on = [], False
    for ln in stdout.splitlines():
        w = ln.split()
        if w[:2] == ["Step", "Temp"]:
            on = True
            continue
        if on:
            if ln.startswith("Loop time"):
                on = False
                continue
            try:
                rows.append([float(v) for v in w])
            except ValueError:
                pass
    return np.array(rows)


def run_deck(exe, deck, extra):
    wd = tempfile.mkdtemp(prefix="sphbvf_deck_")
    with open(os.path.join(wd, "in.lmp"), "w") as fh:
        fh.write(deck)
    r = subprocess.run([exe, "-in", "in.lmp", "-log", "log.lammps", "-echo", "none"] + extra, cwd=wd, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, "%s failed:\n%s\n%s" % (exe, r.stdout[-3000:], r.stderr[-2000:])
    return wd, r.stdout


@pytest.mark.parametrize("name", sorted(DECKS))
def test_deck_unchanged_with_sf_cuda(name):
    _check_deck(name)


@pytest.mark.parametrize("name", ["cavity3d", "cavity2d", "fsi2d", "ring2d", "natconv2d", "react2d"])
def test_deck_unchanged_on_two_gpus(name, monkeypatch):
    """One LAMMPS process (one MPI rank) driving two GPUs: the engine splits the atoms into bricks, one context and
    one worker thread per GPU, NCCL halo / migration inside the library; same decks, same bar."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("SPHBVF_NGPU", "2")
    out = _check_deck(name)
    assert "sphbvf: 2 GPUs" in out, out[-1500:]


def _check_deck(name):
    if not (os.path.exists(REF) and os.path.exists(CUDA)):
        pytest.skip("lmp_serial / lmp_cuda not built (make -C oracle ref; make -C sph-bvf_b200/lammps)")
    wd_ref, out_ref = run_deck(REF, DECKS[name], [])
    wd_cuda, out_cuda = run_deck(CUDA, DECKS[name], ["-sf", "cuda"])
    # thermo output (temperature from the synced host velocities, pressure from the device virial,
    # sphbvf_virial == Pair::virial_fdotr_compute): printed with 5-6 significant digits
    ta, tb = read_thermo(out_ref), read_thermo(out_cuda)
    assert ta.shape == tb.shape and ta.shape[0] >= 3 and ta.shape[1] >= 6, (ta.shape, tb.shape)
    if ta.shape[1] == 7:     # the deck prints the (adaptive) timestep as a 7th column
        assert len(set(ta[:, 6])) > 1, "dt/adaptive did not change the timestep"
        assert np.abs(ta[:, 6] - tb[:, 6]).max() <= 2e-5 * ta[:, 6].max(), (ta[:, 6], tb[:, 6])
    assert np.array_equal(ta[:, 0], tb[:, 0])
    for col, what in ((1, "Temp"), (5, "Press")):
        scale = np.abs(ta[:, col]).max()
        assert scale > 0, what
        # Press = (kinetic term + sum over local AND ghost atoms of x.f) / volume (Pair::virial_fdotr_compute).
        # Upstream's forces on SOLID partners are not the mirror image of the forces on their fluid
        # neighbours (different viscosity model; in transportVelocity also the un-flipped pressure switch,
        # SURVEY.md A.5), so which half of such a pair lands on which atom -- and with it sum x.f --
        # depends on half-list orientation.  The device virial (sphbvf_virial) is the orientation-free
        # value, checked exactly against the oracle in tests/test_gpu_parity.py; against upstream it can
        # agree only where no fluid-solid pair matters: tight for the ring deck (solid away from the
        # periodic faces), loose for the channel deck (walls cross a periodic face), not at all for the
        # wall-bounded transportVelocity decks.
        #
        # Wall-bounded transportVelocity decks: not skipped but held to the measured size of that artefact (max over
        # the thermo rows of |P_cuda - P_ref| / max |P_ref|, printed below; bounds = twice the values measured on a
        # B200 in round 2), so that a regression of the device virial beyond it still fails.
        if what == "Press" and name.startswith(("cavity", "natconv", "react")):
            err = np.abs(ta[:, col] - tb[:, col]).max() / scale
            print("thermo Press, %s: orientation artefact of the reference = %.3e of max |P|" % (name, err))
            assert err <= PRESS_ARTEFACT_BOUND[name], (name, err, ta[:, col], tb[:, col])
            continue
        tol = 2e-3 if (what == "Press" and name == "fsi2d") else 2e-5
        if what == "Press" and name == "fsi2d":
            ta, tb = ta[1:], tb[1:]        # step 0 of the first setup: stale ghost velocities upstream (D.9)
        assert np.abs(ta[:, col] - tb[:, col]).max() <= tol * scale, (name, what, ta[:, col], tb[:, col])
    ref, got = read_dumps(wd_ref), read_dumps(wd_cuda)
    assert sorted(ref) == sorted(got) and len(ref) >= 3, (sorted(ref), sorted(got))
    cols = ref[min(ref)][0]
    scale = {c: max(np.abs(ref[s][1][:, k]).max() for s in ref) for k, c in enumerate(cols)}
    worst = {}
    for s in sorted(ref):
        a, b = ref[s][1], got[s][1]
        assert a.shape == b.shape, (name, s, a.shape, b.shape)
        assert np.array_equal(a[:, 0], b[:, 0]) and np.array_equal(a[:, 1], b[:, 1])
        # forces on FIXED solids are never consumed and orientation dependent in the reference
        fixed = np.isin(a[:, 1], [2]) if name.startswith(("cavity", "natconv", "react")) else np.isin(a[:, 1], [3]) if name == "fsi2d" else np.zeros(len(a), bool)
        for k, c in enumerate(cols[2:], start=2):
            x, y = a[:, k], b[:, k]
            if c in ("fx", "fy", "fz"):
                x, y = x[~fixed], y[~fixed]
            fin = np.isfinite(x)
            assert np.array_equal(fin, np.isfinite(y)), (name, s, c)
            err = np.abs(x[fin] - y[fin]).max() / max(scale[c], 1e-300) if fin.any() else 0.0
            worst[c] = max(worst.get(c, 0.0), err)
    bad = {c: e for c, e in worst.items() if e > TOL}
    assert not bad, "%s: columns beyond %g: %s (all: %s)" % (name, TOL, bad, worst)
    return out_cuda


# ---------------------------------------------------------------------------------------------
# short-horizon field statistics (BASELINE.json north_star: "cavity centreline velocity, Nusselt
# number matching within 1e-6"), computed by the same post-processing on both outputs
# ---------------------------------------------------------------------------------------------
def _final_dump(deck, exe, extra, nsteps):
    deck = deck.replace("run 28", "run %d" % nsteps).replace("run 20\nrun 20", "run %d" % nsteps)
    deck = deck.replace("custom 7 ", "custom %d " % nsteps).replace("custom 10 ", "custom %d " % nsteps)
    wd, _ = run_deck(exe, deck, extra)
    d = read_dumps(wd)
    cols, data = d[max(d)]
    return {c: data[:, k] for k, c in enumerate(cols)}


def test_cavity_centreline_velocity_statistic():
    if not (os.path.exists(REF) and os.path.exists(CUDA)):
        pytest.skip("lmp_serial / lmp_cuda not built")
    nsteps = 400
    prof = []
    for exe, extra in ((REF, []), (CUDA, ["-sf", "cuda"])):
        f = _final_dump(CAVITY2D, exe, extra, nsteps)
        d = 1.0 / 20
        fluid = f["type"] == 1
        strip = fluid & (np.abs(f["x"] - 0.5) < d)
        bins = np.clip((f["y"][strip] / d).astype(int), 0, 19)
        # u(y) on the vertical centreline and v(x) on the horizontal one (Ghia-style profiles)
        u = np.bincount(bins, weights=f["vx"][strip], minlength=20) / np.maximum(np.bincount(bins, minlength=20), 1)
        strip = fluid & (np.abs(f["y"] - 0.5) < d)
        bins = np.clip((f["x"][strip] / d).astype(int), 0, 19)
        v = np.bincount(bins, weights=f["vy"][strip], minlength=20) / np.maximum(np.bincount(bins, minlength=20), 1)
        prof.append(np.concatenate([u, v]))
    scale = np.abs(prof[0]).max()
    assert scale > 1e-3
    assert np.abs(prof[0] - prof[1]).max() / scale < 1e-6


def test_heated_cavity_nusselt_statistic():
    if not (os.path.exists(REF) and os.path.exists(CUDA)):
        pytest.skip("lmp_serial / lmp_cuda not built")
    nsteps = 400
    nu = []
    for exe, extra in ((REF, []), (CUDA, ["-sf", "cuda"])):
        f = _final_dump(NATCONV2D, exe, extra, nsteps)
        d = 1.0 / 24
        fluid = f["type"] == 1
        # wall heat flux of the hot (left) wall: -dC/dx from the first two fluid columns, averaged in y,
        # over the conductive flux (C_hot - C_cold) / L = 1
        c1 = f["c_cc"][fluid & (f["x"] < d)].mean()
        c2 = f["c_cc"][fluid & (f["x"] > d) & (f["x"] < 2 * d)].mean()
        nu.append(-(c2 - c1) / d)
    assert abs(nu[0]) > 1e-3
    assert abs(nu[0] - nu[1]) / abs(nu[0]) < 1e-6, nu


def test_heated_cavity_with_stochastic_term_as_shipped():
    """The reference's natural-convection decks set ssa_tsdpd/e = 1e-6 (SI units): the random stress is
    then ~1e-8 of the deterministic force and unreproducible upstream (clock() seed).  The /cuda
    styles apply their counter-based version of the term; fields must agree to that noise level."""
    if not (os.path.exists(REF) and os.path.exists(CUDA)):
        pytest.skip("lmp_serial / lmp_cuda not built")
    deck = NATCONV2D.replace("set group all ssa_tsdpd/e 0.", "set group all ssa_tsdpd/e 1e-6")
    assert deck != NATCONV2D
    wd_ref, _ = run_deck(REF, deck, [])
    wd_cuda, _ = run_deck(CUDA, deck, ["-sf", "cuda"])
    ref, got = read_dumps(wd_ref), read_dumps(wd_cuda)
    assert sorted(ref) == sorted(got)
    cols = ref[min(ref)][0]
    for s in sorted(ref):
        a, b = ref[s][1], got[s][1]
        fluid = a[:, 1] == 1
        for k, c in enumerate(cols[2:], start=2):
            x, y = a[:, k], b[:, k]
            if c in ("fx", "fy"):
                x, y = x[fluid], y[fluid]
            scale = max(np.abs(ref[q][1][:, k]).max() for q in ref)
            assert np.abs(x - y).max() <= 1e-5 * max(scale, 1e-300), (s, c)


# ---------------------------------------------------------------------------------------------
# output path: thermo-only steps are served from the device (compute temp/cuda =
# sphbvf_ke_tensor, pressure from sphbvf_virial) without downloading the per-atom arrays; dump steps
# still see current host data
# ---------------------------------------------------------------------------------------------
def test_thermo_only_steps_need_no_download(monkeypatch):
    if not (os.path.exists(REF) and os.path.exists(CUDA)):
        pytest.skip("lmp_serial / lmp_cuda not built (make -C oracle ref; make -C sph-bvf_b200/lammps)")
    deck = CAVITY2D.replace("dump d all custom 7 ", "dump d all custom 21 ").replace("thermo 7", "thermo 3").replace("run 28", "run 42")
    wd_ref, out_ref = run_deck(REF, deck, [])
    monkeypatch.setenv("SPHBVF_VERBOSE", "1")
    wd_lazy, out_lazy = run_deck(CUDA, deck, ["-sf", "cuda"])
    monkeypatch.setenv("SPHBVF_OUTPUT", "full")
    wd_full, out_full = run_deck(CUDA, deck, ["-sf", "cuda"])
    stats = {}
    for tag, out in (("lazy", out_lazy), ("full", out_full)):
        ln = [l for l in out.splitlines() if l.startswith("sphbvf:") and "full downloads" in l]
        assert ln, out[-2000:]
        m = re.search(r"sphbvf: (\d+) full downloads, (\d+) output steps served from the device, (\d+) device kinetic-energy "
                      r"reductions, (\d+) bytes device->host", ln[-1])
        assert m, ln[-1]
        stats[tag] = tuple(int(v) for v in m.groups())   # full downloads, steps from the device, KE reductions, bytes
    # 15 output steps (0, 3, ..., 42); dumps at 0, 21, 42; the final download of Fix::post_run is not counted here
    assert stats["lazy"][1] >= 11 and stats["lazy"][0] == 0, stats
    assert stats["full"][1] == 0 and stats["full"][0] >= 14, stats
    assert stats["lazy"][2] >= 11, stats
    # the three dump steps copy the columns the dump lists (x, v, f and, through the /cuda per-atom computes, rho and
    # phi: 11 of the 43 doubles per atom this variant owns), the twelve thermo-only steps copy nothing
    natoms = 26 * 26
    assert stats["lazy"][3] == 3 * natoms * 8 * 11, (stats, natoms)
    assert stats["full"][3] >= 15 * natoms * 8 * 43, stats
    ta, tl, tf = read_thermo(out_ref), read_thermo(out_lazy), read_thermo(out_full)
    assert ta.shape == tl.shape == tf.shape and ta.shape[0] == 15
    scale = np.abs(ta[:, 1]).max()
    assert np.abs(ta[:, 1] - tl[:, 1]).max() <= 2e-5 * scale, (ta[:, 1], tl[:, 1])
    # device reduction vs host loop over the downloaded velocities: same numbers at print precision
    assert np.abs(tf[:, 1:] - tl[:, 1:]).max() <= 2e-5 * np.abs(tf[:, 1:]).max()
    ref, got = read_dumps(wd_ref), read_dumps(wd_lazy)
    assert sorted(ref) == sorted(got) == [0, 21, 42]
    for s in ref:
        a, b = ref[s][1], got[s][1]
        for k, c in enumerate(ref[s][0]):
            if c in ("fx", "fy"):
                continue
            sc = max(np.abs(a[:, k]).max(), 1e-300)
            assert np.abs(a[:, k] - b[:, k]).max() <= TOL * sc, (s, c)


# ---------------------------------------------------------------------------------------------
# atom_style ssa_tsdpd/atomic/cuda (picked by -sf cuda): the host arrays of the pair-sweep outputs are allocated when
# something on the host first asks for them
# ---------------------------------------------------------------------------------------------
def _host_mirror_stats(out):
    ln = [l for l in out.splitlines() if l.startswith("sphbvf: host mirrors")]
    assert ln, out[-2000:]
    m = re.search(r"host mirrors of (\d+) of the (\d+) pair-sweep output arrays were allocated .* (\d+) bytes of host arrays per atom slot", ln[-1])
    assert m, ln[-1]
    return tuple(int(v) for v in m.groups())


def test_lazy_host_mirrors(monkeypatch):
    if not (os.path.exists(REF) and os.path.exists(CUDA)):
        pytest.skip("lmp_serial / lmp_cuda not built (make -C oracle ref; make -C sph-bvf_b200/lammps)")
    monkeypatch.setenv("SPHBVF_VERBOSE", "1")
    # (a) the dump lists positions, velocities and rho only: no output array is ever mirrored, < 300 B per atom on the
    # host (upstream's style: ~ 830 B per atom)
    lean = CAVITY3D.replace("vx vy vz fx fy fz c_crho c_cphi", "vx vy vz c_crho")
    assert lean != CAVITY3D
    # (b) phi through its /cuda compute and drho through property/atom (AtomVec::pack_property_atom): two mirrors
    two = CAVITY3D.replace("compute cphi all ssa_tsdpd/phi/atom\n", "compute cphi all ssa_tsdpd/phi/atom\ncompute pa all property/atom drho\n")
    two = two.replace("vx vy vz fx fy fz c_crho c_cphi", "vx vy vz c_crho c_cphi c_pa")
    assert two != CAVITY3D
    for deck, nmirrors in ((lean, 0), (two, 2)):
        wd_ref, _ = run_deck(REF, deck, [])
        wd_cuda, out = run_deck(CUDA, deck, ["-sf", "cuda"])
        got_n, total, per_slot = _host_mirror_stats(out)
        assert (got_n, total) == (nmirrors, 21), (got_n, total)
        assert 200 < per_slot < 300 + 8 * nmirrors, per_slot
        ref, got = read_dumps(wd_ref), read_dumps(wd_cuda)
        assert sorted(ref) == sorted(got) and len(ref) >= 3
        for s in ref:
            assert ref[s][0] == got[s][0]
            for k, c in enumerate(ref[s][0]):
                sc = max(max(np.abs(ref[q][1][:, k]).max() for q in ref), 1e-300)
                assert np.abs(ref[s][1][:, k] - got[s][1][:, k]).max() <= TOL * sc, (s, c)
        if nmirrors:
            cols = ref[max(ref)][0]
            assert np.abs(ref[max(ref)][1][:, cols.index("c_pa")]).max() > 0
