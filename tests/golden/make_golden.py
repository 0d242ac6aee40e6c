#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Test infrastructure.  Needs /root/reference and `make -C oracle ref` (oracle/_ref/ref_harness);
it therefore only runs in the build container, never on the GPU box -- the fixtures it writes
(*.npz) are committed.

For every case the reference's own example deck (examples/ssa-tsdpd/...) is read where it lies,
edited IN MEMORY (dump vtk line dropped, resolution variable reduced so the fixture stays small,
run length shortened, `e` forced to 0 where the deck sets a non-zero value because the
reference's random stress term is seeded from clock(), snapshot fix + parameter echoes appended)
and written to a scratch directory, then executed by ref_harness.  Two extra cases are decks of
our own (synthetic 3D cavity lattice of SURVEY.md 8(d), and a 3D periodic box with a free elastic
solid) to cover 3D and paths no shipped deck reaches.

Fixture (.npz) content:
  meta           JSON: dim, periodic, box, dt, skin, every/delay/check, variant, S, ntypes,
                 types[{mass,rho0,c0,G0}], pairs[{i,j,eta,h,cutc,kappa}], fixes[...], integrate_groupbit,
                 steps, nbuild_steps
  init_<field>   state handed to the implementation under test (tag order)
  s<step>_<f>    reference fields after <step> steps
  p<step>        canonical (min tag, max tag) pair list used in step <step>
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from refsnap import canonical_pairs, read_snapshot  # noqa: E402

REF = os.environ.get("SPHBVF_REFERENCE", "/root/reference")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
EX = os.path.join(REF, "examples", "ssa-tsdpd")

ECHO_CMDS = ("pair_coeff", "pair_style", "mass", "fix", "neighbor", "neigh_modify", "timestep",
             "atom_style", "dimension", "boundary")

VARIANTS = {"ssa_tsdpd/bvf/transportVelocity": 0, "ssa_tsdpd/bvf/mechanics": 1, "ssa_tsdpd/bvf/fsi": 2}

SYNTH3D = """
# synthetic 3D lid-driven cavity lattice (SURVEY.md 8d): n^3 simple-cubic particles, 3-layer walls
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
{perturb}
fix integration all ssa_tsdpd/bvf/transportVelocity
variable skin equal 0.01*v_h
neighbor ${{skin}} bin
variable dt equal 0.05*v_h/10.0
timestep ${{dt}}
thermo 1000
run 0
"""

PERTURB = """
displace_atoms fluid random $(0.1*v_delta) $(0.1*v_delta) $(0.1*v_delta) 20261018 units box
variable ux atom 0.1*sin(PI*x)*cos(PI*y)
variable uy atom -0.1*cos(PI*x)*sin(PI*y)
velocity fluid set v_ux v_uy 0.0 units box
variable r atom 1.0+0.01*sin(2*PI*x)*sin(2*PI*y)*sin(2*PI*z)
set group fluid ssa_tsdpd/rho v_r
"""

SOLID3D = """
# 3D fully periodic box: fluid with an elastic FREE solid slab and species; exercises ghosts in 3D,
# Jaumann rate, artificial stress, deviatoric force, species flux.  Deck of our own.
dimension 3
units si
atom_style ssa_tsdpd/atomic 1 0 0
boundary p p p
variable n equal {n}
variable delta equal 1.0/v_n
region domain block 0 1 0 1 0 1 units box
create_box 2 domain
lattice sc ${{delta}} origin 0.5 0.5 0.5
create_atoms 1 box
region slab block 0.3 0.7 0.3 0.7 0.3 0.7 units box
group slab region slab
set group slab type 2
group fluid subtract all slab
variable m equal v_delta*v_delta*v_delta
mass 1 ${{m}}
mass 2 $(1.2*v_m)
set group fluid ssa_tsdpd/rho 1.0
set group slab ssa_tsdpd/rho 1.2
set group all ssa_tsdpd/e 0.
set group slab ssa_tsdpd/solid_tag 1 free
variable h equal 2.6*v_delta
pair_style {pair}
pair_coeff 1 1 1.0 10.0 1e-2 ${{h}} ${{h}} 0.0 1e-3
pair_coeff 1 2 1.0 10.0 1e-2 ${{h}} ${{h}} 0.0 1e-3
pair_coeff 2 2 1.2 20.0 1e-2 ${{h}} ${{h}} 50.0 2e-3
displace_atoms all random $(0.05*v_delta) $(0.05*v_delta) $(0.05*v_delta) 4711 units box
variable ux atom 3.0*sin(2*PI*x)*cos(2*PI*y)
variable uy atom -3.0*cos(2*PI*x)*sin(2*PI*y)
variable uz atom 1.0*sin(2*PI*z)
velocity all set v_ux v_uy v_uz units box
variable cc atom 0.5+0.5*sin(2*PI*x)*sin(2*PI*y)
set group all ssa_tsdpd/C 0 v_cc
fix integration all {fix}
variable skin equal 0.05*v_h
neighbor ${{skin}} bin
neigh_modify delay 2 every 1 check yes
timestep 2e-4
thermo 1000
# a first `run 0` makes vest = v and rhoI = rho on the owned atoms, so the ghosts created by the
# setup of the measured run are consistent with their owners (avoids reference quirk SURVEY D.9)
run 0
run 0
"""


def edit_reference_deck(text, subs, nsteps, force_e0):
    """Return the edited deck text (reference deck stays untouched on disk)."""
    out = []
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("dump ") or s.startswith("dump\t") or s.startswith("dump_modify"):
            continue
        for pat, rep in subs:
            line = re.sub(pat, rep, line)
        out.append(line)
    text = "\n".join(out) + "\n"
    # drop the run command; it is re-added by finish_deck
    text = re.sub(r"(?m)^\s*run\s+.*$", "", text)
    if force_e0:
        text += "set group all ssa_tsdpd/e 0.\n"
    text += "run 0\n"
    return text


def finish_deck(text, nsteps, snap_every):
    """Add parameter echoes after the interesting commands, the snapshot fix and the run."""
    lines = []
    for line in text.splitlines():
        lines.append(line)
        s = line.split("#", 1)[0].strip()
        if not s:
            continue
        cmd = s.split()[0]
        if cmd in ECHO_CMDS:
            lines.append('print """SPHBVF_ECHO %s"""' % s)
    text = "\n".join(lines) + "\n"
    run = "fix zz_snap all sphbvf/snapshot %d snap pairs\nrun %d\n" % (snap_every, nsteps)
    # replace the LAST "run 0"
    idx = text.rfind("run 0")
    return text[:idx] + run + text[idx + len("run 0"):]


def expand_range(tok, n):
    """LAMMPS force->bounds() for a type wildcard token."""
    if tok == "*":
        return 1, n
    if "*" in tok:
        a, b = tok.split("*")
        return (int(a) if a else 1), (int(b) if b else n)
    return int(tok), int(tok)


def parse_run(workdir, S, ntypes):
    log = open(os.path.join(workdir, "out.txt")).read()
    echoes = [ln[len("SPHBVF_ECHO "):].split() for ln in log.splitlines() if ln.startswith("SPHBVF_ECHO ")]
    meta = {"types": [dict(mass=0.0, rho0=0.0, c0=0.0, G0=0.0) for _ in range(ntypes)], "pairs": {},
            "fixes": [], "every": 1, "delay": 10, "check": 1}
    fixbits = {}
    for ln in open(os.path.join(workdir, "snap.meta.txt")):
        w = ln.split()
        if w[0] == "neighbor":
            meta["skin"] = float(w[1])
            meta["every"], meta["delay"], meta["check"] = int(w[2]), int(w[3]), int(w[4])
        elif w[0] == "pair_style":
            meta["pair_style"] = w[1]
            meta["variant"] = VARIANTS[w[1]]
        elif w[0] == "fix":
            fixbits[w[1]] = (w[2], int(w[3]))
    for w in echoes:
        if w[0] == "pair_coeff":
            ilo, ihi = expand_range(w[1], ntypes)
            jlo, jhi = expand_range(w[2], ntypes)
            rho0, c0, eta, h, cutc, G0 = [float(v) for v in w[3:9]]
            kappa = [float(v) for v in w[9:9 + S]]
            for i in range(ilo, ihi + 1):      # pair_...transport_velocity.cpp:999-1019
                meta["types"][i - 1].update(rho0=rho0, c0=c0, G0=G0)
                for j in range(max(jlo, i), jhi + 1):
                    meta["pairs"]["%d %d" % (i, j)] = dict(i=i, j=j, eta=eta, h=h, cutc=cutc, kappa=kappa)
        elif w[0] == "fix":
            fid, style = w[1], w[3]
            bit = fixbits[fid][1]
            a = w[4:]
            if style.startswith("ssa_tsdpd/bvf/"):
                meta["integrate_groupbit"] = bit
                meta["fix_style"] = style
            elif style == "ssa_tsdpd/buoyancy":
                meta["fixes"].append(dict(kind="buoyancy", groupbit=bit, gravity=int(a[0] == "gravity"),
                                          accel=float(a[1]), coord=int(a[2]), k=int(a[3]), Cref=float(a[4])))
            elif style == "ssa_tsdpd/forcing":
                kind = {"tsdpd": 0, "velocity": 1}[a[0]]
                shape = {"circle": 0, "rectangle": 1}[a[3]]
                if shape == 0:
                    cx, cy, p, q, val = float(a[4]), float(a[5]), float(a[6]), 0.0, float(a[7])
                else:
                    cx, cy, p, q, val = [float(v) for v in a[4:9]]
                meta["fixes"].append(dict(kind="forcing", groupbit=bit, what=kind, step=int(a[1]), idx=int(a[2]),
                                          shape=shape, cx=cx, cy=cy, a=p, b=q, value=val))
            elif style == "ssa_tsdpd/buffer":
                kind = {"tsdpd": 0, "velocity": 1, "density": 2}[a[0]]
                axis = {"x": 0, "y": 1}[a[1]]
                meta["fixes"].append(dict(kind="buffer", groupbit=bit, what=kind, axis=axis, step=int(a[2]),
                                          idx=int(a[3]), cx=float(a[4]), cy=float(a[5]), length=float(a[6]),
                                          width=float(a[7]), value=float(a[8])))
            elif style == "setforce":
                meta["fixes"].append(dict(kind="setforce", groupbit=bit, fx=float(a[0]), fy=float(a[1]), fz=float(a[2])))
            elif style == "ssa_tsdpd/chem_rxn_mass_action":
                nr = int(a[1])
                reactants = [int(v) for v in a[2:2 + nr]]
                npd = int(a[2 + nr])
                products = [int(v) for v in a[3 + nr:3 + nr + npd]]
                meta["fixes"].append(dict(kind="chem_rxn", groupbit=bit, k=float(a[0]), reactants=reactants, products=products))
            elif style == "sphbvf/snapshot":
                pass
            else:
                raise RuntimeError("unhandled fix style " + style)
    meta["pairs"] = list(meta["pairs"].values())
    return meta


TV_FIELDS = ["x", "v", "vest", "f", "rho", "rhoI", "drho", "phi", "number_density", "nw", "ddv",
             "rhoAux1", "rhoAux2"]
SOLID_FIELDS = ["dev", "ddev"]
MECH_FIELDS = ["ddx", "Pnew"]
SPECIES_FIELDS = ["C", "Q"]


def make_case(name, deck_text, nsteps, keep_steps, pair_steps, consistent_ghosts=False):
    with tempfile.TemporaryDirectory() as wd:
        with open(os.path.join(wd, "deck.lmp"), "w") as fh:
            fh.write(finish_deck(deck_text, nsteps, 1))
        with open(os.path.join(wd, "out.txt"), "w") as out:
            subprocess.run([HARNESS, "-in", "deck.lmp", "-log", "none"], cwd=wd, stdout=out,
                           stderr=subprocess.STDOUT, check=True)
        s0 = read_snapshot(os.path.join(wd, "snap.0.bin"))
        meta = parse_run(wd, s0["S"], s0["ntypes"])
        for t in range(s0["ntypes"]):
            meta["types"][t]["mass"] = s0["mass"][t]
        meta.update(name=name, dim=s0["dim"], periodic=s0["periodic"], boxlo=s0["boxlo"], boxhi=s0["boxhi"],
                    dt=s0["dt"], S=s0["S"], ntypes=s0["ntypes"], natoms=s0["natoms"], nsteps=nsteps,
                    steps=sorted(keep_steps), pair_steps=sorted(pair_steps),
                    consistent_ghosts=bool(consistent_ghosts))
        arrays = {}
        f0 = s0["fields"]
        for k in ("tag", "type", "mask", "solid_tag", "fixed_tag", "x", "v", "rho", "e", "C", "dev"):
            arrays["init_" + k] = f0[k]
        fields = list(TV_FIELDS)
        if meta["variant"] != 0:
            fields += MECH_FIELDS
        if (f0["solid_tag"] * (1 - f0["fixed_tag"])).any():
            fields += SOLID_FIELDS
        if s0["S"]:
            fields += SPECIES_FIELDS
        meta["fields"] = fields
        build_steps = []
        for step in range(nsteps + 1):
            snap = read_snapshot(os.path.join(wd, "snap.%d.bin" % step))
            if snap["ago"] == 0 and step > 0:
                build_steps.append(step)
            if step in keep_steps:
                for k in fields:
                    arrays["s%d_%s" % (step, k)] = snap["fields"][k]
            if step in pair_steps:
                arrays["p%d" % step] = canonical_pairs(snap["pairs"]).astype(np.int32)
        meta["build_steps"] = build_steps
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print("%-22s %6d atoms  S=%d variant=%d builds@%s  -> %s (%.1f kB)" % (
            name, s0["natoms"], s0["S"], meta["variant"], build_steps, os.path.basename(path),
            os.path.getsize(path) / 1e3))


REACT2D = """
# reaction-diffusion in a wall-bounded 2D box: three species, two mass-action reactions and a zeroth-order source
# (fix ssa_tsdpd/chem_rxn_mass_action, SURVEY.md 8f-3); fixed timestep so that the plain-C oracle restates it
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
neighbor $(0.3*v_h) bin
timestep 2e-4
run 0
"""


MIXEDH2D = """
# 2D wall-bounded box with TYPE-PAIR dependent smoothing lengths, viscosities and masses: exercises the per-pair
# neighbour cutoffs (neighbor.cpp:278-310) and the non-uniform coefficient path of the pair styles
dimension 2
units si
atom_style ssa_tsdpd/atomic 0 0 0
boundary f f p
variable n equal 30
variable d equal 1.0/(v_n-6)
variable lo equal -3*v_d
variable hi equal 1.0+3*v_d
region box block ${lo} ${hi} ${lo} ${hi} 0 ${d} units box
create_box 3 box
lattice sq ${d} origin 0.5 0.5 0.0
create_atoms 3 box
region inner block 0 1 0 1 0 ${d} units box
group fluid region inner
set group fluid type 1
region blob sphere 0.5 0.5 0.0 0.25 units box
group heavy region blob
set group heavy type 2
group solid subtract all fluid
mass 1 $(v_d*v_d)
mass 2 $(1.5*v_d*v_d)
mass 3 $(v_d*v_d)
set group all ssa_tsdpd/rho 1.0
set group heavy ssa_tsdpd/rho 1.5
set group all ssa_tsdpd/e 0.
set group solid ssa_tsdpd/solid_tag 1 fixed
pair_style ssa_tsdpd/bvf/transportVelocity
pair_coeff 1 1 1.0 10.0 1e-2 $(2.5*v_d) $(2.5*v_d) 0.0
pair_coeff 1 2 1.0 10.0 2e-2 $(2.8*v_d) $(2.8*v_d) 0.0
pair_coeff 1 3 1.0 10.0 1e-2 $(2.5*v_d) $(2.5*v_d) 0.0
pair_coeff 2 2 1.5 8.0 3e-2 $(3.1*v_d) $(3.1*v_d) 0.0
pair_coeff 2 3 1.5 8.0 2e-2 $(2.2*v_d) $(2.2*v_d) 0.0
pair_coeff 3 3 1.0 10.0 1e-2 $(2.5*v_d) $(2.5*v_d) 0.0
variable ux atom 0.8*sin(PI*x)*cos(PI*y)
variable uy atom -0.8*cos(PI*x)*sin(PI*y)
velocity fluid set v_ux v_uy 0.0 units box
fix integ all ssa_tsdpd/bvf/transportVelocity
neighbor $(0.02*v_d) bin
timestep 1e-4
run 0
"""


def ref_deck(rel):
    return open(os.path.join(EX, rel)).read()


def main():
    only = set(sys.argv[1:])

    def want(n):
        return not only or n in only

    if want("cavity_n50"):
        t = edit_reference_deck(ref_deck("lid_driven_cavity/Re100_N50/lid_driven_cavity.lmp"), [], 0, False)
        make_case("cavity_n50", t, 21, {0, 1, 2, 11, 20, 21}, {0, 11, 21})
    if want("cavity_n20"):
        t = edit_reference_deck(ref_deck("lid_driven_cavity/Re100_N50/lid_driven_cavity.lmp"),
                                [(r"variable\s+nx equal 50", "variable nx equal 20")], 0, False)
        make_case("cavity_n20", t, 45, {0, 1, 2, 10, 20, 40, 45}, {0, 10, 40})
    if want("cavity_mech_n20"):
        # the shipped cavity deck with its commented-out alternative pair style enabled
        # (lid_driven_cavity.lmp:148) and the matching integrator: mechanics variant, no ghosts
        t = edit_reference_deck(ref_deck("lid_driven_cavity/Re100_N50/lid_driven_cavity.lmp"),
                                [(r"variable\s+nx equal 50", "variable nx equal 20"),
                                 (r"^pair_style\s+ssa_tsdpd/bvf/transportVelocity", "pair_style ssa_tsdpd/bvf/mechanics"),
                                 (r"ssa_tsdpd/bvf/transportVelocity\s*$", "ssa_tsdpd/bvf/mechanics")], 0, False)
        make_case("cavity_mech_n20", t, 21, {0, 1, 2, 20, 21}, {0, 21})
    if want("natconv_n40"):
        t = edit_reference_deck(ref_deck("natural_convection/Ra_10E4/natural_convection.lmp"),
                                [(r"variable\s+Nxint equal 200", "variable Nxint equal 40"),
                                 (r"variable\s+Nyint equal 200", "variable Nyint equal 40")], 0, True)
        make_case("natconv_n40", t, 25, {0, 1, 2, 20, 25}, {0, 25})
    if want("fsi_nx20"):
        t = edit_reference_deck(ref_deck("fsi/fsi.lmp"), [(r"variable\s+nx equal 60", "variable nx equal 20")], 0, False)
        make_case("fsi_nx20", t, 22, {0, 1, 2, 3, 20, 22}, {0, 22})
    if want("yeast_nx40"):
        t = edit_reference_deck(ref_deck("cell_polarization/case_1/cell_polarization.lmp"),
                                [(r"variable\s+nx equal 100", "variable nx equal 40")], 0, False)
        make_case("yeast_nx40", t, 12, {0, 1, 2, 3, 12}, {0, 12})
    if want("synth3d_n14"):
        make_case("synth3d_n14", SYNTH3D.format(n=14, perturb=PERTURB), 12, {0, 1, 2, 12}, {0, 12})
    if want("synth3d_n14_lattice"):
        make_case("synth3d_n14_lattice", SYNTH3D.format(n=14, perturb=""), 2, {0, 1, 2}, {0})
    if want("mixedh2d_n30"):
        make_case("mixedh2d_n30", MIXEDH2D, 24, {0, 1, 2, 3, 12, 24}, {0, 12, 24})
    if want("react2d_n26"):
        make_case("react2d_n26", REACT2D, 24, {0, 1, 2, 3, 12, 24}, {0, 24})
    for var, pair, fix in (("tv", "ssa_tsdpd/bvf/transportVelocity", "ssa_tsdpd/bvf/transportVelocity"),
                           ("mech", "ssa_tsdpd/bvf/mechanics", "ssa_tsdpd/bvf/mechanics"),
                           ("fsi", "ssa_tsdpd/bvf/fsi", "ssa_tsdpd/bvf/fsi")):
        nm = "solid3d_%s_n10" % var
        if want(nm):
            make_case(nm, SOLID3D.format(n=10, pair=pair, fix=fix), 24, {0, 1, 2, 3, 12, 24}, {0, 12, 24},
                      consistent_ghosts=True)


if __name__ == "__main__":
    main()
