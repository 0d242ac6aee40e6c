#!/usr/bin/env python
"""Write the reference's SHIPPED example decks, at shipped size, as runnable test inputs.

Test infrastructure.  Needs /root/reference, so it runs in the build container only (called by
__graft_entry__.build()); the GPU box gets the result the same way it gets oracle/_ref/lmp_serial:
oracle/_ref/ is git-ignored (no reference text enters the history) but travels with the snapshot.

Each deck of examples/ssa-tsdpd is read where it lies and edited IN MEMORY, keeping every physics
line as shipped:
  * `dump ... vtk ...`  ->  `dump ... custom ... id type x y vx vy fx fy <the deck's own computes>`
    with `dump_modify sort id format float %.17g` (USER-VTK needs the VTK library, absent here);
  * `run ${nt}`         ->  `run NSTEPS` (bounded horizon; output every NSTEPS/..., see DECKS);
  * the two output-frequency variables are set to the test's cadence;
  * for the `_e0` variant of the natural-convection deck `set group all ssa_tsdpd/e 0.` is
    appended before the run: upstream seeds its random stress from clock(), so only e = 0 is
    reproducible to 1e-10 (the as-shipped e is compared at the noise level of that term).
tests/test_shipped_decks.py runs every deck through oracle/_ref/lmp_serial (unmodified reference) and through
lmp_cuda -sf cuda and compares the dumps.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SPHBVF_REFERENCE", "/root/reference")
EX = os.path.join(REF, "examples", "ssa-tsdpd")
OUT = os.path.join(ROOT, "oracle", "_ref", "decks")

# name -> (path under examples/ssa-tsdpd, steps, dump every, force e = 0)
DECKS = {
    "cavity_Re1000_N200": ("lid_driven_cavity/Re1000_N200/lid_driven_cavity.lmp", 24, 12, False),
    "natconv_Ra1e4": ("natural_convection/Ra_10E4/natural_convection.lmp", 24, 12, False),
    "natconv_Ra1e4_e0": ("natural_convection/Ra_10E4/natural_convection.lmp", 24, 12, True),
    "fsi": ("fsi/fsi.lmp", 24, 12, False),
    "cell_polarization_case1": ("cell_polarization/case_1/cell_polarization.lmp", 24, 12, False),
}


def edit(text, nsteps, every, force_e0):
    out = []
    for line in text.splitlines():
        s = line.split("#", 1)[0].strip()
        w = s.split()
        if len(w) >= 4 and w[0] == "dump" and w[3] == "vtk":
            # dump ID group vtk N file id type vx vy vz c_...  ->  custom, full precision, positions and forces added
            cols = [c for c in w[6:] if c not in ("id", "type")]
            out.append("dump %s %s custom %d dump.*.txt id type x y %s fx fy" % (w[1], w[2], every, " ".join(cols)))
            out.append("dump_modify %s sort id format float %%.17g" % w[1])
            continue
        if len(w) >= 2 and w[0] == "run":
            if force_e0:
                out.append("set group all ssa_tsdpd/e 0.")
            out.append("run %d" % nsteps)
            continue
        if len(w) >= 4 and w[0] == "variable" and w[1] in ("freq_results", "freq_screen") and w[2] == "equal":
            out.append("variable %s equal %d" % (w[1], every))
            continue
        out.append(line)
    return "\n".join(out) + "\n"


def main():
    if not os.path.isdir(EX):
        print("make_decks: %s not present, nothing written" % EX)
        return 0
    os.makedirs(OUT, exist_ok=True)
    for name, (rel, nsteps, every, e0) in DECKS.items():
        text = open(os.path.join(EX, rel)).read()
        new = edit(text, nsteps, every, e0)
        assert re.search(r"(?m)^dump \S+ \S+ custom ", new) and re.search(r"(?m)^run %d$" % nsteps, new), name
        with open(os.path.join(OUT, name + ".lmp"), "w") as fh:
            fh.write(new)
        print("make_decks: %s (%d lines)" % (name, new.count("\n")))
    return 0


if __name__ == "__main__":
    sys.exit(main())
