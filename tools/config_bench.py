"""Per-config throughput of BASELINE.json configs 1-4: the reference's SHIPPED example decks (oracle/_ref/decks, written
by tests/golden/make_decks.py from /root/reference/examples/ssa-tsdpd) through `lmp_cuda -sf cuda`, at shipped size and
with the deck's own resolution variable scaled up so that the GPU is loaded.

  python tools/config_bench.py [--scales 1,8] [--steps 200] [--warm 40] [--decks cavity,natconv,fsi,cellpol]
                               [--exe lmp_cuda|lmp_serial] [--out gpurun_out/config_bench.json]

Edits made in memory (physics lines untouched): `variable <nx> equal N` -> N * scale, timestep -> dt / scale (same
Courant number at the finer spacing), dump lines removed, thermo every 10^6 steps, `run` -> `run WARM` + `run STEPS`
(the second LAMMPS "Loop time" is what is reported).  Every deck runs twice: once plain (the throughput) and once
with SPHBVF_PROFILE=1 (CUDA events around every kernel family; the pair kernel's share and launches per step).
--exe lmp_serial times the unmodified reference on one host core on the same edited deck (keep --scales 1).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECKDIR = os.path.join(ROOT, "oracle", "_ref", "decks")
# short name -> (deck file, resolution variables, pair kernel instantiation it exercises)
DECKS = {
    "cavity": ("cavity_Re1000_N200.lmp", ["nx"], "pair_kernel<TV, S=0, SOLIDS=1>  (BASELINE config 1)"),
    "natconv": ("natconv_Ra1e4_e0.lmp", ["Nxint", "Nyint"], "pair_kernel<TV, S=1, SOLIDS=1> + buoyancy/forcing fixes  (config 2)"),
    "fsi": ("fsi.lmp", ["nx"], "pair_kernel<MECH, S=0, SOLIDS=2> + buffer fixes  (config 3)"),
    "cellpol": ("cell_polarization_case1.lmp", ["nx"], "pair_kernel<FSI, S=1, SOLIDS=2>, full-list style + forcing  (config 4)"),
}


def edit(text, resvars, scale, warm, steps):
    out = []
    for line in text.splitlines():
        s = line.split("#", 1)[0].strip()
        w = s.split()
        if len(w) >= 4 and w[0] == "variable" and w[1] in resvars and w[2] == "equal":
            out.append("variable %s equal %d" % (w[1], int(round(float(w[3]))) * scale))
            continue
        if w and w[0] in ("dump", "dump_modify"):
            continue
        if len(w) >= 2 and w[0] == "thermo":
            out.append("thermo 1000000")
            continue
        if len(w) >= 2 and w[0] == "timestep":
            out.append("timestep $(%s/%d.0)" % (w[1].replace("${", "v_").replace("}", ""), scale) if scale != 1 else line)
            continue
        if len(w) >= 2 and w[0] == "run":
            out.append("run %d" % warm)
            out.append("run %d" % steps)
            continue
        out.append(line)
    return "\n".join(out) + "\n"


def run_one(exe, deck, env_extra, timeout):
    wd = tempfile.mkdtemp(prefix="cfgbench_")
    open(os.path.join(wd, "in.lmp"), "w").write(deck)
    cmd = [exe, "-in", "in.lmp", "-log", "none", "-echo", "none"]
    if exe.endswith("lmp_cuda"):
        cmd += ["-sf", "cuda"]
    t0 = time.time()
    p = subprocess.run(cmd, cwd=wd, env=dict(os.environ, **env_extra), capture_output=True, text=True, timeout=timeout)
    wall = time.time() - t0
    loops = re.findall(r"Loop time of ([0-9.eE+-]+) on (\d+) procs for (\d+) steps with (\d+) atoms", p.stdout)
    prof = [l for l in p.stdout.splitlines() if l.startswith("sphbvf profile")]
    if p.returncode or len(loops) < 2:
        return {"error": (p.stdout[-1500:] + p.stderr[-500:]), "wall_s": wall}
    t, _, st, atoms = loops[-1]
    r = {"atoms": int(atoms), "steps": int(st), "loop_s": float(t), "us_per_step": float(t) / int(st) * 1e6,
         "atom_steps_per_s": int(atoms) * int(st) / float(t), "wall_s": wall}
    if prof:
        fam = {}
        for name, ms, nl in re.findall(r"(\w+) ([0-9.]+) / (\d+);", prof[-1]):
            fam[name] = {"ms": float(ms), "launches": int(nl)}
        r["families"] = fam
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scales", default="1,8")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warm", type=int, default=40)
    ap.add_argument("--decks", default="cavity,natconv,fsi,cellpol")
    ap.add_argument("--exe", default="lmp_cuda")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config_bench.json"))
    ap.add_argument("--timeout", type=int, default=1500)
    a = ap.parse_args()
    exe = os.path.join(ROOT, "sph-bvf_b200", "lammps", "_build", "lmp_cuda") if a.exe == "lmp_cuda" else os.path.join(ROOT, "oracle", "_ref", "lmp_serial")
    rows = []
    for name in a.decks.split(","):
        fn, resvars, what = DECKS[name]
        text = open(os.path.join(DECKDIR, fn)).read()
        for scale in [int(s) for s in a.scales.split(",")]:
            deck = edit(text, resvars, scale, a.warm, a.steps)
            plain = run_one(exe, deck, {}, a.timeout)
            row = {"deck": name, "file": fn, "scale": scale, "kernel": what, "exe": a.exe, **plain}
            if "error" not in plain and a.exe == "lmp_cuda":
                prof = run_one(exe, deck, {"SPHBVF_PROFILE": "1"}, a.timeout)
                if "families" in prof:
                    fam = prof["families"]
                    tot = sum(v["ms"] for v in fam.values())
                    row["profiled"] = {"us_per_step": prof["us_per_step"], "families_ms": {k: v["ms"] for k, v in fam.items()},
                                       "launches": {k: v["launches"] for k, v in fam.items()},
                                       "pair_ms_per_step": fam["pair"]["ms"] / max(fam["pair"]["launches"], 1),
                                       "pair_share_of_gpu_time": fam["pair"]["ms"] / tot if tot > 0 else None,
                                       "note": "families cover setup + warm-up + timed run of the profiled pass"}
            rows.append(row)
            if "error" in row:
                print("%-8s x%-2d ERROR %s" % (name, scale, row["error"][-400:]), flush=True)
            else:
                pp = row.get("profiled", {})
                print("%-8s x%-2d %9d atoms %9.1f us/step %10.4g atom-steps/s   pair %.3f ms/launch (%.0f %% of GPU time)" % (
                    name, scale, row["atoms"], row["us_per_step"], row["atom_steps_per_s"], pp.get("pair_ms_per_step", float("nan")),
                    100 * (pp.get("pair_share_of_gpu_time") or 0)), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, "w"), indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
