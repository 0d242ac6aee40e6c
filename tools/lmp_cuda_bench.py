"""The synthetic 3D cavity deck of bench.py as an unmodified LAMMPS input through lmp_cuda -sf cuda (one LAMMPS
process, SPHBVF_NGPU GPUs): what a user of the reference sees end to end, including LAMMPS' own setup.
  python tools/lmp_cuda_bench.py N NGPU [steps]"""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
n, ngpu = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
exe = os.path.join(ROOT, "sph-bvf_b200", "lammps", "_build", "lmp_cuda")
wd = tempfile.mkdtemp(prefix="lmpcuda_")
open(os.path.join(wd, "in.lmp"), "w").write(bench.REF_DECK.format(n=n, steps=steps, warm=10))
env = dict(os.environ, SPHBVF_NGPU=str(ngpu), SPHBVF_VERBOSE="1")
t0 = time.time()
out = subprocess.run([exe, "-in", "in.lmp", "-log", "none", "-echo", "none", "-sf", "cuda"], cwd=wd, env=env,
                     capture_output=True, text=True, timeout=3000)
wall = time.time() - t0
loops = re.findall(r"Loop time of ([0-9.eE+-]+) on (\d+) procs for (\d+) steps with (\d+) atoms", out.stdout)
print([l for l in out.stdout.splitlines() if l.startswith("sphbvf")])
if out.returncode or not loops:
    print(out.stdout[-2000:], out.stderr[-1000:])
    sys.exit(1)
t, _, st, atoms = loops[-1]
print("lmp_cuda n=%d ngpu=%d: %s atoms, %s steps in %s s -> %.4g atom-steps/s (whole process wall %.1f s)" % (
    n, ngpu, atoms, st, t, int(atoms) * int(st) / float(t), wall))
