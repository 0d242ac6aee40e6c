"""One-off diagnostic (variant library built with -DTB_DIAG_NOSTORE): time the list build with and without
the scattered emission stores.  SPHBVF_LIB must point at the variant."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = bench.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
meta = bench.cavity_meta(n)
atoms = bench.cavity_atoms(meta, meta["boxlo"], meta["boxhi"])
eng = pkg.Engine(meta)
eng.set_atoms(atoms["tag"], atoms["type"], atoms["mask"], atoms["solid"], atoms["fixed"], atoms["x"], atoms["v"], atoms["rho"], atoms["e"])
eng.set_run_length(10 ** 9)
eng.setup()
for mode in ("0", "1", "0", "1"):
    os.environ["SPHBVF_TB_NOSTORE"] = mode
    eng.profiling(True)
    for _ in range(3):
        eng.build_neighbors()
    eng.sync()
    ms, cnt = eng.kernel_ms(3)
    eng.profiling(False)
    print("nostore=%s: rebuild %.3f ms each" % (mode, ms / 3))
eng.close()
