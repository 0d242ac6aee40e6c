#!/bin/bash
# tools/ab_bench.sh NAME...  -- run bench.py (short, device-resident leg only) once per variant library
# and print the per-family kernel times; results land in gpurun_out/ab_NAME.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  lib=sph-bvf_b200/variants/libsphbvf_$v.so
  [ "$v" = base ] && lib=sph-bvf_b200/libsphbvf.so
  nofuse=0
  [ "$v" = nofuse ] && lib=sph-bvf_b200/libsphbvf.so && nofuse=1
  SPHBVF_NO_FUSE=$nofuse SPHBVF_LIB=$PWD/$lib python bench.py --steps ${STEPS:-30} --warmup 5 --no-cpu --no-e2e ${BENCH_ARGS} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().splitlines()[-1])
    k = d["kernels"]; s = d["steps"]
    print("%-12s step %.3f ms | pair %.3f  init %.3f  final %.3f  fused %.3f  rebuild/step %.3f  pack %.3f" % (
        v, d["ms_per_step"], k["pair"]["ms"] / s, k["initial_integrate"]["ms"] / s, k["final_integrate"]["ms"] / s,
        k.get("final_initial_pack_fused", {"ms": 0})["ms"] / s, k["neighbor_rebuild"]["ms"] / s, k["pack_halo"]["ms"] / s))
except Exception as ex:
    print(v, "FAILED", ex, open("gpurun_out/ab_%s.err" % v).read()[-500:])
PY
done
