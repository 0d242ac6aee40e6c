#!/bin/bash
# tools/scale_run.sh -- on an 8-GPU box: brick-decomposition parity at 2/4/8 ranks, then the weak-scaling bench at 4 and 8
cd "$(dirname "$0")/.."
python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
for n in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29740+n)) \
    bench.py --gpus $n --steps 50 --warmup 5 --no-e2e > gpurun_out/bench_g${n}_r1g.json 2> gpurun_out/bench_g${n}_r1g.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open("gpurun_out/bench_g%s_r1g.json" % n).read().strip().splitlines()[-1])
    print("N=%s: %.4g atom-steps/s, %.3f ms/step, %d atoms" % (n, d["value"], d["ms_per_step"], d["config"]["atoms"]))
except Exception as ex:
    print("N=%s failed" % n, ex, open("gpurun_out/bench_g%s_r1g.err" % n).read()[-800:])
PY
done
