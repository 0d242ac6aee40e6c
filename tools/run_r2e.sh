#!/bin/bash
# round 2, GPU call E (8 GPUs): multi-rank parity tests (2/4/8 bricks) + bench at N=8, overlapped vs serial halo
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/r2e_gpus.txt
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q -p no:cacheprovider -s > $O/r2e_pytest_8gpu.log 2>&1; echo "pytest rc=$?" > $O/r2e_steps.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 10 > $O/r2e_bench_g8.json 2> $O/r2e_bench_g8.err; echo "bench g8 rc=$?" >> $O/r2e_steps.log
SPHBVF_HALO=serial timeout 900 $TR --master-port 29522 bench.py --gpus 8 --steps 50 --warmup 10 --no-e2e --no-parity > $O/r2e_bench_g8_serial.json 2> $O/r2e_bench_g8_serial.err; echo "bench g8 serial rc=$?" >> $O/r2e_steps.log
cat $O/r2e_steps.log; tail -3 $O/r2e_pytest_8gpu.log
