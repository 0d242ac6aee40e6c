#!/bin/bash
# round 2, GPU call D (2 GPUs): multi-rank parity tests + bench with overlapped vs serial halo
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/r2d_gpus.txt
timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_lammps_dropin.py -m gpu -q -p no:cacheprovider -k "bricks or two_gpus" -s > $O/r2d_pytest_2gpu.log 2>&1; echo "pytest rc=$?" > $O/r2d_steps.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 > $O/r2d_bench_g2.json 2> $O/r2d_bench_g2.err; echo "bench g2 rc=$?" >> $O/r2d_steps.log
SPHBVF_HALO=serial timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 10 --no-e2e --no-parity > $O/r2d_bench_g2_serial.json 2> $O/r2d_bench_g2_serial.err; echo "bench g2 serial rc=$?" >> $O/r2d_steps.log
SPHBVF_PAIR=tile timeout 900 $TR --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e > $O/r2d_bench_g2_tile.json 2> $O/r2d_bench_g2_tile.err; echo "bench g2 tile rc=$?" >> $O/r2d_steps.log
cat $O/r2d_steps.log; tail -3 $O/r2d_pytest_2gpu.log
