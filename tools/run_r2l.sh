#!/bin/bash
# round 2, GPU call L: species / solid pair instantiations after the register work: parity + per-config throughput
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_tile_vs_gather.py tests/test_fused_integrator.py tests/test_random_stress.py -q -p no:cacheprovider --maxfail=12 > $O/r2l_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2l_steps.log
timeout 1500 python tools/config_bench.py --scales 8 --steps 200 --warm 40 --out $O/r2l_config_bench.json > $O/r2l_config_bench.txt 2>&1; echo "config rc=$?" >> $O/r2l_steps.log
cat $O/r2l_steps.log; tail -3 $O/r2l_pytest.log; grep -E "^FAILED|^ERROR" $O/r2l_pytest.log | head; cat $O/r2l_config_bench.txt
