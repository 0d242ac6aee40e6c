#!/bin/bash
# round 2, GPU call Y: the integrator kernels changed signature for the early halo: single-GPU regression of the final library
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_fused_integrator.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_lammps_dropin.py -q -p no:cacheprovider --maxfail=10 > $O/r2y_pytest.log 2>&1; echo "pytest rc=$?" > $O/r2y_steps.log
timeout 300 python bench.py --no-cpu --no-lammps --steps 50 > $O/r2y_bench_g1.json 2> $O/r2y_bench_g1.err; echo "bench rc=$?" >> $O/r2y_steps.log
cat $O/r2y_steps.log; tail -2 $O/r2y_pytest.log; python -c "
import json; d=json.loads(open('gpurun_out/r2y_bench_g1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['parity_check']['ok'], d['e2e']['value'])"
