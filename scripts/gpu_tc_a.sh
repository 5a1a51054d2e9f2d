#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gat_gpu.py -q -k "tensor_core" --timeout 200 -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -30 gpurun_out/pytest_tc.log
timeout 300 python bench.py --workload powerlaw_20m --steps 5 --warmup 3 --no-cpu --no-e2e --algo 2 > gpurun_out/bench_tc20m.log 2>&1; tail -c 1500 gpurun_out/bench_tc20m.log
