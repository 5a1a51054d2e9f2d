#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  +(Assert|assert)" gpurun_out/pytest_gpu.log | cut -c1-200 | head -40
timeout 300 python bench.py --workload powerlaw_20m --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c20m.log 2>&1; tail -c 700 gpurun_out/bench_c20m.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c200m.log 2>&1; tail -c 900 gpurun_out/bench_c200m.log
