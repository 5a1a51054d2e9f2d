#!/bin/bash
# first GPU pass: parity tests, smoke, small + full benches (no profiler)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --workload elliptic --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_elliptic.log 2>&1
echo "exit $?" >> gpurun_out/bench_elliptic.log
timeout 600 python bench.py --workload powerlaw_20m --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_pl20m.log 2>&1
echo "exit $?" >> gpurun_out/bench_pl20m.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_pl200m.log 2>&1
echo "exit $?" >> gpurun_out/bench_pl200m.log
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench_elliptic.log; tail -2 gpurun_out/bench_pl20m.log; tail -2 gpurun_out/bench_pl200m.log
