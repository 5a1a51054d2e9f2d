#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gat_gpu.py -q -k "bf16 or tensor_core" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_bf16.log 2>&1; tail -3 gpurun_out/pytest_bf16.log; grep -E "^E  +(Assert|assert)" gpurun_out/pytest_bf16.log | head -5
timeout 300 python bench.py --workload powerlaw_20m --bf16 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_bf16_20m.log 2>&1; grep -o '"ms_per_step": [0-9.]*\|"stages_ms": {[^}]*}' gpurun_out/bench_bf16_20m.log
timeout 600 python bench.py --bf16 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_bf16_200m.json 2>/dev/null; grep -o '"ms_per_step": [0-9.]*\|"stages_ms": {[^}]*}\|"layer": {[^}]*}' gpurun_out/bench_bf16_200m.json
