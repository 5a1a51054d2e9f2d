#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_model.py > gpurun_out/bench_models.json 2> gpurun_out/bench_models.err; echo "exit $?"; cat gpurun_out/bench_models.json; tail -3 gpurun_out/bench_models.err
timeout 300 python bench.py --workload elliptic --steps 30 --warmup 5 --no-cpu --no-e2e 2>/dev/null | grep -o '"ms_per_step": [0-9.]*\|"stages_ms": {[^}]*}'
