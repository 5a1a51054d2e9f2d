#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/diag_in.py > gpurun_out/diag_in.log 2>&1; echo "diag exit $?"
grep -E "BAD|==|WORST|Error|error" gpurun_out/diag_in.log | head -30
bash scripts/gpu_bench_in.sh
