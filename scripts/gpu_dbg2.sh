#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
echo "--- run $i"; timeout 300 python scripts/diag_in.py 2>&1 | grep -E "BAD|WORST|rror" | head -8
done
bash scripts/gpu_bench_in.sh
