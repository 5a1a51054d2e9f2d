#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gat_gpu.py -q -k "tensor_core or fp32_mean" --timeout 300 -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tc.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  +(Assert|assert)|exit" gpurun_out/pytest_tc.log | cut -c1-200 | head -20
for WS in 1 0; do
GNNFD_GEMM_WS=$WS timeout 300 python bench.py --workload powerlaw_20m --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_l20m_$WS.log 2>&1; echo "WS=$WS"; grep -o '"ms_per_step": [0-9.]*\|"stages_ms": {[^}]*}' gpurun_out/bench_l20m_$WS.log
done
