#!/bin/bash
mkdir -p gpurun_out
GNNFD_GEMM_WS=2 timeout 240 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "tc or TC or project or gemm" 2>&1 | tail -15
for W in 1 2 1 2; do
echo -n "WS=$W "; GNNFD_GEMM_WS=$W timeout 300 python bench.py --workload powerlaw_20m --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | grep -E "stages_ms" | python -c "
import json,sys
for l in sys.stdin:
    print(json.loads(l)['roofline']['stages_ms']['project_fwd'])"
done
