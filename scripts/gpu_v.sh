#!/bin/bash
mkdir -p gpurun_out
for D in 0 15 4 0; do
echo -n "DBG=$D "; GNNFD_GEMM_DBG=$D timeout 300 python bench.py --workload powerlaw_20m --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | grep -E "stages_ms" | python -c "
import json,sys
for l in sys.stdin:
    print(json.loads(l)['roofline']['stages_ms']['project_fwd'])"
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stages_ms'])"
