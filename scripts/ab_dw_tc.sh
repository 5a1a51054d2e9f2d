#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "tc or TC or project or gemm or dw" 2>&1 | tail -15
for M in 1 2 1 2; do
echo -n "DW=$M "; GNNFD_DW_TC=$M timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['layer']['frac'], d['roofline']['stages_ms'])"
done
