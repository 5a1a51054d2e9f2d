#!/bin/bash
# same-box A/B of the packed-rows forward kernel (GNNFD_FWD_PACK=0/1) + the GPU suite with packing forced on
mkdir -p gpurun_out
GNNFD_FWD_PACK=1 timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "not headline" 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "not headline" 2>&1 | tail -1
for P in 0 1 0 1; do
echo -n "elliptic PACK=$P "; GNNFD_FWD_PACK=$P timeout 300 python bench.py --workload elliptic --steps 30 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stages_ms'])"
done
for P in 0 1; do
echo -n "200m PACK=$P "; GNNFD_FWD_PACK=$P timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stages_ms'])"
done
