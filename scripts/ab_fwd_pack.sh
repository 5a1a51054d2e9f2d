#!/bin/bash
# same-box A/B of the packed-rows kernels (GNNFD_FWD_PACK / GNNFD_BWD_PACK = 0/1) after the GPU suite with both on (default)
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4
GNNFD_FWD_PACK=0 GNNFD_BWD_PACK=0 timeout 240 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "not headline" 2>&1 | tail -1
for P in 0 1 0 1; do
echo -n "elliptic PACK=$P "; GNNFD_FWD_PACK=$P GNNFD_BWD_PACK=$P timeout 300 python bench.py --workload elliptic --steps 30 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['stages_ms'])"
done
for P in 0 1 0 1; do
echo -n "200m PACK=$P "; GNNFD_FWD_PACK=$P GNNFD_BWD_PACK=$P timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['layer']['frac'], d['roofline']['stages_ms'])"
done
echo -n "skew PACK=1 "; timeout 300 python bench.py --workload skew --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['layer']['frac'])"
