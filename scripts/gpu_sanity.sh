#!/bin/bash
# last check of the committed build: all GPU tests, smoke(), both formulations on the 20M-edge graph
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu_sanity.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_sanity.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_sanity.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_sanity.log | cut -c1-300
for a in 0 3; do
timeout 300 python bench.py --workload powerlaw_20m --algo $a --steps 10 --warmup 3 > gpurun_out/sanity_20m_a$a.json 2> gpurun_out/sanity_20m_a$a.err; echo "20m algo $a exit $?"
python -c "
import json; d=json.load(open('gpurun_out/sanity_20m_a$a.json')); r=d['roofline']; print(d['ms_per_step'], {k: round(v,2) for k,v in r['stages_ms'].items()}, 'csr', round(d['config']['csr_build_ms'],1), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1), 'traffic', r['traffic'])"
done
