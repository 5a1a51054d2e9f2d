#!/bin/bash
# quick check of the input-space path: its parity tests, then the 20M-edge bench with both algorithms
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_input_space_gpu.py -q -x -p no:cacheprovider > gpurun_out/pytest_in_quick.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_in_quick.log
for a in ${ALGOS:-3}; do
timeout 300 python bench.py --workload powerlaw_20m --algo $a --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_in_20m_a$a.json 2> gpurun_out/bench_in_20m_a$a.err; echo "20m algo $a exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_in_20m_a$a.json')); r=d['roofline']; print(d['ms_per_step'], {k: round(v,2) for k,v in r['stages_ms'].items()}, d['clocks'])"
done
