#!/bin/bash
mkdir -p gpurun_out
for a in 3 0; do
timeout 300 python bench.py --workload powerlaw_20m --algo $a --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_in_20m_a$a.json 2> gpurun_out/bench_in_20m_a$a.err; echo "20m algo $a exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_in_20m_a$a.json')); r=d['roofline']; print(d['ms_per_step'], {k: round(v,2) for k,v in r['stages_ms'].items()}, {k: round(v) for k,v in r['stages_gbs'].items()}, r['layer']['frac'], r['layer_own_model']['frac'])"
done
timeout 500 python bench.py --algo 3 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_in_200m_a3.json 2> gpurun_out/bench_in_200m_a3.err; echo "200m algo 3 exit $?"; tail -3 gpurun_out/bench_in_200m_a3.err
python -c "
import json; d=json.load(open('gpurun_out/bench_in_200m_a3.json')); r=d['roofline']; print(d['ms_per_step'], {k: round(v,2) for k,v in r['stages_ms'].items()}, {k: round(v) for k,v in r['stages_gbs'].items()}, r['layer']['frac'], r['layer_own_model']['frac'], d['config']['workload'])"
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
