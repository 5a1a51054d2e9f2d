#!/bin/bash
# 8-GPU: exchange mechanisms of the input-space partition on the headline workload, then the time-step sharded TGN step
G=${1:-8}
mkdir -p gpurun_out
for X in ${EXCH:-nccl multicast}; do
GNNFD_EXCHANGE=$X GNNFD_BENCH_DEBUG=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_scale_${G}_200m_$X.log 2> gpurun_out/bench_scale_${G}_200m_$X.err
echo "G=$G exchange=$X exit $?"; grep "stages_ms" gpurun_out/bench_scale_${G}_200m_$X.err | head -3; tail -1 gpurun_out/bench_scale_${G}_200m_$X.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['exchange'][:40], d['timing'])"
done
bash scripts/gpu_tgn.sh $G
