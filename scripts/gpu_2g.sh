#!/bin/bash
# multi-GPU check (run under gpurun --gpus N): NCCL / peer-memory parity test of every partition variant, then the scaling bench
G=${1:-2}
WL=${2:-"powerlaw_20m powerlaw_200m"}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_partition_gpu.py -x -q -s -p no:cacheprovider > gpurun_out/pytest_mgpu_${G}g.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|Error|error" gpurun_out/pytest_mgpu_${G}g.log | tail -5
for W in $WL; do
for X in ${EXCH:-auto}; do
GNNFD_EXCHANGE=$X GNNFD_BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --workload $W --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_scale_${G}_${W}_$X.log 2> gpurun_out/bench_scale_${G}_${W}_$X.err
echo "G=$G $W exchange=$X exit $?"; grep "stages_ms" gpurun_out/bench_scale_${G}_${W}_$X.err | head -8; tail -1 gpurun_out/bench_scale_${G}_${W}_$X.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['exchange'], d['timing'], d['e2e'])"
done; done
