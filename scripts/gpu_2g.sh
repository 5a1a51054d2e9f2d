#!/bin/bash
# multi-GPU check (run under gpurun --gpus N): NCCL parity test of every partition variant, then the scaling bench
G=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_partition_gpu.py -x -q -p no:cacheprovider > gpurun_out/pytest_mgpu_${G}g.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_mgpu_${G}g.log
for W in powerlaw_20m powerlaw_200m; do
GNNFD_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --workload $W --steps 5 --warmup 3 > gpurun_out/bench_scale_${G}_$W.log 2> gpurun_out/bench_scale_${G}_$W.err
echo "G=$G $W exit $?"; grep "stages_ms" gpurun_out/bench_scale_${G}_$W.err | head -8; tail -1 gpurun_out/bench_scale_${G}_$W.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['stages_ms'], d['e2e'])"
done
