#!/bin/bash
# scaling point at G GPUs (run under gpurun --gpus G): [parity test on the first <=4 GPUs] + headline workload with the e2e leg and
# the in-bench self-check + the time-step sharded TGN step
G=${1:-4}
mkdir -p gpurun_out
if [ "${2:-test}" = "test" ]; then
timeout 400 python -m pytest tests/test_partition_gpu.py -x -q -s -p no:cacheprovider > gpurun_out/pytest_mgpu_${G}g.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/pytest_mgpu_${G}g.log | tail -2
fi
GNNFD_BENCH_DEBUG=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_scale_${G}.log 2> gpurun_out/bench_scale_${G}.err
echo "G=$G exit $?"; grep "stages_ms" gpurun_out/bench_scale_${G}.err | head -8; tail -1 gpurun_out/bench_scale_${G}.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['timing'], d['e2e'], d['config']['csr_build_ms'], d['selfcheck'])"
bash scripts/gpu_tgn.sh $G
