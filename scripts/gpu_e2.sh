#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 600 python -m pytest tests/test_partition_gpu.py -q --timeout 500 -p no:cacheprovider > gpurun_out/pytest_mgpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_mgpu.log; tail -15 gpurun_out/pytest_mgpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload powerlaw_20m --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_2gpu_20m.log 2>&1
echo "exit $?" >> gpurun_out/bench_2gpu_20m.log; tail -c 1500 gpurun_out/bench_2gpu_20m.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_200m.log 2>&1
echo "exit $?" >> gpurun_out/bench_2gpu_200m.log; tail -c 1500 gpurun_out/bench_2gpu_200m.log
