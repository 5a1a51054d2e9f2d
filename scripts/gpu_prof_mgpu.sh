#!/bin/bash
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 scripts/prof_mgpu.py > gpurun_out/prof_mgpu_$NG.log 2>&1; echo "exit $?"; grep -E " ms" gpurun_out/prof_mgpu_$NG.log | tail -12
