#!/bin/bash
# scaling run: N = 8 (or whatever is visible), then 4, 2
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for G in $NG 4 2; do
  if [ "$G" -le "$NG" ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $G --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_scale_${G}.log 2>&1
    echo "G=$G exit $?"; grep -o '"ms_per_step": [0-9.]*\|"stages_ms": {[^}]*}\|"csr_build_ms": [0-9.]*' gpurun_out/bench_scale_${G}.log | tr '\n' ' '; echo
  fi
done
