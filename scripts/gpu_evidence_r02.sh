#!/bin/bash
# round-2 evidence set on the shipped build: gpu tests, default bench, launch list and `ncu --set full` captures of the
# two packed-rows kernels + the src-major pass + the forward GEMM on the headline workload.  Every ncu pass runs only
# after the same command exited 0 on its own; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
T=${1:-r02a}
timeout 420 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_$T.log
timeout 400 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench exit $?"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
timeout 200 $CMD > gpurun_out/ncu_plain_$T.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_launch_$T.log 2>&1
echo "ncu launch list exit $?"
# step 4 of 4 (3 warm-ups): skip the first three launches of each kernel, capture the fourth
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'gat_fwd_items_pack|gat_bwd_dst_items_pack|gat_bwd_src_rows|gemm_tc_ws2|dw_tc2' -s 15 -c 5 \
    -o gpurun_out/prof_$T -f $CMD > gpurun_out/ncu_full_$T.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/ncu_full_$T.log
