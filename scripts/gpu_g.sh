#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r1_default.json 2> gpurun_out/bench_r1_default.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_r1_default.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_reference.json 2> gpurun_out/bench_r1_reference.err; echo "ref exit $?"
CMD="python bench.py --workload powerlaw_20m --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gat_(fwd_items|bwd_dst_items|bwd_src_rows)|gemm_tc|dw_tc' -s 15 -c 5 -o gpurun_out/prof_r1b -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"
