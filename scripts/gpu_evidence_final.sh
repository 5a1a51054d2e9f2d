#!/bin/bash
# Evidence set of the shipped build on ONE B200 (tag = $1): gpu tests, smoke, the default bench line and the reference arm,
# the secondary workloads, then -- only after the same command exited 0 on its own -- the ncu launch list (per-kernel time
# and DRAM bytes) and `ncu --set full` captures of the main kernels of both formulations.  Numbers printed under ncu are
# never bench values.
mkdir -p gpurun_out
T=${1:-r02}
timeout 420 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu_$T.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_$T.log | cut -c1-400
timeout 500 python bench.py > gpurun_out/bench_${T}_default.json 2> gpurun_out/bench_${T}_default.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${T}_reference.json 2>/dev/null; echo "ref exit $?"
timeout 300 python bench.py --algo 3 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${T}_inputspace_200m.json 2>/dev/null; echo "input-space 200m exit $?"
timeout 300 python bench.py --workload elliptic --steps 30 --warmup 5 > gpurun_out/bench_${T}_elliptic.json 2>/dev/null; echo "elliptic exit $?"
timeout 300 python bench.py --workload skew --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${T}_skew.json 2>/dev/null; echo "skew exit $?"
timeout 300 python bench.py --workload tgn_snapshots --steps 20 --warmup 3 > gpurun_out/bench_${T}_tgn_1gpu.json 2>/dev/null; echo "tgn exit $?"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
if timeout 200 $CMD > gpurun_out/ncu_plain_$T.log 2>&1; then
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${T}_200m.csv $CMD > gpurun_out/ncu_launch_$T.log 2>&1; echo "ncu launch list exit $?"
# step 4 of 4 (3 warm-ups): skip the first three launches of each kernel, capture the fourth
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'gat_fwd_items_pack|gat_bwd_dst_items_pack|gat_bwd_src_rows|in_proj_gemm|dw_tc2' -s 15 -c 5 \
    -o /tmp/prof_${T}_200m -f $CMD > gpurun_out/ncu_full_$T.log 2>&1; echo "ncu full exit $?"
# the report itself is too big to travel (gpurun_out is capped at 64 MiB): keep its raw page as CSV
ncu -i /tmp/prof_${T}_200m.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${T}_200m.csv 2>/dev/null
fi
CMD2="python bench.py --workload powerlaw_20m --algo 3 --steps 1 --warmup 3 --no-e2e --no-cpu"
if timeout 200 $CMD2 > gpurun_out/ncu_plain_in_$T.log 2>&1; then
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${T}_in_20m.csv $CMD2 > gpurun_out/ncu_launch_in_$T.log 2>&1; echo "ncu launch list (input-space) exit $?"
timeout 500 ncu --set full --clock-control none --import-source on \
    -k regex:'in_alpha_items|gat_in_fwd_items|gat_in_bwd_items|in_out_gemm|in_dw_gemm|in_proj_gemm|in_logits_kernel' -s 27 -c 9 \
    -o /tmp/prof_${T}_in_20m -f $CMD2 > gpurun_out/ncu_full_in_$T.log 2>&1; echo "ncu full (input-space) exit $?"
ncu -i /tmp/prof_${T}_in_20m.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${T}_in_20m.csv 2>/dev/null
fi
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
du -sh gpurun_out
