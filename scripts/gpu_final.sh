#!/bin/bash
# round-end style check: gpu tests, smoke, default bench, reference arm, launch list + ncu of the dominant kernel
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log | cut -c1-300
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_final_reference.json 2>/dev/null; echo "ref exit $?"
timeout 300 python bench.py --workload elliptic --steps 30 --warmup 5 > gpurun_out/bench_final_elliptic.json 2>/dev/null; echo "ell exit $?"
timeout 300 python bench.py --workload skew --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_final_skew.json 2>/dev/null; echo "skew exit $?"
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/ncu_plain200.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_200m.csv $CMD > gpurun_out/ncu_launch200.log 2>&1
echo "ncu exit $?"
