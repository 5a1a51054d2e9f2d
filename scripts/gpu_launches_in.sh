#!/bin/bash
# launch list (per-kernel durations) of one input-space step on the 20M-edge graph
mkdir -p gpurun_out
CMD="python bench.py --workload powerlaw_20m --algo 3 --steps 1 --warmup 3 --no-e2e --no-cpu"
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -s 150 -c 60 --csv --log-file gpurun_out/launches_in.csv $CMD > gpurun_out/ncu_launch_in.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_in.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:60]),{})[r[mi]]=r[vi]
for (i,k),m in list(d.items())[-34:]:
    print(i,k, m.get('gpu__time_duration.sum'), 'lsu%',m.get('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed'),'issue%',m.get('smsp__issue_active.avg.pct_of_peak_sustained_active'),'warps%',m.get('sm__warps_active.avg.pct_of_peak_sustained_active'),'dram%',m.get('dram__throughput.avg.pct_of_peak_sustained_elapsed'))
PY
