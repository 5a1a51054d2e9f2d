#!/bin/bash
# per-kernel launch list of one TemporalGNN training step (tgn_snapshots workload, 1 GPU, eager)
mkdir -p gpurun_out
CMD="python bench.py --workload tgn_snapshots --steps 1 --warmup 3 --no-cpu --graph off"
timeout 120 $CMD > gpurun_out/tgn_plain.log 2>&1 || { echo plain failed; tail -3 gpurun_out/tgn_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_tgn.csv $CMD > gpurun_out/ncu_launch_tgn.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_tgn.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
items=[(int(r[ii]), r[ki], float(r[vi].replace(',',''))) for r in rows[1:] if r[ii].isdigit()]
# last step = after the last bce kernel start... take the last quarter of launches (4 steps: 3 warmup + 1)
n=len(items); last=items[3*n//4:]
agg=collections.OrderedDict()
for _,k,t in last:
    nm=k.split('(')[0].replace('void ','')[:70]
    a=agg.setdefault(nm,[0,0.0]); a[0]+=1; a[1]+=t/1e3
tot=sum(a[1] for a in agg.values())
print(f"{len(last)} launches in the last step, {tot/1e3:.3f} ms of kernel time")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:45]:
    print(f"{a[1]:9.1f} us {100*a[1]/tot:5.1f}% x{a[0]:3d}  {k}")
PY
