"""Summarise an ncu launch list (gpu__time_duration + dram bytes) for the last step of a bench run."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if 'Kernel Name' in r:
        hdr, start = r, i
        break
ki, mi, vi, idi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
data = collections.OrderedDict()
for r in rows[start + 2:]:
    if len(r) <= vi:
        continue
    d = data.setdefault(r[idi], {'name': r[ki]})
    d[r[mi]] = float(r[vi].replace(',', ''))
items = list(data.values())
idx = [i for i, d in enumerate(items) if 'build_b_images' in d['name']]
last = items[idx[-1]:]
agg = collections.OrderedDict()
for d in last:
    nm = d['name'].split('(')[0][:64]
    a = agg.setdefault(nm, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get('gpu__time_duration.sum', 0) / 1e6
    a[2] += d.get('dram__bytes_read.sum', 0) / 1e9; a[3] += d.get('dram__bytes_write.sum', 0) / 1e9
tot = sum(a[1] for a in agg.values())
print(f"last step: {len(last)} launches, {tot:.2f} ms (ncu: cold cache, serialised -- compare shares, not absolutes)")
print(f"{'ms':>9} {'share':>6} {'n':>3} {'rd GB':>8} {'wr GB':>8} {'GB/s':>7}  kernel")
for k, a in agg.items():
    print(f"{a[1]:9.3f} {100 * a[1] / tot:5.1f}% {a[0]:3d} {a[2]:8.2f} {a[3]:8.2f} {(a[2] + a[3]) / max(a[1], 1e-9) * 1e3:7.0f}  {k}")
