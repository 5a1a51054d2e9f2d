"""Opcode histogram (executed count, shared wavefronts vs ideal, stall samples) from `ncu --page source --csv`."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return int(float(r[ix[k]]))
    except Exception: return 0
seen = set(); ops = collections.Counter(); wf = collections.Counter(); idl = collections.Counter(); smp = collections.Counter()
for r in rows[2:]:
    if len(r) <= ix['# Samples'] or not r[0].startswith('0x') or r[0] in seen: continue
    seen.add(r[0])
    s = r[ix['Source']].split()
    if not s: continue
    op = s[1] if s[0].startswith('@') and len(s) > 1 else s[0]
    ops[op] += num(r, 'Instructions Executed'); wf[op] += num(r, 'L1 Wavefronts Shared')
    idl[op] += num(r, 'L1 Wavefronts Shared Ideal'); smp[op] += num(r, '# Samples')
tot = sum(ops.values()); ts = max(sum(smp.values()), 1)
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(op.ljust(28), n, '%.1f%%' % (100 * n / tot), 'wf', wf[op], 'ideal', idl[op], 'samples %.1f%%' % (100 * smp[op] / ts))
print('total', tot, 'samples', ts)
