"""Instructions executed / stall samples per CUDA source line from an ncu report:
   python scripts/ncu_lines.py report.ncu-rep kernel_regex [launch_skip] [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{rx}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
out, fname, hdr = [], None, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0]:
        ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            n, s = int(float(r[ie] or 0)), int(float(r[sm] or 0))
        except ValueError:
            continue
        if n > 0 or s > 0:
            out.append((n, s, fname, r[0], r[1].strip()[:100]))
tot, ts = sum(o[0] for o in out), max(sum(o[1] for o in out), 1)
print("total instructions", tot, "samples", ts)
for n, s, f, l, src in sorted(out, reverse=True)[:top]:
    print(f"{n:>11} {100 * n / tot:5.1f}%  smp {100 * s / ts:5.1f}%  {f}:{l}  {src}")
