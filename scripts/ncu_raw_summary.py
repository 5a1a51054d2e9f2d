"""Key counters per kernel from `ncu -i report --page raw --csv`:  python scripts/ncu_raw_summary.py raw.csv [more.csv ...]"""
import csv, sys
WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__occupancy_limit_registers", "CTAs/SM (regs)"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem)"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe %peak"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instr"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor (hmma) active %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("smsp__inst_executed.sum", "warp instructions")]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    print(f"### {path.split('/')[-1]}\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        print(f"**`{name}`**  grid {r[hdr.index('Grid Size')]}, block {r[hdr.index('Block Size')]}\n")
        cells = []
        for key, label in WANT:
            if key in hdr and r[hdr.index(key)] not in ("", "n/a"):
                v, u = r[hdr.index(key)], units[hdr.index(key)]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:,.3f}".rstrip("0").rstrip(".") if abs(f) < 1e6 else f"{f:,.0f}"
                except ValueError:
                    pass
                cells.append(f"{label}: {v} {u}".strip())
        stalls = {h.split("issue_stalled_")[1].split("_per")[0]: float(r[hdr.index(h)]) for h in hdr
                  if "issue_stalled" in h and "per_issue_active" in h and r[hdr.index(h)] not in ("", "n/a")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:5]
        print("; ".join(cells) + "\n")
        print("top stalls (cycles per issued instruction): " + ", ".join(f"{k} {v:.2f}" for k, v in top) + "\n")
