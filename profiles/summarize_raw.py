"""Key figures of `ncu --page raw --csv` exports (one kernel launch each) as a markdown table.
usage: summarize_raw.py <tag> <raw.csv> [<raw.csv> ...]   (writes to stdout)"""
import csv, sys
KEYS = [("gpu__time_duration.sum", "time us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem KB"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("smsp__inst_executed.sum", "warp instr"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %")]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "no_instruction", "math_pipe_throttle", "mio_throttle", "lg_throttle", "not_selected", "branch_resolving", "membar"]
def load(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
    hdr, units, vals = rows[h[0]], rows[h[0] + 1], rows[h[0] + 2]
    return {k: (v, u) for k, u, v in zip(hdr, units, vals)}
print("| kernel | " + " | ".join(n for _, n in KEYS) + " | top stalls (warps per issue) |")
print("|---|" + "---|" * (len(KEYS) + 1))
for p in sys.argv[2:]:
    d = load(p)
    name = d.get("Kernel Name", ("?", ""))[0].split("(")[0]
    cells = []
    for k, _ in KEYS:
        v, u = d.get(k, ("", ""))
        try:
            f = float(v.replace(",", ""))
            v = ("%.4g" % f) if abs(f) < 1e6 else "%.4g" % f
        except ValueError:
            pass
        cells.append(("%s %s" % (v, u)).strip())
    st = []
    for s in STALLS:
        v = d.get("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s, ("0", ""))[0]
        try:
            st.append((float(v), s))
        except ValueError:
            pass
    st.sort(reverse=True)
    print("| %s | " % name + " | ".join(cells) + " | " + ", ".join("%s %.2f" % (s, v) for v, s in st[:4]) + " |")
