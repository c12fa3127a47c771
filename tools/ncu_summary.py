"""Key metrics of every kernel in an .ncu-rep (no GPU needed): python tools/ncu_summary.py x.ncu-rep [--json out.json] [--points N]
--json writes one record per launch {kernel, time_us, dram_bytes, dram_pct, tensor_pct, l1tex_pct, issue_pct, regs, grid, block, points}:
the file bench.py reads `roofline.traffic` / `roofline.ncu` from (profiles/r02*_ncu.json)."""
import csv, json, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread ", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
def _f(r, key):
    try:
        return float(r[hdr.index(key)].replace(",", ""))
    except Exception:
        return None
def _bytes(r, key):
    v = _f(r, key)
    if v is None:
        return None
    u = units[hdr.index(key)].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
def _us(r):
    v = _f(r, "gpu__time_duration.sum")
    u = units[hdr.index("gpu__time_duration.sum")].lower()
    return None if v is None else v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
if "--json" in sys.argv:
    pts = int(sys.argv[sys.argv.index("--points") + 1]) if "--points" in sys.argv else None
    recs = []
    for r in rows[2:]:
        rd, wr = _bytes(r, "dram__bytes_read.sum"), _bytes(r, "dram__bytes_write.sum")
        recs.append({"kernel": r[hdr.index("Kernel Name")], "time_us": _us(r), "dram_bytes": (rd or 0) + (wr or 0) if rd is not None else None,
                     "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "dram_pct": _f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                     "tensor_pct": _f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                     "l1tex_pct": _f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                     "issue_pct": _f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "regs": _f(r, "launch__registers_per_thread"), "grid": r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else None,
                     "block": r[hdr.index("launch__block_size")] if "launch__block_size" in hdr else None, "points": pts})
    json.dump(recs, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("## " + name[:110])
    for h, u, v in zip(hdr, units, r):
        if any((h + " ").startswith(k) or h == k.strip() for k in KEYS):
            print("  %-90s %12s %s" % (h, v, u))
