#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + stall samples / instructions by CUDA source line."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for i, h in enumerate(hdr):
    if h in want or ("average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        try:
            if float(vals[i].replace(",", "")) < 0.05 and "stalled" in h: continue
        except ValueError: pass
        print(f"{h:90s} {units[i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
data = []; cur = None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 8 and r[0].isdigit() and r[2] == "-":
        try: data.append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()[:105]))
        except ValueError: pass
    elif r and r[0] == "File Path": cur = r[1].split("/")[-1]
tot = sum(d[0] for d in data) or 1; toti = sum(d[1] for d in data) or 1
print("total samples", tot, "warp-instructions", toti)
for s, i, f, ln, text in sorted(data, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% ins {f}:{ln}: {text}")
