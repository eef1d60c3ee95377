"""Summarise an .ncu-rep (read here, no GPU): one line per captured launch with the roofline-relevant metrics.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
for r in rows[2:]:
    print("; ".join(f"{w}={r[idx[w]]}{(' ' + units[idx[w]]) if units[idx[w]] else ''}" for w in want if w in idx))
