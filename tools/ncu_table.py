"""Side-by-side table of key metrics for every kernel in an .ncu-rep: python tools/ncu_table.py rep [extra_metric ...]."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__cluster_max_active", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"] + sys.argv[2:]
want += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
ki = hdr.index("Kernel Name")
print("| metric | unit | " + " | ".join(f"`{r[ki][:40]}`" for r in rows[2:]) + " |")
print("|---|---|" + "---:|" * (len(rows) - 2))
for w in want:
    if w in hdr:
        i = hdr.index(w)
        name = w.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")
        print(f"| {name} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")
