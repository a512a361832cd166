"""Summarise an .ncu-rep (raw page) into the handful of metrics profiles/ cites.  usage: ncu_summary.py rep [metric-prefix ...]"""
import sys, csv, subprocess
rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
                        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
                        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
                        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled", "gpu__compute_memory_throughput",
                        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts", "smsp__pcsamp_warps_issue_stalled"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = [i for i, h in enumerate(hdr) if any(h.startswith(w) for w in want)]
kn = hdr.index("Kernel Name")
for r in rows[2:]:
    print("#### " + r[kn][:110])
    for i in idx:
        if r[i] not in ("", "0", "n/a"):
            print(f"{hdr[i]} [{units[i]}] = {r[i]}")
