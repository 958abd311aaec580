"""Print the key counters of every kernel in an .ncu-rep (reads `ncu --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct", "smsp__issue_active.avg.pct",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor_cycles_active.avg.pct", "sm__pipe_fma_cycles_active.avg.pct",
        "l1tex__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts.sum.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__throughput.avg.pct", "gpu__dram_throughput.avg.pct", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__shared_mem_per_block", "launch__occupancy_per",
        "smsp__average_warps_issue_stalled", "smsp__warp_issue_stalled", "sm__cycles_elapsed.max", "launch__waves_per"]
idx = [i for i, h in enumerate(hdr) if any(h.startswith(w) for w in want) and "per_second" not in h and "pct_of_peak_sustained_elapsed" not in h.replace("throughput.avg.pct_of_peak_sustained_elapsed", "")]
for r in rows[2:]:
    print("=" * 100)
    for i in idx:
        v = r[i]
        if v in ("", "0", "n/a"):
            continue
        print(f"{hdr[i][:86]:88s} {units[i]:10s} {v[:60]}")
