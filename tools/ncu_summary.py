"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py report.ncu-rep [pattern ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
                        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "launch__registers_per_thread",
                        "launch__grid_size", "launch__block_size", "launch__cluster", "lts__t_sector_hit_rate.pct",
                        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct",
                        "smsp__warp_issue_stalled", "smsp__average_warp", "sm__throughput.avg.pct", "launch__shared_mem",
                        "launch__occupancy_limit", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__cycles_active.avg"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for v in rows[2:]:
    print("==", v[hdr.index("Kernel Name")][:110])
    for i, h in enumerate(hdr):
        if any(p in h for p in pats):
            val = v[i]
            if val in ("", "0", "n/a"):
                continue
            print(f"  {h} = {val} {units[i]}")
