"""Column subset of an `ncu --set full` report as CSV (what profiles/*_ncu_full_summary.csv hold).

    python scripts/summarize_ncu_full.py report.ncu-rep > summary.csv
"""
import csv
import subprocess
import sys

COLS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keep = [c for c in COLS if c in hdr]
w = csv.writer(sys.stdout)
w.writerow(keep)
for r in rows[1:]:
    w.writerow([r[hdr.index(c)].replace("pn2::<unnamed>::", "") for c in keep])
