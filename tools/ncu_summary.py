"""Turn .ncu-rep captures into the compact CSV kept under profiles/ (run where ncu is installed):
    python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/ncu_full_rN.csv"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
]
STALLS = ["long_scoreboard", "barrier", "membar", "wait", "short_scoreboard", "not_selected", "selected",
          "math_pipe_throttle", "lg_throttle", "branch_resolving", "no_instructions", "dispatch_stall", "sleeping",
          "mio_throttle"]
METRICS += [f"smsp__pcsamp_warps_issue_stalled_{s}" for s in STALLS]

out = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    keep = [i for i, h in enumerate(hdr) if h == "Kernel Name" or "__" in h]
    for j, r in enumerate(rows):
        if j == 0 and not first:
            continue  # one header; the units row is repeated per report (ncu picks us / ms per report)
        o = [r[i] for i in keep]
        if j >= 2:
            o[0] = o[0].split("(")[0].split("::")[-1][-48:]
        out.writerow(o)
    first = False
