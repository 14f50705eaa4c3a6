#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the small CSV kept under profiles/:
    python scripts/ncu_summarise.py gpurun_out/prof_fused.ncu-rep profiles/r01_s2_fused_step_ncu_full.csv
One column per captured launch, one row per metric of interest (B200_PROFILING.md's list)."""
import csv
import io
import subprocess
import sys

WANT = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "sm__cycles_elapsed.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    cols = {name: i for i, name in enumerate(header)}
    names = [n for n in WANT if n in cols] + sorted(n for n in header if n.startswith(STALLS) and not n.endswith("_not_issued"))
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit"] + [f"launch{k}" for k in range(len(data))])
        w.writerow(["Kernel Name", ""] + [r[cols["Kernel Name"]] for r in data])
        for n in names:
            w.writerow([n, units[cols[n]]] + [r[cols[n]] for r in data])
    print("wrote", out, "for", len(data), "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
