#!/usr/bin/env python
"""Text summary of an ncu report of one of the staged kernels: the raw-page metrics the DESIGN cites and the warp-stall
samples of the source page grouped by phase of a tile (phases split at BAR.SYNC).

    python scripts/ncu_summary.py gpurun_out/s31/prof_rigid.ncu-rep "header line" > profiles/....txt
"""
import csv
import io
import subprocess
import sys

rep, header = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def page(name, *extra):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


rows = page("raw")
hdr, units, data = rows[0], rows[1], rows[2:]
print(header)
print("(per-launch numbers under ncu are cold-cache and serialised; compare shares, not absolutes)\n")
kn = hdr.index("Kernel Name")
print("kernel:", data[0][kn][:160], "\n")
for i, h in enumerate(hdr):
    stalled = h.startswith("smsp__average_warps_issue_stalled")
    if h in WANT or stalled:
        vals = [r[i] for r in data]
        try:
            if stalled and all(float(v.replace(",", "")) < 0.15 for v in vals):
                continue
        except ValueError:
            pass
        print(f"  {h:85s} {units[i]:16s} {' '.join(vals)}")

src = page("source", "--print-source", "sass")
blocks, cur = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(r)
b = blocks[0]
h2 = b[0]
d2 = [r for r in b[1:] if len(r) == len(h2)]
ix = {h: i for i, h in enumerate(h2)}
KEYS = ["stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_mio", "stall_math", "stall_not_selected", "stall_selected",
        "stall_branch_resolving", "stall_no_inst", "stall_lg", "stall_dispatch"]
phase, ph = 0, {}
for r in d2:
    n = int(r[ix["# Samples"]])
    e = ph.setdefault(phase, [0, 0, {}])
    e[0] += n
    e[1] += int(r[ix["Instructions Executed"]])
    for k in KEYS:
        e[2][k[6:]] = e[2].get(k[6:], 0) + int(r[ix[k]])
    if "BAR.SYNC" in r[ix["Source"]]:
        phase += 1
total = sum(v[0] for v in ph.values())
print(f"\nWarp-stall samples of launch 0 by phase of a tile (SASS split at BAR.SYNC; {total} samples, {len(d2)} instructions):")
for k, v in ph.items():
    if v[1] == 0:
        continue
    print(f"  phase {k}: samples {v[0]:5d} ({100.0 * v[0] / max(1, total):4.1f} %)  warp instructions {v[1]:8d}  " +
          str({a: c for a, c in v[2].items() if c > 8}))
top = sorted(range(len(d2)), key=lambda k: -int(d2[k][ix["# Samples"]]))[:12]
print("\nInstructions with the most samples:")
for k in sorted(top):
    r = d2[k]
    st = {key[6:]: int(r[ix[key]]) for key in KEYS if int(r[ix[key]]) > 0}
    print(f"  #{k:5d} samples {r[ix['# Samples']]:>4s} executed {r[ix['Instructions Executed']]:>7s}  {r[ix['Source']].strip()[:64]:64s} {st}")
