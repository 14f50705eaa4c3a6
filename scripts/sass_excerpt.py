#!/usr/bin/env python
"""Writes profiles/r02_sass_excerpts.txt: for the hot kernels of libconstv_b200.so (sm_100a cubin) the register
count from the build log and the SASS lines that show HOW memory is moved -- vector widths and cache operators of
LDG / STG, asynchronous copies (LDGSTS = cp.async, UBLKCP = cp.async.bulk), mbarrier traffic (SYNCS), MUFU of the Box-Muller,
programmatic dependent launch (ACQBULK / PREEXIT).  Needs cuobjdump (CUDA toolkit); no GPU."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "openmm_constv_b200", "libconstv_b200.so")
KERNELS = [
    ("constv_fused_step_kernel<single, 256 threads, 1 pair/thread, evict-first, early> (the headline kernel)", "constv_fused_step_kernelILi0ELi256ELi1ELb1ELb1E"),
    ("constv_fused_step_indexed_kernel<single, 256> (reordered slots)", "constv_fused_step_indexed_kernelILi0ELi256E"),
    ("constv_fused_step_tma_kernel<single, 256, 4 stages> (TMA ring variant of the fused step)", "constv_fused_step_tma_kernelILi0ELi256ELi4ELb1E"),
    ("constv_settle_step_ws_kernel<single, 4 consumer groups> (rigid-water step, warp-specialised TMA pipeline)", "constv_settle_step_ws_kernelILi0ELi4E"),
    ("constv_settle_step_tiled_kernel<single, 128, identity, evict-first, cp.async stage-in> (rigid-water step, the default)", "constv_settle_step_tiled_kernelILi0ELi128ELb0ELb1ELb1E"),
    ("constv_settle_step_tiled_kernel<single, 128, identity, evict-first, stage-in through registers> (round 1's, kernel variant 3)", "constv_settle_step_tiled_kernelILi0ELi128ELb0ELb1ELb0E"),
    ("constv_drude_step_balanced_kernel<single, 128 threads, 2 slots per thread, identity, evict-first> (Drude step, the default)", "constv_drude_step_balanced_kernelILi0ELi128ELi2ELb0ELb1E"),
    ("constv_drude_step_tiled_kernel<single, 128, identity, evict-first> (round 1's, kernel variant 5)", "constv_drude_step_tiled_kernelILi0ELi128ELb0ELb1E"),
]
PAT = re.compile(r"\b(LDGSTS|LDGDEPBAR|DEPBAR|LDG|STG|LDS|STS|UBLKCP|UTMALDG|UTMASTG|SYNCS|MUFU|ACQBULK|PREEXIT|BAR|ATOMG|ATOMS|RED|LDL|STL|ERRBAR|MEMBAR|FENCE)\b[.\w]*")


def main():
    names = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout  # noqa: F841 (keeps cuobjdump honest)
    out = ["SASS excerpts of openmm_constv_b200/libconstv_b200.so (cuobjdump -sass; arch sm_100a), round 2.",
           "Per kernel: mnemonic histogram of the memory / synchronisation instructions, then their first occurrences in program order.", ""]
    log = open(os.path.join(ROOT, "openmm_constv_b200", "build_ptxas.log")).read()
    listing = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", listing)
    out.append("arch line of the cubin: " + (re.search(r"arch = (\S+)", listing).group(1) if re.search(r"arch = (\S+)", listing) else "?"))
    out.append("")
    for title, key in KERNELS:
        blk = next((b for b in blocks if b.split("\n", 1)[0].find(key) >= 0), None)
        out.append("== " + title)
        if blk is None:
            out.append("   (not found in this build)\n")
            continue
        mangled = blk.split("\n", 1)[0].strip()
        m = re.search(re.escape(mangled) + r"'.*?\n.*?\n(.*?bytes spill loads)\n.*?Used (\d+) registers", log, re.S)
        if m:
            out.append(f"   {m.group(2)} registers; {m.group(1).strip()}")
        hist = collections.Counter()
        firsts = []
        n_instr = 0
        for line in blk.splitlines():
            mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if not mm:
                continue
            n_instr += 1
            ins = mm.group(2).strip()
            hit = PAT.search(ins)
            if hit:
                op = hit.group(0)
                hist[op] += 1
                if len(firsts) < 60 and (hist[op] <= 3):
                    firsts.append(f"   /*{mm.group(1)}*/ {ins}")
        out.append(f"   {n_instr} SASS instructions")
        out.append("   " + ", ".join(f"{k} x{v}" for k, v in sorted(hist.items())))
        out.extend(firsts)
        out.append("")
    path = os.path.join(ROOT, "profiles", "r02_sass_excerpts.txt")
    open(path, "w").write("\n".join(out) + "\n")
    print(path, len(out), "lines")


if __name__ == "__main__":
    main()
