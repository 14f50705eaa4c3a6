#!/usr/bin/env python
"""In-process A/B timing of kernel variants (experiments; not part of the product or of bench.py's line).

Box-to-box and minute-to-minute differences on the shared B200 pool are as large as the effects being
measured (the same binary: 10.8 vs 11.9 us per step in two sessions), so variants are built side by side in
ONE process -- the environment knobs (CONSTV_FUSED_EARLY, CONSTV_FUSED_ZERO, CONSTV_FUSED_BLOCK,
CONSTV_SETTLE_WS) are read when a slab is created -- and timed in alternation, several rounds; the table
printed at the end is the median over rounds.

    python scripts/ab_variants.py fused      # the unconstrained fused step, 1 and 2 chains
    python scripts/ab_variants.py settle     # the rigid-water step: TMA pipeline vs LDG-staged kernel
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
which = sys.argv[1] if len(sys.argv) > 1 else "fused"
atoms = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
PRECISION = sys.argv[3] if len(sys.argv) > 3 else "single"
sys.argv = ["bench.py", "--atoms", str(atoms), "--precision", PRECISION]
import bench  # noqa: E402

args = bench.parse_args()
args.replicas = 3 if atoms > 1_000_000 else 0
env = bench.Env()
peak, _ = bench.peaks()
R = atoms // 2
K, W, ROUNDS = 1500, 30, 4

if which == "chains":
    variants = [("default", {})]
    shard = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED)
    kind, water = "langevin", None
    chain_counts = (2, 3, 4, 6, 8)
elif which == "fused":
    variants = [("early0_zero0", {"CONSTV_FUSED_EARLY": "0", "CONSTV_FUSED_ZERO": "0"}),
                ("early1_zero0", {"CONSTV_FUSED_EARLY": "1", "CONSTV_FUSED_ZERO": "0"}),
                ("early0_zero2", {"CONSTV_FUSED_EARLY": "0", "CONSTV_FUSED_ZERO": "2"}),
                ("early1_zero2", {"CONSTV_FUSED_EARLY": "1", "CONSTV_FUSED_ZERO": "2"}),
                ("early1_zero1", {"CONSTV_FUSED_EARLY": "1", "CONSTV_FUSED_ZERO": "1"})]
    shard = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED)
    kind, water = "langevin", None
    chain_counts = (1, 2)
elif which == "reordered":
    # the per-slot image tables (constv_image_by_slot_kernel) against the gathers through atomIndex -> invAtomOrder / imageCharge
    variants = [("gathers", {"CONSTV_IMAGE_TABLES": "0"}), ("per_slot_image_tables", {"CONSTV_IMAGE_TABLES": "1"})]
    shard = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED)
    perm = env.slabmod.molecule_permutation(shard, shard.meta["n_ion"], 3, shard.meta["n_water"] // 3, seed=7)
    shard, _ = env.slabmod.permute_slab(shard, perm)
    kind, water = "langevin", None
    chain_counts = (1,)
elif which in ("settle_chains", "drude_chains", "settle_reordered", "drude_reordered"):
    variants = [("default", {})]
    drude = None
    if which.startswith("drude"):
        shard, d_normal, d_pairs = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED, maker=env.slabmod.make_drude_slab)
        kind, water, drude = "drude", None, (d_normal, d_pairs)
    else:
        shard, w_atoms, w_params, w_cons = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED, maker=env.slabmod.make_water_slab)
        kind, water = "settle", (w_atoms, w_params, w_cons)
    chain_counts = (2, 3, 4, 6, 8, 12)
    if os.environ.get("AB_CHAINS"):
        chain_counts = tuple(int(x) for x in os.environ["AB_CHAINS"].split(","))
    if which.endswith("reordered"):
        chain_counts = (1,)
    if which.endswith("reordered"):
        per = 4 if which.startswith("drude") else 3
        n_mol = shard.meta["n_pairs"] if which.startswith("drude") else shard.meta["n_water"] // 3
        perm = env.slabmod.molecule_permutation(shard, shard.meta["n_ion"], per, n_mol, seed=7)
        shard, _ = env.slabmod.permute_slab(shard, perm)
    ROUNDS = 3
elif which in ("settle_bulk", "settle_bulk_reordered"):
    # the staged kernel (cp.async stage-in, variant 7) against the same shape with a bulk stage-in by one thread (variant 6);
    # variant 0 = the library's rule (bulk for a chain's launch and for launches of at least two waves of CTAs)
    variants = [("cp_async_staged", {"_variant": 7}), ("bulk_stage_in", {"_variant": 6}), ("automatic", {"_variant": 0})]
    shard, w_atoms, w_params, w_cons = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED, maker=env.slabmod.make_water_slab)
    kind, water = "settle", (w_atoms, w_params, w_cons)
    chain_counts = (1, 2, 3, 6)
    if which.endswith("reordered"):
        perm = env.slabmod.molecule_permutation(shard, shard.meta["n_ion"], 3, shard.meta["n_water"] // 3, seed=7)
        shard, _ = env.slabmod.permute_slab(shard, perm)
        chain_counts = (1,)
    ROUNDS = 3
elif which == "settle_bulk_early":
    # bulk stage-in: the three Philox blocks of a molecule drawn while the bulk copies are in flight (CONSTV_SETTLE_EARLY=1)
    variants = [("bulk_draw_late", {"_variant": 6, "CONSTV_SETTLE_EARLY": "0"}), ("bulk_draw_early", {"_variant": 6, "CONSTV_SETTLE_EARLY": "1"}),
                ("cp_async_staged", {"_variant": 7, "CONSTV_SETTLE_EARLY": "0"})]
    shard, w_atoms, w_params, w_cons = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED, maker=env.slabmod.make_water_slab)
    kind, water = "settle", (w_atoms, w_params, w_cons)
    chain_counts = (1, 3, 6)
    ROUNDS = 3
elif which == "drude":
    # the staged Drude step: thread t <-> work item t (spt1) against the balanced kernel (csrc/constv_drude_kernels.cuh)
    variants = [("item_per_thread", {"CONSTV_DRUDE_SPT": "1"}), ("balanced_2_slots_per_thread", {"CONSTV_DRUDE_SPT": "2"}),
                ("balanced_3_slots_per_thread", {"CONSTV_DRUDE_SPT": "3"})]
    shard, d_normal, d_pairs = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED, maker=env.slabmod.make_drude_slab)
    kind, water, drude = "drude", None, (d_normal, d_pairs)
    chain_counts = (1, 2, 3)
    ROUNDS = 3
else:
    # "_variant" = constv_set_kernel_variant: 0 the default (cp.async stage-in), 3 the stage filled through registers (round 1)
    variants = [("ldg_staged", {"_variant": 3}), ("cp_async_stage_in", {"_variant": 0})]
    shard, w_atoms, w_params, w_cons = env.slabmod.slab_tile(R, 1, 0, PRECISION, seed=bench.SLAB_SEED, maker=env.slabmod.make_water_slab)
    kind, water = "settle", (w_atoms, w_params, w_cons)
    chain_counts = (1, 2, 3)
    ROUNDS = 3

wls = []
for name, knobs in variants:
    os.environ.pop("CONSTV_SETTLE_WS", None)
    os.environ.update({k: v for k, v in knobs.items() if not k.startswith("_")})
    wls.append((name, bench.Workload(env, args, shard, kind=kind, chains=1, water=water, drude=drude if which.startswith("drude") else None,
                                     reorder_perm=perm if which.endswith("reordered") else None, variant=knobs.get("_variant", 0))))
alg = wls[0][1].algorithmic_bytes()
results = {}
for rnd in range(ROUNDS):
    for name, wl in wls:
        for ch in chain_counts:
            wl.set_chains(ch)
            t = bench.time_steps(env, wl, K, W, repeats=7)
            results.setdefault((name, wl.chains), []).append(1e3 * t["ms"] / K)
for (name, ch), us in sorted(results.items()):
    med = float(np.median(us))
    print(json.dumps({"variant": name, "chains": ch, "us_per_step": round(med, 3), "rounds": [round(x, 3) for x in us],
                      "GBps": round(alg / med * 1e-3, 1), "frac_measured_peak": round(alg / med * 1e-3 / peak, 4),
                      "frac_8TBs": round(alg / med * 1e-3 / 8000.0, 4), "atoms": atoms}), flush=True)
