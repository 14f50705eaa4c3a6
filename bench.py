#!/usr/bin/env python
"""Benchmark of the ConstVLangevinIntegrator step path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one fused integrator step over one synthetic slab of ``--atoms`` particles per GPU
(default 1,000,000 = 500k real + 500k image: BASELINE.json configs[2]).  Metric: atom-steps/s,
whole job (sum over ranks / max-over-ranks device time).  Multi-GPU is the range partition of
SURVEY.md 8e: every rank owns a contiguous real-atom range plus its images and there is NO
data-path collective (weak scaling: per-GPU shard fixed, so N=8 with --atoms 2000000 is the 16M
slab of configs[3]); the only collective is the scalar kinetic-energy all-reduce, issued once after
the timed region as a report, never per step.

Chains (default 2): the slab is cut into contiguous ranges of real atoms, each with its own device step
counter and its own stream (constv_set_chains / constv_step_fused_chain), so that the drain of one
range's step k overlaps the ramp-up of another range's step k+1; results are bit-identical to one launch
per step (tests/test_gpu_parity.py::test_chains_are_bit_identical_to_the_single_launch).  --chains 1 is
the single launch per step.

L2 policy: a 1M-atom slab is 58 MB, smaller than B200's 126 MB L2, so the timed loop rotates over
``replicas`` independent copies of the slab whose total footprint exceeds 2x L2: every step streams
its slab from HBM.

The JSON line carries, besides the base contract's keys:
  roofline      algorithmic bytes of one launch / average launch duration (CUDA events around the
                timed region on the launching stream / launches), against MEASURED_PEAKS.json
  cpu_baseline  the oracle (kind "port": the reference has no CPU implementation of this path)
                timed on this box's host cores on a bounded sample of the same workload
  e2e           the same metric through the C ABI with HOST buffers (constv_step_host): per step
                H2D of the real atoms' forces from pinned memory, the fused step, D2H of posq
  clocks        SM clock / throttle reasons sampled during the timed region
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_KELVIN, FRICTION, DT = 300.0, 1.0, 0.001  # example/nacl_1m_spce/nacl_constv.py:12-16
DRUDE_FRICTION_CM, DRUDE_T, DRUDE_FRICTION, DRUDE_WALL = 5.0, 1.0, 20.0, 0.02  # usual Drude-oscillator settings
L2_BYTES = 126 * 1024 * 1024
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100_000)
    ap.add_argument("--warmup", type=int, default=1000)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--atoms", type=int, default=1_000_000, help="particles (real+image) per GPU")
    ap.add_argument("--precision", default="single", choices=["single", "mixed", "double"])
    ap.add_argument("--ensemble", type=int, default=0,
                    help="BASELINE configs[4]: M independent slabs of --atoms particles per GPU advanced by ONE batched "
                         "launch per step (64 slabs of 100k atoms over 8 GPUs = --ensemble 8 --atoms 100000)")
    ap.add_argument("--replicas", type=int, default=0, help="slab copies rotated through (0 = auto: > 2x L2)")
    ap.add_argument("--pairs-per-thread", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0, help="0 auto (one-shot LDG kernel), 1 LDG kernel, 2 TMA ring kernel")
    ap.add_argument("--tma-block", type=int, default=0)
    ap.add_argument("--tma-stages", type=int, default=0)
    ap.add_argument("--tma-ctas", type=int, default=0)
    ap.add_argument("--chains", type=int, default=2,
                    help="cut the slab into this many independent real-atom ranges, each advanced on its own stream "
                         "(constv_set_chains): the drain of one range's step overlaps the ramp of another's; for --ensemble, "
                         "the slabs of a step are advanced by this many batched launches on as many streams")
    ap.add_argument("--rigid-water", default="", choices=["", "fused", "split"],
                    help="BASELINE configs[0-1] shape: the slab's water molecules are rigid (constraints=AllBonds on SPC/E, "
                         "example/nacl_1m_spce/nacl_constv.py:56).  fused = constv_step_fused_settle, one kernel per step; "
                         "split = the reference's structure, part 1 | constraint solver on posDelta | part 2 + image rewrite "
                         "(three launches; implies --chains 1).")
    ap.add_argument("--drude", action="store_true",
                    help="ConstVDrudeLangevinIntegrator on a slab of polarizable 4-site water (SURVEY.md 8f.4): one fused kernel "
                         "per step (per chain), timed beside the reference's own four Drude kernels on the same GPU.")
    ap.add_argument("--reordered", action="store_true",
                    help="the state behind OpenMM after cu.reorderAtoms(): whole molecules (and their images) have traded "
                         "slots inside spatial windows; the step addresses images through atomIndex / invAtomOrder "
                         "(constVLangevin.cu:142-158).  Implies --chains 1.")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cache-image-charge", action="store_true")
    ap.add_argument("--keep-in-l2", action="store_true", help="default cache policy instead of evict-first")
    ap.add_argument("--temperature", type=float, default=T_KELVIN, help="experiments only: 0 disables the noise term")
    ap.add_argument("--e2e-steps", type=int, default=40)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-ref-gpu", action="store_true", help="do not time the reference's own CUDA kernels beside ours")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def traffic_from_profile(atoms, chains):
    """dram bytes per launch from the committed ncu capture, if it was taken on this workload."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        rec = json.load(open(path))
        if int(rec.get("atoms", -1)) == int(atoms) and int(rec.get("chains", 1)) == int(chains):
            return rec["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index, period=0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                                  ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                                  ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                                  ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)):
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
        else:
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index), "--format=csv,noheader,nounits",
                     "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                     "clocks_event_reasons.sw_power_cap"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(int(f[0]))
                self.max_mhz = int(f[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass

    def run(self):
        while not self._halt.is_set():
            self.sample()
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        if not self.samples:
            self.sample()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU legs (the only place bench.py touches oracle/) ---------------------------------------------

def cpu_oracle_rate(slab, seconds, steps_cap=None, window=None, replicas=1):
    """Times the oracle's single-pass fused step (oracle_step_range_f32) with all host threads,
    rotating over `replicas` copies of the slab exactly as the GPU arm does (so neither side runs
    out of its last-level cache).  Returns (atom_steps_per_s, steps, threads, atoms_per_step, s)."""
    import oracle
    copies = [slab.copy() for _ in range(max(1, replicas))]
    s = copies[0]
    prm = oracle.params(T_KELVIN, FRICTION, DT).astype(np.float32 if s.precision == "single" else np.float64)
    dt = (np.float32 if s.precision == "single" else np.float64)(DT)
    xi = oracle.fill_noise(s.padded, 1, 0)
    R = s.num_real
    w = R if window is None else max(1, min(R, window))
    threads = oracle.num_threads()
    per_pass = (R + w - 1) // w

    def one(k):
        c = copies[(k // per_pass) % len(copies)]
        lo = (k % per_pass) * w
        hi = min(R, lo + w)
        oracle.step_range(c.precision, c.num_atoms, c.num_cells, c.zmax, c.posq, c.corr, c.velm,
                          c.force.reshape(-1), prm, dt, xi, 0, lo, hi)
        return (hi - lo) * c.num_cells

    for k in range(2):
        one(k)
    done, atoms = 0, 0
    t0 = time.perf_counter()
    while True:
        atoms += one(done)
        done += 1
        el = time.perf_counter() - t0
        if (steps_cap is not None and done >= steps_cap) or (steps_cap is None and el >= seconds):
            break
    el = time.perf_counter() - t0
    return atoms / el, done, threads, atoms / done, el


def cpu_reference_rate(slab, seconds, steps_cap=None, window=None, replicas=1):
    """Times the REFERENCE'S OWN kernel source compiled for the host (oracle/_ref/
    libconstv_ref_host.so: constVLangevin.cu unmodified, CUDA execution model emulated, one OpenMP
    thread per contiguous range of atoms; part 1, part 2 and the image rewrite as three passes,
    the launch sequence of CudaConstVKernels.cpp:146-166) on all host threads, rotating over
    `replicas` copies of the slab like the GPU arm.  Same return value as cpu_oracle_rate."""
    import oracle
    from oracle import ref
    copies = [slab.copy() for _ in range(max(1, replicas))]
    s = copies[0]
    m = np.float32 if s.precision == "single" else np.float64
    prm = oracle.params(T_KELVIN, FRICTION, DT).astype(m)
    dt = m(DT)
    xi = oracle.fill_noise(s.padded, 1, 0)
    inv = np.arange(s.padded, dtype=np.int32)
    deltas = [np.zeros((s.padded, 4), m) for _ in copies]
    R = s.num_real
    w = R if window is None else max(1, min(R, window))
    threads = ref.host_num_threads()
    per_pass = (R + w - 1) // w

    def one(k):
        j = (k // per_pass) % len(copies)
        c = copies[j]
        lo = (k % per_pass) * w
        hi = min(R, lo + w)
        ref.host_step_range(c.precision, c.num_atoms, c.num_cells, c.zmax, c.posq, c.corr, c.velm,
                            c.force.reshape(-1), deltas[j], prm, dt, xi, 0, inv, lo, hi)
        return (hi - lo) * c.num_cells

    for k in range(2):
        one(k)
    done, atoms = 0, 0
    t0 = time.perf_counter()
    while True:
        atoms += one(done)
        done += 1
        el = time.perf_counter() - t0
        if (steps_cap is not None and done >= steps_cap) or (steps_cap is None and el >= seconds):
            break
    el = time.perf_counter() - t0
    return atoms / el, done, threads, atoms / done, el


def host_thread_count():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_leg():
    """(rate function, kind): the reference's own source when oracle/_ref travelled with the tree,
    else the oracle port."""
    import oracle
    from oracle import ref
    n = host_thread_count()
    oracle.set_num_threads(n)  # torchrun exports OMP_NUM_THREADS=1; the CPU legs use every host thread
    if ref.available("host"):
        ref.host_set_num_threads(n)
        return cpu_reference_rate, "reference"
    return cpu_oracle_rate, "port"


def run_reference_arm(args):
    """--impl reference: the reference's C++ host code needs OpenMM and ships no CPU platform for
    this path, but its kernel source is self-contained: the reference arm times
    that source compiled for the host (oracle/_ref, kind "reference") on all host threads, and
    reports the faster single-pass oracle port beside it.  Without oracle/_ref the port is timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from openmm_constv_b200 import slab as slabmod
    rate_fn, kind = cpu_leg()
    total_atoms = args.atoms * max(1, args.gpus)  # the whole job's slab, as the GPU arm partitions it
    s = slabmod.make_slab(total_atoms // 2, args.precision, seed=12345)
    # bound the run: calibrate, then size the per-step window so warmup+steps stay under ~90 s
    reps = 6 if total_atoms <= 2_000_000 else 2
    rate, _, threads, _, _ = rate_fn(s, 1.0, replicas=reps)
    total = max(1, args.steps + args.warmup)
    window_atoms = min(total_atoms, max(65536, int(rate * 90.0 / total)))
    window = max(1, window_atoms // s.num_cells)
    rate_fn(s, 0, steps_cap=max(1, args.warmup), window=window, replicas=reps)
    rate, done, threads, per_step, el = rate_fn(s, 0, steps_cap=args.steps, window=window, replicas=reps)
    what = ("the reference's constVLangevin.cu compiled for the host (part 1, part 2, image rewrite)" if kind == "reference"
            else "the fused oracle step")
    sample = (f"{done} steps, each {what} over a window of {int(per_step)} atoms walking through "
              f"{reps} replicas of the {total_atoms}-atom slab (like the GPU arm's L2 rotation), injected noise pool")
    port = None
    if kind == "reference":
        prate, pdone, _, _, pel = cpu_oracle_rate(s, min(10.0, args.cpu_seconds), window=window, replicas=reps)
        port = {"value": prate, "unit": "atom-steps/s", "kind": "port", "cores": threads,
                "sample": f"{pdone} single-pass oracle steps over the same windows in {pel:.1f} s"}
    line = {
        "impl": "reference", "metric": "atom-steps/s", "value": rate, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision == "single" else "f64", "data": "synthetic",
        "config": {"workload": f"synthetic slab, {total_atoms} atoms ({total_atoms // 2} real + images), fused ConstV Langevin step only",
                   "atoms_per_step": int(per_step), "precision": args.precision,
                   "note": "the reference ships no CPU platform for this path; its kernel source compiled for the host is timed"
                           if kind == "reference" else "oracle/_ref absent: oracle port timed"},
        "cpu_baseline": {"value": rate, "unit": "atom-steps/s", "cores": threads, "kind": kind, "sample": sample, "port": port},
        "e2e": {"value": rate, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def time_reference_gpu(torch, ctxs, shard, stream, steps, max_blocks, use_graph):
    """The reference's own kernels (constVLangevin.cu compiled by nvcc for sm_100a, oracle/_ref)
    on the same rotating slab replicas: per step the three launches of CudaConstVKernels.cpp:146-166
    (part 1 reading a pre-filled float4 Gaussian pool, part 2, image rewrite), 128-thread blocks,
    grid capped at `max_blocks` (0 = uncapped).  The pool refill (OpenMM's RNG kernel) and the
    reordering kernel are NOT charged to the reference.  Returns ms per step."""
    import oracle
    from oracle import ref
    dev = ctxs[0][0].device
    m = np.float32 if shard.precision == "single" else np.float64
    prm = oracle.params(T_KELVIN, FRICTION, DT).astype(m)
    r = ref.CudaReference(shard.precision, prm, m(DT), dev, max_blocks=max_blocks)
    pool = torch.randn((shard.padded, 4), dtype=torch.float32, device=dev)
    inv = torch.arange(shard.padded, dtype=torch.int32, device=dev)
    sptr = stream.cuda_stream
    n = len(ctxs)

    def launch(first, count):
        for k in range(first, first + count):
            c = ctxs[k % n][0]
            r.step(shard.num_atoms, shard.num_cells, shard.zmax, c.posq, c.corr, c.velm, c.force, c.pos_delta, pool, 0,
                   inv, stream=sptr)

    with torch.cuda.stream(stream):
        launch(0, n)
        stream.synchronize()
        graph, chunk = None, 0
        if use_graph and steps >= 2 * n:
            chunk = n * max(1, min(20, steps // n))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                launch(0, chunk)

        def run(nsteps):
            done = 0
            if graph is not None:
                while done + chunk <= nsteps:
                    graph.replay()
                    done += chunk
            launch(done, nsteps - done)

        run(max(3, steps // 10))
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(steps)
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) / steps


# ---- the B200 arm -------------------------------------------------------------------------------------

def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from openmm_constv_b200 import _capi, integrator as hostapi, partition, slab as slabmod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ConstV step path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.load()

    # ---- workload: this rank's shard of the range-partitioned slab ------------------------------
    R = args.atoms // 2
    water = None
    drude = None
    if args.drude:
        if args.ensemble or args.rigid_water:
            raise SystemExit("--drude, --rigid-water and --ensemble are separate configurations")
        if args.variant == 1:
            args.chains = 1
        shard, d_normal, d_pairs = slabmod.make_drude_slab(R, args.precision, seed=12345 + rank)
        drude = (d_normal, d_pairs)
    elif args.rigid_water:
        if args.ensemble and args.rigid_water != "fused":
            raise SystemExit("--ensemble takes --rigid-water fused only")
        if args.rigid_water == "split" or args.variant == 1:
            args.chains = 1
        shard, w_atoms, w_params, w_cons = slabmod.make_water_slab(R, args.precision, seed=12345 + rank)
        water = (w_atoms, w_params, w_cons)
    else:
        shard = slabmod.make_slab(R, args.precision, seed=12345 + rank)
    shard.global_atom_offset = partition.shard_range(R * world, rank, world)[0]
    reorder_perm = None
    if args.reordered:
        if args.ensemble:
            raise SystemExit("--reordered and --ensemble are separate configurations")
        args.chains = 1
        per = 4 if drude else 3
        n_mol = shard.meta["n_pairs"] if drude else shard.meta["n_water"] // 3
        reorder_perm = slabmod.molecule_permutation(shard, shard.meta["n_ion"], per, n_mol, seed=7 + rank)
        shard, _ = slabmod.permute_slab(shard, reorder_perm)
    # the staged rigid-water kernel finds its slots arithmetically; only the gather kernel (--variant 1) reads lists
    alg_bytes = slabmod.algorithmic_bytes_per_step(shard, rigid_molecules=len(water[0]) if (water and args.variant == 1) else 0,
                                                   drude=drude is not None, drude_pairs=len(drude[1]) if drude else 0)
    rec_bytes = shard.posq.nbytes + shard.velm.nbytes + shard.force.nbytes + (shard.corr.nbytes if shard.corr is not None else 0)
    M = max(1, args.ensemble)  # slabs advanced per step on this GPU (1 = the plain slab configs)
    replicas = args.replicas or max(2, int(np.ceil(2.0 * L2_BYTES / (rec_bytes * M))) + 1)

    flags = _capi.FLAGS_DEFAULT
    if not args.no_cache_image_charge:
        flags |= _capi.FLAG_CACHE_IMAGE_CHARGE
    if args.keep_in_l2:
        flags |= _capi.FLAG_KEEP_IN_L2

    ctxs = []
    for r in range(replicas * M):
        sysm = hostapi.SlabSystem()
        sysm.addParticles(shard.masses)
        sysm.setDefaultPeriodicBoxVectors((shard.box[0], 0, 0), (0, shard.box[1], 0), (0, 0, shard.num_cells * shard.zmax))
        if water:
            sysm.addConstraints(*water[2])
        if drude:
            dforce = hostapi.DrudeForce()
            for a, b in drude[1]:
                dforce.addParticle(int(a), int(b))
            sysm.addForce(dforce)
            it = hostapi.ConstVDrudeLangevinIntegrator(args.temperature, DRUDE_FRICTION_CM, DRUDE_T, DRUDE_FRICTION, DT, shard.num_cells)
            it.setMaxDrudeDistance(DRUDE_WALL)
        else:
            it = hostapi.ConstVLangevinIntegrator(args.temperature, FRICTION, DT, shard.num_cells)
        it.setRandomNumberSeed(2024 + (r % M))
        ctx = hostapi.SlabContext(sysm, it, precision=args.precision, device=local_rank, flags=flags,
                                  global_atom_offset=shard.global_atom_offset, padded=shard.padded)
        ctx.load_slab(shard)
        if reorder_perm is not None:
            ctx.set_atom_order(reorder_perm)
            _capi.check(lib.constv_update_inverse_order(it._kernel.handle, ctx.stream_ptr()))
        if args.pairs_per_thread or args.ctas_per_sm:
            it._kernel.set_launch_shape(args.pairs_per_thread, args.ctas_per_sm)
        it._kernel.set_kernel_variant(args.variant, args.tma_block, args.tma_stages, args.tma_ctas)
        if drude:
            _capi.check(lib.constv_drude_set_parameters(it._kernel.handle, args.temperature, DRUDE_FRICTION_CM, DRUDE_T,
                                                        DRUDE_FRICTION, DRUDE_WALL, DT))
        else:
            _capi.check(lib.constv_set_parameters(it._kernel.handle, args.temperature, FRICTION, DT))
        ctxs.append((ctx, it))
    handles = [it._kernel.handle for _, it in ctxs]

    stream = torch.cuda.Stream(device=dev)
    sptr = stream.cuda_stream

    # ---- kinetic energy report: the ONE collective of the partition (on demand, never per step),
    # taken on the initial state (the timed loop below keeps the injected forces constant, so the
    # state it ends in is not physical)
    ke_dev = torch.zeros(1, dtype=torch.float64, device=dev)
    ctxs[0][1]._kernel.kinetic_energy_device(0.0, ke_dev)
    torch.cuda.synchronize()
    partition.allreduce_scalar_sum(ke_dev)  # NCCL over NVLink when world > 1
    ke_total = float(ke_dev.item())
    n_massive = int(np.sum(shard.velm[:R, 3] != 0)) * world
    temperature = partition.temperature_from_kinetic_energy(ke_total, n_massive, slabmod.BOLTZ)

    groups = [(C.c_void_p * M)(*handles[g * M:(g + 1) * M]) for g in range(replicas)]

    chains = 1
    if args.chains > 1 and not args.ensemble:
        got = C.c_int32(0)
        for h in handles:
            _capi.check(lib.constv_set_chains(h, args.chains))
        _capi.check(lib.constv_get_chains(handles[0], C.byref(got)))
        chains = got.value
    elif args.chains > 1:
        chains = min(args.chains, M)
    side_streams = [torch.cuda.Stream(device=dev) for _ in range(chains - 1)]
    # ensemble: sub-group q of every replica's M slabs = one batched launch on stream q
    bounds = [M * q // chains for q in range(chains + 1)]
    subgroups = [[(C.c_void_p * (bounds[q + 1] - bounds[q]))(*handles[g * M + bounds[q]:g * M + bounds[q + 1]])
                  for q in range(chains)] for g in range(replicas)]

    def launch_steps(first, count):
        if chains > 1:  # fork: chain q's steps on stream q, in step order; join back into `stream`
            fork = torch.cuda.Event()
            fork.record(stream)
            for q, cs in enumerate([stream] + side_streams):
                if q:
                    cs.wait_event(fork)
                for k in range(first, first + count):
                    if args.ensemble and water:  # rigid-water ensemble: one batched launch per sub-group and stream
                        _capi.check(lib.constv_step_fused_settle_batched(subgroups[k % replicas][q], bounds[q + 1] - bounds[q], cs.cuda_stream))
                    elif args.ensemble:
                        _capi.check(lib.constv_step_fused_batched(subgroups[k % replicas][q], bounds[q + 1] - bounds[q], cs.cuda_stream))
                    elif drude:
                        _capi.check(lib.constv_drude_step_fused_chain(handles[k % replicas], q, None, 0, cs.cuda_stream))
                    elif args.rigid_water == "fused":
                        _capi.check(lib.constv_step_fused_settle_chain(handles[k % replicas], q, None, 0, cs.cuda_stream))
                    else:
                        _capi.check(lib.constv_step_fused_chain(handles[k % replicas], q, None, 0, cs.cuda_stream))
            for cs in side_streams:
                join = torch.cuda.Event()
                join.record(cs)
                stream.wait_event(join)
            return
        for k in range(first, first + count):
            if args.ensemble and water:
                _capi.check(lib.constv_step_fused_settle_batched(groups[k % replicas], M, sptr))
            elif args.ensemble:
                _capi.check(lib.constv_step_fused_batched(groups[k % replicas], M, sptr))
            elif drude:
                _capi.check(lib.constv_drude_step_fused(handles[k % replicas], None, 0, sptr))
            elif args.rigid_water == "fused":
                _capi.check(lib.constv_step_fused_settle(handles[k % replicas], None, 0, sptr))
            elif args.rigid_water == "split":
                h = handles[k % replicas]
                _capi.check(lib.constv_step_part1(h, None, 0, sptr))
                _capi.check(lib.constv_settle_positions(h, sptr))
                _capi.check(lib.constv_step_part2_mirror(h, sptr))
            else:
                _capi.check(lib.constv_step_fused(handles[k % replicas], None, 0, sptr))

    # ---- optional CUDA graph: `chunk` consecutive steps per replay --------------------------------
    K, W = args.steps, max(args.warmup, 3)
    use_graph = (not args.no_graph) and K >= 2 * replicas
    graph, chunk = None, 0
    with torch.cuda.stream(stream):
        launch_steps(0, replicas)  # first touch, refreshes image-charge snapshots outside capture
        stream.synchronize()
        if use_graph:
            chunk = replicas * max(1, min(20, K // replicas))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                launch_steps(0, chunk)

        def run(nsteps):
            done = 0
            if graph is not None:
                while done + chunk <= nsteps:
                    graph.replay()
                    done += chunk
            launch_steps(done, nsteps - done)

        run(W)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        launches0 = _capi.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        run(K)
        ev1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        if world > 1:
            dist.barrier()
    ms = ev0.elapsed_time(ev1)
    eager_launches = _capi.launch_count() - launches0

    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    launch_ms_for_ref = ms_max / K
    atoms_per_step_job = shard.num_atoms * M * world
    value = atoms_per_step_job * K / (ms_max * 1e-3)

    # ---- e2e: host buffers through constv_step_host -------------------------------------------------
    # Per step: H2D of the step's input (the real atoms' fixed-point forces, pinned memory), the fused
    # step, the kinetic-energy reduction, and D2H of the step's result.  `value` reads back the scalar
    # a StateDataReporter asks for every step (kinetic energy); `full_state` also reads back posq of
    # all atoms (what a DCD frame needs) and is reported beside it.
    e2e = None
    if not args.skip_e2e and not args.ensemble and not args.rigid_water and not drude:
        ctx0, it0 = ctxs[0]
        n_e2e = max(3, args.e2e_steps)
        host_force = [torch.from_numpy(np.ascontiguousarray(slabmod.new_forces(shard, j)[:, :R])).pin_memory() for j in range(2)]
        esize = 4 if args.precision != "double" else 8
        host_posq = torch.empty((shard.num_atoms, 4), dtype=torch.float32 if esize == 4 else torch.float64).pin_memory()
        h2d = host_force[0].numel() * 8
        ke_host = C.c_double(0.0)

        def e2e_run(with_posq):
            out = host_posq.data_ptr() if with_posq else None
            for j in range(3):
                _capi.check(lib.constv_step_host(it0._kernel.handle, host_force[j % 2].data_ptr(), out, None, C.byref(ke_host), sptr))
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for j in range(n_e2e):
                _capi.check(lib.constv_step_host(it0._kernel.handle, host_force[j % 2].data_ptr(), out, None, C.byref(ke_host), sptr))
            e1.record(stream)
            stream.synchronize()
            ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            return float(ems.item())

        def e2e_pipelined(buffers, code, nsteps):
            """submit step k+1, then collect step k: the upload of k+1 overlaps the kernels and the
            read-back of k.  Every step still uploads its own forces and returns its own energy."""
            h = it0._kernel.handle
            ticket = C.c_uint64(0)

            def loop(n):
                prev = None
                for j in range(n):
                    _capi.check(lib.constv_step_host_submit(h, buffers[j % 2].data_ptr(), code, C.byref(ticket), sptr))
                    if prev is not None:
                        _capi.check(lib.constv_step_host_collect(h, prev, C.byref(ke_host)))
                    prev = ticket.value
                _capi.check(lib.constv_step_host_collect(h, prev, C.byref(ke_host)))

            loop(4)
            best = float("inf")
            for _ in range(3):  # PCIe transfers are noisy on a shared host: best of three passes
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                loop(nsteps)
                e1.record(stream)
                stream.synchronize()
                best = min(best, e0.elapsed_time(e1))
            ems = torch.tensor([best], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            return float(ems.item())

        ms_ke = e2e_run(False)
        ms_full = e2e_run(True)
        n_pipe = 4 * n_e2e
        ms_pipe = e2e_pipelined(host_force, _capi.FORCE_FIXED64, n_pipe)
        host_force32 = [(torch.randn((3, R), dtype=torch.float32) * 1000.0).pin_memory() for _ in range(2)]
        ms_pipe32 = e2e_pipelined(host_force32, _capi.FORCE_FLOAT32, n_pipe)
        e2e = {"value": atoms_per_step_job * n_pipe / (ms_pipe * 1e-3), "unit": "atom-steps/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8, "steps": n_pipe,
               "ms_per_step": ms_pipe / n_pipe,
               "what": "constv_step_host_submit/_collect, two tickets in flight: every step uploads its real-atom forces "
                       "(int64 fixed point, OpenMM's buffer format, pinned memory) and reads back its kinetic energy; the "
                       "upload of step k+1 overlaps the kernels of step k; PCIe-bound",
               "float32_forces": {"value": atoms_per_step_job * n_pipe / (ms_pipe32 * 1e-3), "ms_per_step": ms_pipe32 / n_pipe,
                                  "h2d_bytes_per_step": int(h2d // 2), "d2h_bytes_per_step": 8,
                                  "what": "same with float32 host forces converted to fixed point on the device"},
               "synchronous": {"value": atoms_per_step_job * n_e2e / (ms_ke * 1e-3), "ms_per_step": ms_ke / n_e2e,
                               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
                               "what": "constv_step_host: upload, fused step, kinetic energy, read-back, one step at a time"},
               "full_state": {"value": atoms_per_step_job * n_e2e / (ms_full * 1e-3), "ms_per_step": ms_full / n_e2e,
                              "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(host_posq.numel() * esize + 8),
                              "what": "synchronous, plus D2H of posq for all atoms"}}

    # ---- the reference's own CUDA kernels on this GPU (information beside the headline) ---------------
    ref_gpu = None
    if rank == 0 and world == 1 and not args.skip_ref_gpu and not args.ensemble and not args.rigid_water:
        from oracle import ref as refmod
        if refmod.available("cuda"):
            for ctx, _ in ctxs:
                ctx.load_slab(shard)
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            n_ref = max(60, min(K, 6000))
            ms_cap = time_reference_gpu(torch, ctxs, shard, stream, n_ref, 6 * sms, graph is not None)
            ms_unc = time_reference_gpu(torch, ctxs, shard, stream, n_ref, 0, graph is not None)
            best = min(ms_cap, ms_unc)
            ref_gpu = {"value": shard.num_atoms / (best * 1e-3), "unit": "atom-steps/s", "ms_per_step": best,
                       "ms_per_step_grid_6_blocks_per_sm": ms_cap, "ms_per_step_grid_uncapped": ms_unc, "steps": n_ref,
                       "launches_per_step": 3,
                       "what": "the reference's constVLangevin.cu compiled by nvcc for sm_100a (oracle/_ref): part 1 (reads a "
                               "pre-filled Gaussian pool; refill not charged), part 2, image rewrite; 128-thread blocks; same "
                               "rotating replicas, same CUDA-graph launch; value = the faster of OpenMM's capped grid "
                               "(6 blocks/SM, recalled) and an uncapped grid",
                       "speedup_of_fused_step": best / launch_ms_for_ref}

    if rank == 0 and world == 1 and not args.skip_ref_gpu and drude:
        from oracle import ref as refmod
        import oracle
        if refmod.available("cuda") and hasattr(refmod.cuda(), "constv_ref_drude_cuda_part1_single"):
            for ctx, _ in ctxs:
                ctx.load_slab(shard)
            m = np.float32 if args.precision == "single" else np.float64
            scal = oracle.drude_params(args.temperature, DRUDE_FRICTION_CM, DRUDE_T, DRUDE_FRICTION, DRUDE_WALL, DT)
            rr = refmod.CudaDrudeReference(args.precision, scal, m(DT), dev, drude[0], drude[1], R, shard.padded, max_blocks=0)
            pool = torch.randn((R + 8, 4), dtype=torch.float32, device=dev)
            inv = torch.arange(shard.padded, dtype=torch.int32, device=dev)
            n_ref = max(60, min(K, 6000))

            def ref_launch(first, count):
                for k in range(first, first + count):
                    c = ctxs[k % replicas][0]
                    rr.drude_step(shard.num_atoms, shard.num_cells, shard.zmax, c.posq, c.corr, c.velm, c.force, c.pos_delta,
                                  pool, 0, inv, stream=sptr)

            with torch.cuda.stream(stream):
                ref_launch(0, replicas)
                stream.synchronize()
                rchunk = replicas * max(1, min(20, n_ref // replicas))
                rgraph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(rgraph, stream=stream):
                    ref_launch(0, rchunk)
                for _ in range(3):
                    rgraph.replay()
                stream.synchronize()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = max(1, n_ref // rchunk)
                r0.record(stream)
                for _ in range(reps):
                    rgraph.replay()
                r1.record(stream)
                stream.synchronize()
            ms_ref = r0.elapsed_time(r1) / (reps * rchunk)
            ref_gpu = {"value": shard.num_atoms / (ms_ref * 1e-3), "unit": "atom-steps/s", "ms_per_step": ms_ref, "steps": reps * rchunk,
                       "launches_per_step": 4,
                       "what": "the reference's constVDrudeLangevin.cu + constVLangevin.cu compiled by nvcc for sm_100a (oracle/_ref): "
                               "Drude part 1 (pre-filled Gaussian pool; refill not charged), part 2, hard wall, image rewrite; 128-thread "
                               "blocks, uncapped grid; same rotating replicas, same CUDA-graph launch",
                       "speedup_of_fused_step": ms_ref / launch_ms_for_ref}

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu and not args.ensemble and not args.rigid_water and not drude:
        rate_fn, kind = cpu_leg()
        rate, done, threads, per_step, el = rate_fn(shard, args.cpu_seconds, replicas=replicas)
        what = ("steps of the reference's constVLangevin.cu compiled for the host (part 1, part 2, image rewrite; oracle/_ref)"
                if kind == "reference" else "fused oracle steps")
        cpu = {"value": rate, "unit": "atom-steps/s", "cores": threads, "kind": kind,
               "sample": f"{done} full {what} rotating over {replicas} replicas of the same {shard.num_atoms}-atom slab in {el:.1f} s, all host threads, injected noise pool",
               "host_cpus": os.cpu_count()}
        if kind == "reference":
            prate, pdone, _, _, pel = cpu_oracle_rate(shard, args.cpu_seconds / 2, replicas=replicas)
            cpu["port"] = {"value": prate, "unit": "atom-steps/s", "kind": "port", "cores": threads,
                           "sample": f"{pdone} single-pass oracle steps in {pel:.1f} s (a tuned CPU port, for scale)"}

    if rank == 0:
        peak, peak_src = peaks()
        launch_ms = ms_max / K
        alg_bytes *= M
        achieved = alg_bytes / (launch_ms * 1e-3) / 1e9  # = (alg_bytes / chains) x (K x chains launches) / time
        line = {
            "metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": launch_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "single" else ("f32+f64" if args.precision == "mixed" else "f64"),
            "data": "synthetic",
            "config": {
                "workload": (f"ensemble of {M * world} independent synthetic RIGID-WATER slabs of {shard.num_atoms} atoms, {M} per GPU, advanced by {chains} batched constraint-solving launch(es) per step (BASELINE configs[4] with the example's constraints)"
                             if (args.ensemble and water) else
                             f"ensemble of {M * world} independent synthetic slabs of {shard.num_atoms} atoms, {M} per GPU advanced by one batched launch per step (BASELINE configs[4])"
                             if args.ensemble else
                             f"synthetic slab with RIGID water ({len(water[0])} SPC/E molecules, constraints solved {'inside the step kernel' if args.rigid_water == 'fused' else 'by a separate launch between part 1 and part 2'}), {shard.num_atoms} atoms per GPU, ConstV Langevin step incl. constraints (BASELINE configs[0-1] shape); range partition over {world} GPU(s)"
                             if water else
                             f"synthetic slab of polarizable water ({len(drude[1])} Drude pairs, {len(drude[0])} ordinary particles), {shard.num_atoms} atoms per GPU, fused ConstVDrudeLangevinIntegrator step incl. hard wall (SURVEY 8f.4); range partition over {world} GPU(s)"
                             if drude else
                             f"synthetic slab, {shard.num_atoms} atoms per GPU ({R} real + {shard.num_atoms - R} image), fused ConstV Langevin step only (BASELINE configs[2]); range partition over {world} GPU(s)"),
                "atoms_per_gpu": shard.num_atoms, "total_atoms": atoms_per_step_job, "precision": args.precision,
                "pair_classes": shard.pair_classes(), "rng": "in-kernel Philox4x32-10",
                "l2": f"rotating {replicas} replicas of the per-step working set ({replicas * M * rec_bytes / 1e6:.0f} MB) > 126 MB L2: every step streams from HBM",
                "launch": ("CUDA graph of %d steps, PDL" % chunk) if graph is not None else "eager, PDL",
                "chains": chains, "reordered_slots": bool(args.reordered),
                "image_charge": "re-read from posq" if args.no_cache_image_charge else "side array",
                "kernel_variant": {"variant": args.variant, "tma_block": args.tma_block, "tma_stages": args.tma_stages,
                                   "tma_ctas": args.tma_ctas, "pairs_per_thread": args.pairs_per_thread, "ctas_per_sm": args.ctas_per_sm},
                "T_K": args.temperature, "friction_per_ps": FRICTION, "dt_ps": DT,
            },
            "steps_per_s": K / (ms_max * 1e-3),
            "atom_steps_per_s_per_gpu": value / world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if args.ensemble else traffic_from_profile(shard.num_atoms, chains), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / chains, "launches_per_step": chains,
                         "algorithmic_bytes_per_step": alg_bytes,
                         "note": ("a step is %d launches, one per chain (contiguous range of real atoms) on its own stream; "
                                  "achieved = algorithmic bytes per launch x launches in the timed region / its duration" % chains)
                                 if chains > 1 else "one launch per step",
                         "frac_of_8TBs_nominal": achieved / 8000.0,
                         "kernel": "constv_settle_step_tiled_batched_kernel" if (args.ensemble and water) else
                                   "constv_fused_step_batched_kernel" if args.ensemble else
                                   "constv_drude_step_kernel" if drude else
                                   ("constv_settle_step_tiled_kernel" if args.rigid_water == "fused" else
                                    ("constv_part1_kernel + constv_settle_positions_kernel + constv_part2_mirror_kernel"
                                     if args.rigid_water == "split" else "constv_fused_step_kernel"))},
            "cpu_baseline": cpu,
            "reference_gpu": ref_gpu,
            "e2e": e2e,
            "gpu_launches": int(K * (3 if args.rigid_water == "split" else chains)),
            "eager_launches_in_region": int(eager_launches),
            "clocks": clocks,
            "kinetic_energy_kj_mol": ke_total, "temperature_K": temperature,
        }
        print(json.dumps(line), flush=True)
    for ctx, _ in ctxs:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
