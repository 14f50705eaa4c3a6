#!/usr/bin/env python
"""Benchmark of the ConstVLangevinIntegrator step path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one fused integrator step over this rank's shard of ONE seeded global synthetic slab of
``--atoms`` particles per GPU (default 1,000,000 = 500k real + 500k image per GPU: BASELINE.json
configs[2] at N = 1).  Metric: atom-steps/s, whole job (atoms of all ranks / max-over-ranks device time).
Multi-GPU is the range partition of SURVEY.md 8e: the global slab is ``--gpus`` lateral tiles
(slab.slab_tile), rank r owns partition.shard_range(R_total, r, N) = tile r and its images, and there is NO
data-path collective (weak scaling); the only collective is the scalar kinetic-energy all-reduce, issued
once outside the timed region as a report, never per step.

How the K steps are timed (the same at --steps 20 and at --steps 100000):
  * every launch inside the timed region comes from a CUDA graph -- never an eager launch
    (``eager_launches_in_region`` = 0).  Small K (the driver's --steps 20): ONE graph holds a train of K-step
    regions with an event-record node between regions, so a region is exactly K steps between two events on the
    launching stream and costs no graph relaunch.  Large K: 120-step chunk graphs plus one tail graph, stream
    events between regions;
  * after W warm-up steps and untimed regions that upload every graph and keep the GPU busy, the K-step region
    is run ``repeats`` (>= 9) times back to back; ms_per_step = MEDIAN region time / K, min and max reported;
    for N > 1 the MAX over ranks of each rank's median;
  * clocks and throttle reasons are sampled every 10 ms over the whole repeat loop (the number of repeats is
    raised until the loop lasts ~0.4 s);
  * L2: a 1M-atom slab is 56 MB < B200's 126 MB L2, so the steps rotate over ``replicas`` copies of the slab
    whose footprint exceeds 2x L2: every step streams its slab from HBM.

Chains (default 3): the slab is cut into contiguous ranges of real atoms, each with its own device step
counter and stream (constv_set_chains / constv_step_fused_chain), so that the drain of one range's step k
overlaps the ramp-up of another's step k+1; bit-identical to one launch per step.  (Same box, the driver's
arguments: 2 chains 11.48-11.56 us, 3 chains 11.28-11.29, 4 chains 11.32, 8 chains 11.27.)  A host that runs other
kernels between two steps (OpenMM) cannot use that overlap: ``also.single_launch`` is the same workload with
one launch per step and ``also.plugin_path`` steps through CudaIntegrateConstVLangevinStepKernel::execute
(plugin/tests/BenchPluginPath.cpp), a cache-sweeping kernel between two steps.

The JSON line carries, besides the base contract's keys:
  roofline      algorithmic bytes of one launch / average launch duration, against MEASURED_PEAKS.json
  cpu_baseline  the reference's kernel source compiled for the host (oracle/_ref, kind "reference"; else the
                oracle port) timed on this box's host cores on a bounded sample of the same workload
  e2e           the same metric through the C ABI with HOST buffers: per step H2D of the real atoms' forces
                from pinned memory, the fused step, D2H of the kinetic energy (full_state: + posq)
  clocks        SM clock / throttle reasons sampled during the timed region
  also          (last key, so that it survives a truncated log) single_launch, plugin_path, the 16M-atom slab
                of configs[3] over these N GPUs, the 64 x 100k ensemble of configs[4], the same slab size with
                rigid SPC/E water solved inside the step (rigid_water) and with polarizable water (drude),
                partition_parity, e2e_full_state
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
GRAPH_CHUNK_STEPS = 120
SLAB_SEED = 12345


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10_000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--atoms", type=int, default=1_000_000, help="particles (real+image) per GPU")
    ap.add_argument("--slab", default="", choices=["", "16M"],
                    help="16M = BASELINE configs[3] as the headline: ONE 16M-atom slab range-partitioned over the N GPUs "
                         "(strong scaling: 16M/N atoms per GPU)")
    ap.add_argument("--precision", default="single", choices=["single", "mixed", "double"])
    ap.add_argument("--ensemble", type=int, default=0,
                    help="BASELINE configs[4] as the headline: M independent slabs of --atoms particles per GPU advanced by "
                         "batched launches (64 slabs of 100k atoms over 8 GPUs = --ensemble 8 --atoms 100000)")
    ap.add_argument("--replicas", type=int, default=0, help="slab copies rotated through (0 = auto: > 2x L2)")
    ap.add_argument("--repeats", type=int, default=0, help="timed K-step regions (0 = auto: >= 9, ~0.4 s in total)")
    ap.add_argument("--pairs-per-thread", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0, help="0 auto (one-shot LDG kernel), 1 LDG kernel, 2 TMA ring kernel")
    ap.add_argument("--tma-block", type=int, default=0)
    ap.add_argument("--tma-stages", type=int, default=0)
    ap.add_argument("--tma-ctas", type=int, default=0)
    ap.add_argument("--chains", type=int, default=3,
                    help="cut the slab into this many independent real-atom ranges, each advanced on its own stream "
                         "(constv_set_chains); for --ensemble, the slabs of a step are advanced by this many batched launches")
    ap.add_argument("--rigid-water", default="", choices=["", "fused", "split"],
                    help="BASELINE configs[0-1] shape: rigid SPC/E water (constraints=AllBonds, nacl_constv.py:56).  fused = "
                         "constv_step_fused_settle, one kernel per step; split = part 1 | solver | part 2 + image rewrite")
    ap.add_argument("--drude", action="store_true",
                    help="ConstVDrudeLangevinIntegrator on polarizable 4-site water (SURVEY.md 8f.4), one fused kernel per step")
    ap.add_argument("--reordered", action="store_true",
                    help="the state behind OpenMM after cu.reorderAtoms(): slots addressed through atomIndex / invAtomOrder "
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
    ap.add_argument("--skip-also", action="store_true", help="headline only: no single-launch / plugin / 16M / ensemble / rigid-water / Drude / parity legs")
    ap.add_argument("--only", default="", help="experiments: comma list of `also` legs to run (single,plugin,slab16m,ensemble,rigid,drude,parity)")
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
    """dram bytes per launch from the committed ncu capture, if it was taken on this workload: the capture's bytes per
    STEP (what the slab's records cost in DRAM traffic does not depend on how many launches a step is cut into)
    divided by this run's launches per step."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        rec = json.load(open(path))
        if int(rec.get("atoms", -1)) == int(atoms):
            per_step = rec["dram_bytes_per_launch"] * int(rec.get("chains", 1))
            return int(round(per_step / max(1, int(chains))))
    except Exception:
        pass
    return None


def auto_replicas(bytes_per_step_working_set):
    """Copies to rotate through so that no step finds its data in L2: footprint > 2x L2."""
    if bytes_per_step_working_set >= 2 * L2_BYTES:
        return 1
    return int(np.ceil(2.0 * L2_BYTES / bytes_per_step_working_set)) + 1


def workload_config(atoms_per_gpu, n_gpus, precision, slab16m=False):
    """The `config` object: identical in the B200 arm and the reference arm (what is computed, not how)."""
    total = 16_000_000 if slab16m else atoms_per_gpu * n_gpus
    per = total // n_gpus
    rec = per * (56 if precision == "single" else 104 if precision == "mixed" else 88)
    reps = auto_replicas(rec)
    name = ("synthetic 16M-atom slab (BASELINE configs[3])" if slab16m else
            "synthetic slab (BASELINE configs[2] per GPU)")
    return {"workload": f"{name}: {total} atoms ({total // 2} real + {total // 2} image), {per} per GPU over {n_gpus} GPU(s) by "
                        f"contiguous real-atom range, ConstV Langevin step only, injected forces",
            "atoms_per_gpu": per, "total_atoms": total, "precision": precision,
            "T_K": T_KELVIN, "friction_per_ps": FRICTION, "dt_ps": DT,
            "l2": f"steps rotate over {reps} replica(s) of the per-GPU slab ({reps * rec / 1e6:.0f} MB > 126 MB L2): inputs larger than L2, "
                  f"every step streams from memory"}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index, period=0.01):
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
            self.period = max(period, 0.05)

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
                "sm_min_mhz": float(np.min(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU legs (the only place bench.py times anything under oracle/) -----------------------------------

def _cpu_rate(step_range_fn, slab, seconds, steps_cap, window, replicas, threads):
    copies = [slab.copy() for _ in range(max(1, replicas))]
    R = slab.num_real
    w = R if window is None else max(1, min(R, window))
    per_pass = (R + w - 1) // w

    def one(k):
        j = (k // per_pass) % len(copies)
        lo = (k % per_pass) * w
        hi = min(R, lo + w)
        step_range_fn(copies[j], j, lo, hi)
        return (hi - lo) * slab.num_cells

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


def cpu_oracle_rate(slab, seconds, steps_cap=None, window=None, replicas=1):
    """Times the oracle's single-pass fused step (oracle_step_range_f32) with all host threads, rotating over
    `replicas` copies of the slab exactly as the GPU arm does (so neither side runs out of its last-level
    cache).  Returns (atom_steps_per_s, steps, threads, atoms_per_step, seconds)."""
    import oracle
    m = np.float32 if slab.precision == "single" else np.float64
    prm = oracle.params(T_KELVIN, FRICTION, DT).astype(m)
    dt = m(DT)
    xi = oracle.fill_noise(slab.padded, 1, 0)

    def fn(c, j, lo, hi):
        oracle.step_range(c.precision, c.num_atoms, c.num_cells, c.zmax, c.posq, c.corr, c.velm,
                          c.force.reshape(-1), prm, dt, xi, 0, lo, hi)

    return _cpu_rate(fn, slab, seconds, steps_cap, window, replicas, oracle.num_threads())


def cpu_reference_rate(slab, seconds, steps_cap=None, window=None, replicas=1):
    """Times the REFERENCE'S OWN kernel source compiled for the host (oracle/_ref/libconstv_ref_host.so:
    constVLangevin.cu unmodified, CUDA execution model emulated, one OpenMP thread per contiguous range of
    atoms; part 1, part 2 and the image rewrite as three passes, the launch sequence of
    CudaConstVKernels.cpp:146-166) on all host threads, rotating over `replicas` copies like the GPU arm."""
    import oracle
    from oracle import ref
    m = np.float32 if slab.precision == "single" else np.float64
    prm = oracle.params(T_KELVIN, FRICTION, DT).astype(m)
    dt = m(DT)
    xi = oracle.fill_noise(slab.padded, 1, 0)
    inv = np.arange(slab.padded, dtype=np.int32)
    deltas = {}

    def fn(c, j, lo, hi):
        if j not in deltas:
            deltas[j] = np.zeros((c.padded, 4), m)
        ref.host_step_range(c.precision, c.num_atoms, c.num_cells, c.zmax, c.posq, c.corr, c.velm,
                            c.force.reshape(-1), deltas[j], prm, dt, xi, 0, inv, lo, hi)

    return _cpu_rate(fn, slab, seconds, steps_cap, window, replicas, ref.host_num_threads())


def host_thread_count():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_leg():
    """(rate function, kind): the reference's own source when oracle/_ref travelled with the tree, else the
    oracle port."""
    import oracle
    from oracle import ref
    n = host_thread_count()
    oracle.set_num_threads(n)  # torchrun exports OMP_NUM_THREADS=1; the CPU legs use every host thread
    if ref.available("host"):
        ref.host_set_num_threads(n)
        return cpu_reference_rate, "reference"
    return cpu_oracle_rate, "port"


def run_reference_arm(args):
    """--impl reference: the reference's C++ host code needs OpenMM and ships no CPU platform for this path
    (SURVEY.md 0.1, 8c), but its kernel source is self-contained: the reference arm times that source compiled
    for the host (oracle/_ref, kind "reference") on all host threads, and reports the faster single-pass oracle
    port beside it.  Without oracle/_ref the port is timed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from openmm_constv_b200 import slab as slabmod
    rate_fn, kind = cpu_leg()
    n_gpus = max(1, args.gpus)
    config = workload_config(args.atoms, n_gpus, args.precision, args.slab == "16M")
    total_atoms = config["total_atoms"]
    # the whole job's slab, as the GPU arm partitions it; the CPU walks it in windows
    tiles = [slabmod.slab_tile(total_atoms // 2, n_gpus, t, args.precision, seed=SLAB_SEED) for t in range(n_gpus)]
    s = slabmod.concat_slabs(tiles) if n_gpus > 1 else tiles[0]
    reps = auto_replicas(s.posq.nbytes + s.velm.nbytes + s.force.nbytes + (s.corr.nbytes if s.corr is not None else 0))
    # bound the run: calibrate, then size the per-step window so warmup+steps stay under ~90 s
    rate, _, threads, _, _ = rate_fn(s, 1.0, replicas=reps)
    total = max(1, args.steps + args.warmup)
    window_atoms = min(total_atoms, max(65536, int(rate * 90.0 / total)))
    window = max(1, window_atoms // s.num_cells)
    rate_fn(s, 0, steps_cap=max(1, args.warmup), window=window, replicas=reps)
    rate, done, threads, per_step, el = rate_fn(s, 0, steps_cap=args.steps, window=window, replicas=reps)
    what = ("the reference's constVLangevin.cu compiled for the host (part 1, part 2, image rewrite)" if kind == "reference"
            else "the fused oracle step")
    sample = (f"{done} steps, each {what} over a window of {int(per_step)} atoms walking through "
              f"{reps} replica(s) of the {total_atoms}-atom slab (like the GPU arm's L2 rotation), injected noise pool")
    port = None
    if kind == "reference":
        prate, pdone, _, _, pel = cpu_oracle_rate(s, min(10.0, args.cpu_seconds), window=window, replicas=reps)
        port = {"value": prate, "unit": "atom-steps/s", "kind": "port", "cores": threads,
                "sample": f"{pdone} single-pass oracle steps over the same windows in {pel:.1f} s"}
    line = {
        "impl": "reference", "metric": "atom-steps/s", "value": rate, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / done, "higher_is_better": True, "scaling": "strong" if args.slab == "16M" else "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision == "single" else "f64", "data": "synthetic",
        "config": config,
        "note": ("the reference ships no CPU platform for this path; its kernel source compiled for the host is timed"
                 if kind == "reference" else "oracle/_ref absent: oracle port timed"),
        "cpu_baseline": {"value": rate, "unit": "atom-steps/s", "cores": threads, "kind": kind, "sample": sample, "port": port},
        "e2e": {"value": rate, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- the B200 arm ------------------------------------------------------------------------------------------

class Env:
    """Process-wide handles of one rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from openmm_constv_b200 import _capi, integrator as hostapi, partition, slab as slabmod
        self.torch, self.dist = torch, dist
        self.capi, self.hostapi, self.partition, self.slabmod = _capi, hostapi, partition, slabmod
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the ConstV step path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = _capi.load()
        self.stream = torch.cuda.Stream(device=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())


class Workload:
    """`replicas` rotating groups of `M` slabs on this rank, and how one step of them is launched."""

    def __init__(self, env, args, shard, *, kind="langevin", M=1, replicas=0, chains=1, water=None, drude=None,
                 reorder_perm=None, variant=0, precision=None, per_slab_forces=False):
        torch, lib, capi, hostapi = env.torch, env.lib, env.capi, env.hostapi
        self.env, self.args, self.shard, self.kind, self.M = env, args, shard, kind, M
        self.water, self.drude = water, drude
        self.reordered = reorder_perm is not None
        precision = precision or shard.precision
        rec_bytes = shard.posq.nbytes + shard.velm.nbytes + shard.force.nbytes + (shard.corr.nbytes if shard.corr is not None else 0)
        self.rec_bytes = rec_bytes
        self.replicas = replicas or auto_replicas(rec_bytes * M)
        flags = capi.FLAGS_DEFAULT
        if not args.no_cache_image_charge:
            flags |= capi.FLAG_CACHE_IMAGE_CHARGE
        if args.keep_in_l2:
            flags |= capi.FLAG_KEEP_IN_L2
        self.ctxs = []
        for r in range(self.replicas * M):
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
            ctx = hostapi.SlabContext(sysm, it, precision=precision, device=env.local_rank, flags=flags,
                                      global_atom_offset=shard.global_atom_offset, padded=shard.padded)
            ctx.load_slab(shard)
            if per_slab_forces and (r % M):  # ensemble: every slab sits at its own applied voltage = its own forces
                ctx.force.copy_(torch.from_numpy(env.slabmod.new_forces(shard, r % M)))
            if reorder_perm is not None:
                ctx.set_atom_order(reorder_perm)
                capi.check(lib.constv_update_inverse_order(it._kernel.handle, ctx.stream_ptr()))
            if args.pairs_per_thread or args.ctas_per_sm:
                it._kernel.set_launch_shape(args.pairs_per_thread, args.ctas_per_sm)
            it._kernel.set_kernel_variant(variant, args.tma_block, args.tma_stages, args.tma_ctas)
            if drude:
                capi.check(lib.constv_drude_set_parameters(it._kernel.handle, args.temperature, DRUDE_FRICTION_CM, DRUDE_T,
                                                           DRUDE_FRICTION, DRUDE_WALL, DT))
            else:
                capi.check(lib.constv_set_parameters(it._kernel.handle, args.temperature, FRICTION, DT))
            self.ctxs.append((ctx, it))
        self.handles = [it._kernel.handle for _, it in self.ctxs]
        self.groups = [(C.c_void_p * M)(*self.handles[g * M:(g + 1) * M]) for g in range(self.replicas)]
        self.side_streams = []
        self.set_chains(chains)
        torch.cuda.synchronize()

    def set_chains(self, chains):
        env, lib, capi, M = self.env, self.env.lib, self.env.capi, self.M
        want = max(1, chains)
        if M == 1:
            got = C.c_int32(0)
            for h in self.handles:
                capi.check(lib.constv_set_chains(h, want))
            capi.check(lib.constv_get_chains(self.handles[0], C.byref(got)))
            self.chains = got.value
        else:
            self.chains = min(want, M)
        while len(self.side_streams) < self.chains - 1:
            self.side_streams.append(env.torch.cuda.Stream(device=env.dev))
        # ensemble: sub-group q of every replica's M slabs = one batched launch on stream q
        ch = self.chains
        self.bounds = [M * q // ch for q in range(ch + 1)]
        self.subgroups = [[(C.c_void_p * (self.bounds[q + 1] - self.bounds[q]))(
            *self.handles[g * M + self.bounds[q]:g * M + self.bounds[q + 1]]) for q in range(ch)] for g in range(self.replicas)]

    @property
    def launches_per_step(self):
        return 3 if self.kind == "settle_split" else self.chains

    def launch_steps(self, first, count):
        env, lib, capi = self.env, self.env.lib, self.env.capi
        stream, torch = env.stream, env.torch
        check, kind, M, reps = capi.check, self.kind, self.M, self.replicas
        sptr = stream.cuda_stream
        if self.chains > 1:  # fork: chain q's steps on stream q, in step order; join back into `stream`
            fork = torch.cuda.Event()
            fork.record(stream)
            for q, cs in enumerate([stream] + self.side_streams[:self.chains - 1]):
                if q:
                    cs.wait_event(fork)
                for k in range(first, first + count):
                    if M > 1 and kind == "settle":
                        check(lib.constv_step_fused_settle_batched(self.subgroups[k % reps][q], self.bounds[q + 1] - self.bounds[q], cs.cuda_stream))
                    elif M > 1:
                        check(lib.constv_step_fused_batched(self.subgroups[k % reps][q], self.bounds[q + 1] - self.bounds[q], cs.cuda_stream))
                    elif kind == "drude":
                        check(lib.constv_drude_step_fused_chain(self.handles[k % reps], q, None, 0, cs.cuda_stream))
                    elif kind == "settle":
                        check(lib.constv_step_fused_settle_chain(self.handles[k % reps], q, None, 0, cs.cuda_stream))
                    else:
                        check(lib.constv_step_fused_chain(self.handles[k % reps], q, None, 0, cs.cuda_stream))
            for cs in self.side_streams[:self.chains - 1]:
                join = torch.cuda.Event()
                join.record(cs)
                stream.wait_event(join)
            return
        for k in range(first, first + count):
            if M > 1 and kind == "settle":
                check(lib.constv_step_fused_settle_batched(self.groups[k % reps], M, sptr))
            elif M > 1:
                check(lib.constv_step_fused_batched(self.groups[k % reps], M, sptr))
            elif kind == "drude":
                check(lib.constv_drude_step_fused(self.handles[k % reps], None, 0, sptr))
            elif kind == "settle":
                check(lib.constv_step_fused_settle(self.handles[k % reps], None, 0, sptr))
            elif kind == "settle_split":
                h = self.handles[k % reps]
                check(lib.constv_step_part1(h, None, 0, sptr))
                check(lib.constv_settle_positions(h, sptr))
                check(lib.constv_step_part2_mirror(h, sptr))
            else:
                check(lib.constv_step_fused(self.handles[k % reps], None, 0, sptr))

    def algorithmic_bytes(self):
        sm = self.env.slabmod
        water, drude = self.water, self.drude
        gather = water is not None and self.args.variant == 1  # only the gather kernel reads the item / molecule lists
        total = self.M * sm.algorithmic_bytes_per_step(self.shard, rigid_molecules=len(water[0]) if gather else 0,
                                                       drude=drude is not None, drude_pairs=len(drude[1]) if drude else 0)
        if self.reordered:  # atomIndex of the two slots a thread owns (8 B per real atom), invAtomOrder of each rewritten image (4 B)
            c = self.shard.pair_classes()
            total += 8 * self.shard.num_real + 4 * (c["massive_charged"] + c["massless_charged"]) * (self.shard.num_cells - 1)
        return total

    def reload(self):
        for ctx, _ in self.ctxs:
            ctx.load_slab(self.shard)

    def close(self):
        self.env.torch.cuda.synchronize()
        for ctx, _ in self.ctxs:
            ctx.close()
        self.ctxs = []


def time_steps(env, wl, K, W, use_graph=True, repeats=0, target_s=0.4, sample_clocks=False):
    """W warm-up steps, then the K-step region `repeats` times back to back (see the module docstring).
    Returns {"ms": median region time (max over ranks), "ms_min", "ms_max", "repeats", "eager_launches",
    "graph": description, "clocks": ...}."""
    torch, stream, capi = env.torch, env.stream, env.capi
    reps = wl.replicas
    chunk = reps * max(1, GRAPH_CHUNK_STEPS // reps)
    graphs = {}

    def graph_for(offset, count):
        key = (offset % reps, count)
        if key not in graphs:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                wl.launch_steps(key[0], count)
            graphs[key] = g
        return graphs[key]

    def plan(start, n):
        """The launches of steps [start, start+n): graphs only (eager only with --no-graph)."""
        if not use_graph:
            return [("eager", start, n)]
        out, done = [], 0
        while n - done >= chunk and n > 2 * chunk:
            out.append(graph_for(start + done, chunk))
            done += chunk
        if n - done:
            out.append(graph_for(start + done, n - done))
        return out

    def run(p):
        for g in p:
            if isinstance(g, tuple):
                wl.launch_steps(g[1], g[2])
            else:
                g.replay()

    lps = wl.launches_per_step
    in_graph_error = None
    if use_graph and K * lps <= 400:
        try:
            return time_steps_in_graph(env, wl, K, W, repeats, target_s, sample_clocks)
        except Exception as exc:  # noqa: BLE001 -- e.g. a runtime without external events: stream events between graph launches
            in_graph_error = str(exc)[:200]
            torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        wl.launch_steps(0, reps)  # first touch: refreshes image-charge snapshots outside capture
        stream.synchronize()
        pos = 0
        run(plan(pos, W))
        pos += W
        # one synchronous region: an estimate of its duration (also replays its graphs once)
        first = plan(pos, K)
        stream.synchronize()
        t0 = time.perf_counter()
        run(first)
        stream.synchronize()
        est = max(1e-5, time.perf_counter() - t0)
        pos += K
        n_rep = repeats or int(min(400, max(9, np.ceil(target_s / est))))
        period = reps // int(np.gcd(K, reps))  # distinct graph offsets a sequence of regions walks through
        n_pre = int(max(period, min(200, np.ceil(0.03 / est))))
        plans = [plan(pos + r * K, K) for r in range(n_pre + n_rep)]
        stream.synchronize()
        env.barrier()
        torch.cuda.synchronize()
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(env.local_rank)
            sampler.start()
        launches0 = capi.launch_count()
        events = [torch.cuda.Event(enable_timing=True) for _ in range(n_rep + 1)]
        for r in range(n_pre):  # untimed: every graph replayed once, clocks up, and the GPU busy while the host queues ahead
            run(plans[r])
        events[0].record(stream)
        for r in range(n_rep):
            run(plans[n_pre + r])
            events[r + 1].record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        eager = capi.launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        env.barrier()
    ms = np.asarray([events[r].elapsed_time(events[r + 1]) for r in range(n_rep)])
    med = float(np.median(ms))
    out = {"ms": env.max_over_ranks(med), "ms_min": env.min_over_ranks(float(ms.min())), "ms_max": env.max_over_ranks(float(ms.max())),
           "ms_this_rank": med, "repeats": n_rep, "untimed_regions": n_pre + 1, "eager_launches": int(eager),
           "graph": (f"{len(graphs)} CUDA graphs ({K}-step regions" + (f" as {chunk}-step chunks + tail" if K > 2 * chunk else "") + "), PDL")
                    if use_graph else "eager, PDL",
           "clocks": clocks}
    if in_graph_error:
        out["graph"] += "; in-graph events unavailable: " + in_graph_error
    graphs.clear()
    return out


def time_steps_in_graph(env, wl, K, W, repeats, target_s, sample_clocks):
    """Short regions (K x launches per step <= 400 kernels, e.g. the driver's --steps 20): the timed regions AND the
    event records between them are nodes of one CUDA graph -- [event][K steps][event][K steps] ... -- so that a region
    boundary costs an event node, not the relaunch of a graph (measured: ~7 us per 20-step graph relaunch, 3 % of a
    230 us region).  The event nodes are side branches: with the node IN the chain of regions the next region's kernels
    waited for it, and its ~5 us of latency was charged to every 20-step region (11.53 against 11.26 us per step, session 62;
    at 200-step regions 11.03 against 11.02).  BENCH_EVENT_BRANCH=0 restores the in-line nodes.  Events are created external=True (cudaEventRecordExternal), which is what makes an event recorded
    by a graph node usable with elapsed_time.  The graph holds `per` + 1 regions; the first interval of every replay
    (the GPU was idle before it) is discarded."""
    torch, stream, capi = env.torch, env.stream, env.capi
    lps = wl.launches_per_step
    with torch.cuda.stream(stream):
        wl.launch_steps(0, wl.replicas)  # first touch: refreshes image-charge snapshots outside capture
        stream.synchronize()
        warm = torch.cuda.CUDAGraph()
        with torch.cuda.graph(warm, stream=stream):
            wl.launch_steps(0, W)
        warm.replay()
        stream.synchronize()
        t0 = time.perf_counter()
        warm.replay()
        stream.synchronize()
        est = max(2e-6, (time.perf_counter() - t0) / W) * K  # rough: includes the launch of a small graph
        n_rep = repeats or int(min(400, max(9, np.ceil(target_s / est))))
        per = int(max(2, min(n_rep, 8000 // (K * lps))))
        replays = int(np.ceil(n_rep / per))
        events = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(per + 2)]
        graph = torch.cuda.CUDAGraph()
        # The event record nodes hang off the chain of regions as side branches (recorded on `tick`, which waits for the
        # region's last kernel): an event is stamped when its region completes, and the next region's first kernels depend
        # on that same completion, not on the event node -- so the node's own latency is not charged to the next region.
        side_branch = os.environ.get("BENCH_EVENT_BRANCH", "1") != "0"
        tick = torch.cuda.Stream() if side_branch else None
        with torch.cuda.graph(graph, stream=stream):
            events[0].record(stream)
            for r in range(per + 1):
                wl.launch_steps(W + r * K, K)
                if side_branch:
                    done = torch.cuda.Event()
                    done.record(stream)
                    tick.wait_event(done)
                    events[r + 1].record(tick)
                else:
                    events[r + 1].record(stream)
            if side_branch:
                back = torch.cuda.Event()
                back.record(tick)
                stream.wait_event(back)
        graph.replay()  # untimed: uploads the graph, warms clocks and caches
        stream.synchronize()
        env.barrier()
        torch.cuda.synchronize()
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(env.local_rank)
            sampler.start()
        launches0 = capi.launch_count()
        ms = []
        for _ in range(replays):
            graph.replay()
            stream.synchronize()
            ms.extend(events[r + 1].elapsed_time(events[r + 2]) for r in range(per))
        eager = capi.launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        env.barrier()
    ms = np.asarray(ms)
    med = float(np.median(ms))
    return {"ms": env.max_over_ranks(med), "ms_min": env.min_over_ranks(float(ms.min())), "ms_max": env.max_over_ranks(float(ms.max())),
            "ms_this_rank": med, "repeats": int(len(ms)), "untimed_regions": per + 1 + replays, "eager_launches": int(eager),
            "graph": f"one CUDA graph of {per + 1} regions of {K} steps with an (external) event record node between regions"
                     + (" (a side branch off the region's last kernels: the next region depends on those kernels, not on the event node)" if side_branch else "")
                     + f", replayed {replays}x; the first region of every replay is untimed; PDL",
            "clocks": clocks}


def rate_record(atoms_job, K, t, alg_bytes, launches_per_step, peak, kernel, extra=None):
    """A compact result of one timed leg; alg_bytes = algorithmic bytes per step ON ONE GPU."""
    ms_step = t["ms"] / K
    achieved = alg_bytes / (ms_step * 1e-3) / 1e9
    rec = {"atom_steps_per_s": atoms_job * K / (t["ms"] * 1e-3), "us_per_step": 1e3 * ms_step,
           "GBps_per_gpu": achieved, "frac_measured_peak": achieved / peak, "frac_8TBs": achieved / 8000.0,
           "launches_per_step": launches_per_step, "steps": K, "repeats": t["repeats"], "us_min_max": [1e3 * t["ms_min"] / K, 1e3 * t["ms_max"] / K],
           "eager_launches": t["eager_launches"], "kernel": kernel}
    if extra:
        rec.update(extra)
    return rec


def time_reference_gpu(env, wl, steps, max_blocks, use_graph):
    """The reference's own kernels (constVLangevin.cu compiled by nvcc for sm_100a, oracle/_ref) on the same
    rotating slab replicas: per step the three launches of CudaConstVKernels.cpp:146-166 (part 1 reading a
    pre-filled float4 Gaussian pool, part 2, image rewrite), 128-thread blocks, grid capped at `max_blocks`
    (0 = uncapped).  The pool refill (OpenMM's RNG kernel) and the reordering kernel are NOT charged to the
    reference.  Returns ms per step."""
    import oracle
    from oracle import ref
    torch, stream, shard, ctxs = env.torch, env.stream, wl.shard, wl.ctxs
    dev = env.dev
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
        chunk = n * max(1, min(20, steps // n))
        graph = None
        if use_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                launch(0, chunk)
        reps = max(1, steps // chunk)

        def run():
            for _ in range(reps):
                if graph is not None:
                    graph.replay()
                else:
                    launch(0, chunk)

        run()
        stream.synchronize()
        times = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run()
            e1.record(stream)
            stream.synchronize()
            times.append(e0.elapsed_time(e1) / (reps * chunk))
    return float(np.median(times))


def time_reference_gpu_drude(env, wl, steps):
    import oracle
    from oracle import ref as refmod
    torch, stream, shard, ctxs, drude = env.torch, env.stream, wl.shard, wl.ctxs, wl.drude
    args = wl.args
    R = shard.num_real
    m = np.float32 if args.precision == "single" else np.float64
    scal = oracle.drude_params(args.temperature, DRUDE_FRICTION_CM, DRUDE_T, DRUDE_FRICTION, DRUDE_WALL, DT)
    rr = refmod.CudaDrudeReference(args.precision, scal, m(DT), env.dev, drude[0], drude[1], R, shard.padded, max_blocks=0)
    pool = torch.randn((R + 8, 4), dtype=torch.float32, device=env.dev)
    inv = torch.arange(shard.padded, dtype=torch.int32, device=env.dev)
    sptr = stream.cuda_stream
    n = len(ctxs)

    def ref_launch(first, count):
        for k in range(first, first + count):
            c = ctxs[k % n][0]
            rr.drude_step(shard.num_atoms, shard.num_cells, shard.zmax, c.posq, c.corr, c.velm, c.force, c.pos_delta,
                          pool, 0, inv, stream=sptr)

    with torch.cuda.stream(stream):
        ref_launch(0, n)
        stream.synchronize()
        rchunk = n * max(1, min(20, steps // n))
        rgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(rgraph, stream=stream):
            ref_launch(0, rchunk)
        reps = max(1, steps // rchunk)
        for _ in range(3):
            rgraph.replay()
        stream.synchronize()
        times = []
        for _ in range(5):
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record(stream)
            for _ in range(reps):
                rgraph.replay()
            r1.record(stream)
            stream.synchronize()
            times.append(r0.elapsed_time(r1) / (reps * rchunk))
    return float(np.median(times)), reps * rchunk


def run_e2e(env, wl, args, atoms_job):
    """e2e: host buffers through constv_step_host*.  Per step: H2D of the step's input (the real atoms'
    fixed-point forces, pinned memory), the fused step, the kinetic-energy reduction, and D2H of the step's
    result.  `value` reads back the scalar a StateDataReporter asks for every step (kinetic energy);
    `full_state` also reads back posq of all atoms (what a DCD frame needs)."""
    torch, lib, capi, stream, slabmod = env.torch, env.lib, env.capi, env.stream, env.slabmod
    shard = wl.shard
    R = shard.num_real
    sptr = stream.cuda_stream
    ctx0, it0 = wl.ctxs[0]
    n_e2e = max(3, args.e2e_steps)
    # pinned buffers on the GPU's memory node where the host says which one that is (openmm_constv_b200/hostmem.py;
    # the pool's boxes are single-node VMs: `placement` then records that nothing was done)
    from openmm_constv_b200 import hostmem
    placement = {}
    with hostmem.near_gpu(torch.cuda.current_device(), placement):
        host_force = [torch.from_numpy(np.ascontiguousarray(slabmod.new_forces(shard, j)[:, :R])).pin_memory() for j in range(2)]
        esize = 4 if args.precision != "double" else 8
        host_posq = torch.empty((shard.num_atoms, 4), dtype=torch.float32 if esize == 4 else torch.float64).pin_memory()
        host_posq2 = torch.empty_like(host_posq).pin_memory()
        host_force32 = [(torch.randn((3, R), dtype=torch.float32) * 1000.0).pin_memory() for _ in range(2)]
    h2d = host_force[0].numel() * 8
    ke_host = C.c_double(0.0)
    h = it0._kernel.handle

    def e2e_sync(with_posq):
        out = host_posq.data_ptr() if with_posq else None
        for j in range(3):
            capi.check(lib.constv_step_host(h, host_force[j % 2].data_ptr(), out, None, C.byref(ke_host), sptr))
        best = []
        for _ in range(3):
            env.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for j in range(n_e2e):
                capi.check(lib.constv_step_host(h, host_force[j % 2].data_ptr(), out, None, C.byref(ke_host), sptr))
            e1.record(stream)
            stream.synchronize()
            best.append(e0.elapsed_time(e1))
        return env.max_over_ranks(float(np.median(best)))

    def e2e_pipelined(buffers, code, nsteps, posq_out=None):
        """submit step k+1, then collect step k: the upload of k+1 overlaps the kernels and the read-back of
        k.  Every step still uploads its own forces and returns its own energy (and, with posq_out, its own posq
        of all atoms into one of two alternating host buffers: constv_step_host_submit_state)."""
        ticket = C.c_uint64(0)

        def loop(n):
            prev = None
            for j in range(n):
                if posq_out is None:
                    capi.check(lib.constv_step_host_submit(h, buffers[j % 2].data_ptr(), code, C.byref(ticket), sptr))
                else:
                    capi.check(lib.constv_step_host_submit_state(h, buffers[j % 2].data_ptr(), code, posq_out[j % 2].data_ptr(), None,
                                                                 C.byref(ticket), sptr))
                if prev is not None:
                    capi.check(lib.constv_step_host_collect(h, prev, C.byref(ke_host)))
                prev = ticket.value
            capi.check(lib.constv_step_host_collect(h, prev, C.byref(ke_host)))

        loop(4)
        ts = []
        for _ in range(5):  # PCIe transfers are noisy on a shared host: median of five passes
            env.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            loop(nsteps)
            e1.record(stream)
            stream.synchronize()
            ts.append(e0.elapsed_time(e1))
        return env.max_over_ranks(float(np.median(ts)))

    with torch.cuda.stream(stream):
        ms_ke = e2e_sync(False)
        ms_full = e2e_sync(True)
        n_pipe = 4 * n_e2e
        ms_pipe = e2e_pipelined(host_force, capi.FORCE_FIXED64, n_pipe)
        ms_pipe32 = e2e_pipelined(host_force32, capi.FORCE_FLOAT32, n_pipe)
        n_state = 2 * n_e2e
        ms_state = e2e_pipelined(host_force32, capi.FORCE_FLOAT32, n_state, posq_out=[host_posq, host_posq2])
    return {"value": atoms_job * n_pipe / (ms_pipe32 * 1e-3), "unit": "atom-steps/s",
            "h2d_bytes_per_step": int(h2d // 2), "d2h_bytes_per_step": 8, "steps": n_pipe,
            "ms_per_step": ms_pipe32 / n_pipe, "host_placement": placement,
            "what": "constv_step_host_submit/_collect with HOST buffers, two tickets in flight: every step uploads the real atoms' forces "
                    "from pinned memory (float32 x, y, z planes, 12 B per real atom; converted to OpenMM's fixed point on the device as "
                    "OpenMM's force kernels do), runs the fused step and the kinetic-energy reduction and reads the energy back; "
                    "PCIe-bound; median of 5 passes",
            "fixed64_forces": {"value": atoms_job * n_pipe / (ms_pipe * 1e-3), "ms_per_step": ms_pipe / n_pipe,
                               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
                               "what": "same with the host forces already in OpenMM's int64 fixed-point buffer format (24 B per real atom)"},
            "synchronous": {"value": atoms_job * n_e2e / (ms_ke * 1e-3), "ms_per_step": ms_ke / n_e2e,
                            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8},
            "full_state": {"value": atoms_job * n_e2e / (ms_full * 1e-3), "ms_per_step": ms_full / n_e2e,
                           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(host_posq.numel() * esize + 8),
                           "what": "synchronous constv_step_host, plus D2H of posq for all atoms"},
            "full_state_pipelined": {"value": atoms_job * n_state / (ms_state * 1e-3), "ms_per_step": ms_state / n_state,
                                     "h2d_bytes_per_step": int(h2d // 2), "d2h_bytes_per_step": int(host_posq.numel() * esize + 8),
                                     "what": "constv_step_host_submit_state/_collect: float32 forces up, posq of all atoms and the energy back "
                                             "every step; the read-back of step k (the other PCIe direction) overlaps the upload of step k+1"}}


def run_plugin_path(args, atoms):
    """Steps through the C++ plugin class against the OpenMM shim (plugin/tests/BenchPluginPath.cpp)."""
    exe = os.path.join(ROOT, "openmm_constv_b200", "plugin", "bin", "BenchPluginPath")
    if not os.path.exists(exe):
        return {"unavailable": "plugin/bin/BenchPluginPath has not been built"}
    out = {"what": "us per CudaIntegrateConstVLangevinStepKernel::execute (one launch per step, no graph; rigid_water_* = the same System with "
                   "SPC/E constraints, solved inside the step), a sweep between two steps: "
                   "forces = a kernel copying into the force buffer (24 B/atom), copy = a kernel copying 128 MB -> 128 MB, read = a "
                   "kernel reading 256 MB; bracketed = events around execute(), marginal = loop time with minus without the step"}
    for mode in ("identity", "reordered", "rigid_water_identity", "rigid_water_reordered"):
        cmd = [exe, "--atoms", str(atoms), "--steps", "300", "--precision", args.precision] + (["--reordered"] if mode.endswith("reordered") else []) + \
              (["--rigid-water"] if mode.startswith("rigid_water") else [])
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
            if line:
                r = json.loads(line[-1])
                out[mode] = {"algorithmic_bytes_per_step": r["algorithmic_bytes_per_step"]}
                if "max_constraint_error_nm" in r:  # sanity of the timed steps: a few ulp of the largest coordinate (the injected forces
                    out[mode]["max_constraint_error_nm"] = r["max_constraint_error_nm"]  # are constant: the molecules drift far)
                    out[mode]["max_abs_coordinate_nm"] = r.get("max_abs_coordinate_nm")
                for sweep in ("forces", "copy", "read"):
                    out[mode][sweep] = {"marginal_us": r[sweep]["marginal_us"], "bracketed_us": r[sweep]["bracketed_us"],
                                        "bracketed_frac_8TBs": r[sweep]["bracketed_frac_8TBs"]}
            else:
                out[mode] = {"error": (res.stderr or res.stdout)[-300:]}
        except Exception as exc:  # noqa: BLE001
            out[mode] = {"error": str(exc)[-300:]}
    return out


def run_partition_parity(env, steps=8, tile_real=20000):
    """One small seeded slab, range-partitioned over the ranks (over 2 shards on this GPU when N = 1): after
    `steps` Philox steps every shard's posq / velm / force must equal, bit for bit, the same slots of the
    UNPARTITIONED slab advanced on this GPU, and that must equal the CPU oracle fed the Gaussians the device
    publishes for (seed, step) (rank 0).  The oracle is used as the checker only."""
    torch, lib, capi, hostapi, slabmod, partition = env.torch, env.lib, env.capi, env.hostapi, env.slabmod, env.partition
    parts = max(2, env.world)
    R = tile_real * parts
    tiles = [slabmod.slab_tile(R, parts, t, "single", seed=777) for t in range(parts)]
    full = slabmod.concat_slabs(tiles)
    mine = [env.rank] if env.world > 1 else list(range(parts))

    def advance(s, offset):
        sysm = hostapi.SlabSystem()
        sysm.addParticles(s.masses)
        sysm.setDefaultPeriodicBoxVectors((s.box[0], 0, 0), (0, s.box[1], 0), (0, 0, 2 * s.zmax))
        it = hostapi.ConstVLangevinIntegrator(T_KELVIN, FRICTION, DT)
        it.setRandomNumberSeed(4242)
        ctx = hostapi.SlabContext(sysm, it, precision="single", device=env.local_rank, global_atom_offset=offset, padded=s.padded)
        ctx.load_slab(s)
        return ctx, it

    ok = True
    with torch.cuda.stream(env.stream):
        ctx_full, it_full = advance(full, 0)
        # rank 0: the unpartitioned GPU trajectory against the oracle, step by step
        oracle_ok = None
        if env.rank == 0:
            import oracle
            ref = full.copy()
            prm = oracle.params(T_KELVIN, FRICTION, DT).astype(np.float32)
            inv = np.arange(full.padded, dtype=np.int32)
            delta = np.zeros((full.padded, 4), np.float32)
            noise = torch.zeros((full.padded, 4), dtype=torch.float32, device=env.dev)
            oracle_ok = True
            for k in range(steps):
                capi.check(lib.constv_generate_noise(it_full._kernel.handle, k, noise.data_ptr(), ctx_full.stream_ptr()))
                xi = noise.cpu().numpy()
                it_full.step(1)
                oracle.step("single", full.num_atoms, 2, full.zmax, ref.posq, None, ref.velm, ref.force.reshape(-1),
                            delta, prm, np.float32(DT), xi, 0, inv, True, True)
            oracle_ok = (np.array_equal(ctx_full.posq.cpu().numpy().view(np.uint32), ref.posq.view(np.uint32))
                         and np.array_equal(ctx_full.velm.cpu().numpy().view(np.uint32), ref.velm.view(np.uint32))
                         and np.array_equal(ctx_full.force.cpu().numpy(), ref.force))
            ok = ok and oracle_ok
        else:
            it_full.step(steps)
        Rf = full.num_real
        for t in mine:
            lo, hi = partition.shard_range(Rf, t, parts)
            shard = slabmod.slice_slab(full, lo, hi)
            assert shard.global_atom_offset == tiles[t].global_atom_offset == lo
            assert np.array_equal(shard.posq, tiles[t].posq) and np.array_equal(shard.force, tiles[t].force)
            ctx, it = advance(tiles[t], lo)
            capi.check(lib.constv_set_chains(it._kernel.handle, 2))
            it.step(steps)
            torch.cuda.synchronize()
            r = hi - lo
            for cell in range(2):
                a = slice(cell * r, (cell + 1) * r)
                b = slice(cell * Rf + lo, cell * Rf + hi)
                ok = ok and bool(torch.equal(ctx.posq[a].view(torch.int32), ctx_full.posq[b].view(torch.int32)))
                ok = ok and bool(torch.equal(ctx.velm[a].view(torch.int32), ctx_full.velm[b].view(torch.int32)))
                ok = ok and bool(torch.equal(ctx.force[:, a], ctx_full.force[:, b]))
            ctx.close()
        ctx_full.close()
    all_ok = env.min_over_ranks(1.0 if ok else 0.0) == 1.0
    return {"result": "bit-exact" if all_ok else "MISMATCH", "shards": parts, "ranks": env.world, "steps": steps,
            "atoms": full.num_atoms, "vs": "unpartitioned slab on the GPU == CPU oracle (rank 0); every shard == its slots of the unpartitioned slab"}


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    env = Env()
    torch, lib, capi, slabmod, partition = env.torch, env.lib, env.capi, env.slabmod, env.partition
    world, rank = env.world, env.rank
    peak, peak_src = peaks()
    only = set(filter(None, args.only.split(",")))

    def want(leg):
        return (not args.skip_also) and (not only or leg in only)

    # ---- workload: this rank's shard = its tile of the ONE global slab -------------------------------------
    slab16m = args.slab == "16M"
    config = workload_config(args.atoms, world, args.precision, slab16m)
    R_total = config["total_atoms"] // 2
    R = R_total // world
    plain = not (args.drude or args.rigid_water or args.ensemble or args.reordered)
    water = drude = None
    kind = "langevin"
    if args.drude:
        if args.ensemble or args.rigid_water:
            raise SystemExit("--drude, --rigid-water and --ensemble are separate configurations")
        if args.variant == 1:
            args.chains = 1
        shard, d_normal, d_pairs = slabmod.slab_tile(R_total, world, rank, args.precision, seed=SLAB_SEED, maker=slabmod.make_drude_slab)
        drude = (d_normal, d_pairs)
        kind = "drude"
    elif args.rigid_water:
        if args.ensemble and args.rigid_water != "fused":
            raise SystemExit("--ensemble takes --rigid-water fused only")
        if args.rigid_water == "split" or args.variant == 1:
            args.chains = 1
        shard, w_atoms, w_params, w_cons = slabmod.slab_tile(R_total, world, rank, args.precision, seed=SLAB_SEED, maker=slabmod.make_water_slab)
        water = (w_atoms, w_params, w_cons)
        kind = "settle" if args.rigid_water == "fused" else "settle_split"
    else:
        shard = slabmod.slab_tile(R_total, world, rank, args.precision, seed=SLAB_SEED)
    lo, hi = partition.shard_range(R_total, rank, world)
    assert (lo, hi) == (shard.global_atom_offset, shard.global_atom_offset + shard.num_real), "tile != partition range"
    if args.ensemble:
        shard.global_atom_offset = 0  # independent slabs: each is its own system
    reorder_perm = None
    if args.reordered:
        if args.ensemble:
            raise SystemExit("--reordered and --ensemble are separate configurations")
        args.chains = 1
        per = 4 if drude else 3
        n_mol = shard.meta["n_pairs"] if drude else shard.meta["n_water"] // 3
        reorder_perm = slabmod.molecule_permutation(shard, shard.meta["n_ion"], per, n_mol, seed=7 + rank)
        shard, _ = slabmod.permute_slab(shard, reorder_perm)
    M = max(1, args.ensemble)
    wl = Workload(env, args, shard, kind=kind, M=M, replicas=args.replicas, chains=args.chains, water=water, drude=drude,
                  reorder_perm=reorder_perm, variant=args.variant, per_slab_forces=bool(args.ensemble))
    chains = wl.chains
    n_replicas = wl.replicas
    alg_bytes = wl.algorithmic_bytes()

    # ---- kinetic energy report: the ONE collective of the partition (on demand, never per step), on the
    # initial state (the timed loops keep the injected forces constant, so the state they end in is not physical)
    ke_dev = torch.zeros(1, dtype=torch.float64, device=env.dev)
    wl.ctxs[0][1]._kernel.kinetic_energy_device(0.0, ke_dev)
    torch.cuda.synchronize()
    partition.allreduce_scalar_sum(ke_dev)  # NCCL over NVLink when world > 1
    ke_total = float(ke_dev.item())
    cnt = torch.tensor([float(np.sum(shard.velm[:shard.num_real, 3] != 0))], dtype=torch.float64, device=env.dev)
    partition.allreduce_scalar_sum(cnt)
    temperature = partition.temperature_from_kinetic_energy(ke_total, int(cnt.item()), slabmod.BOLTZ)

    # ---- the headline ------------------------------------------------------------------------------------------
    K, W = args.steps, max(args.warmup, 3)
    use_graph = not args.no_graph
    t = time_steps(env, wl, K, W, use_graph=use_graph, repeats=args.repeats, sample_clocks=True)
    ms_step = t["ms"] / K
    atoms_job = shard.num_atoms * M * world
    value = atoms_job / (ms_step * 1e-3)
    kernel_name = ("constv_settle_step_tiled_batched_kernel" if (args.ensemble and water) else
                   "constv_fused_step_batched_kernel" if args.ensemble else
                   "constv_drude_step_balanced_kernel" if (drude and args.variant != 1) else "constv_drude_step_kernel" if drude else
                   "constv_settle_step_tiled_kernel" if kind == "settle" else
                   "constv_part1_kernel + constv_settle_positions_kernel + constv_part2_mirror_kernel" if kind == "settle_split" else
                   "constv_fused_step_indexed_kernel" if args.reordered else "constv_fused_step_kernel")
    if kind == "settle" and env.capi.last_kernel():
        kernel_name = env.capi.last_kernel()  # the library's choice of stage-in (constv_settle_step_{bulk,tiled,ws}[_batched]_kernel), as launched

    also = {}
    # ---- one launch per step on the same slab (what a host with other kernels between two steps gets at best)
    if plain and chains > 1 and want("single"):
        wl.set_chains(1)
        t1 = time_steps(env, wl, K, W, use_graph=use_graph, repeats=args.repeats)
        also["single_launch"] = rate_record(atoms_job, K, t1, alg_bytes, 1, peak, "constv_fused_step_kernel")
        wl.set_chains(chains)

    # ---- e2e -------------------------------------------------------------------------------------------------------
    e2e = None
    if not args.skip_e2e and plain:
        wl.set_chains(1)
        e2e = run_e2e(env, wl, args, atoms_job)
        wl.set_chains(chains)
        also["e2e_full_state"] = {"atom_steps_per_s": e2e["full_state"]["value"], "ms_per_step": e2e["full_state"]["ms_per_step"],
                                  "d2h_bytes_per_step": e2e["full_state"]["d2h_bytes_per_step"]}

    # ---- the reference's own CUDA kernels on this GPU (information beside the headline) ---------------------
    ref_gpu = None
    if rank == 0 and world == 1 and not args.skip_ref_gpu and not args.ensemble and not args.rigid_water and not args.reordered:
        from oracle import ref as refmod
        if refmod.available("cuda") and not drude:
            wl.reload()
            sms = torch.cuda.get_device_properties(env.dev).multi_processor_count
            n_ref = max(60, min(K, 2400))
            ms_cap = time_reference_gpu(env, wl, n_ref, 6 * sms, use_graph)
            ms_unc = time_reference_gpu(env, wl, n_ref, 0, use_graph)
            best = min(ms_cap, ms_unc)
            ref_gpu = {"value": shard.num_atoms / (best * 1e-3), "unit": "atom-steps/s", "ms_per_step": best,
                       "ms_per_step_grid_6_blocks_per_sm": ms_cap, "ms_per_step_grid_uncapped": ms_unc, "steps": n_ref,
                       "launches_per_step": 3,
                       "what": "the reference's constVLangevin.cu compiled by nvcc for sm_100a (oracle/_ref): part 1 (pre-filled Gaussian "
                               "pool; refill not charged), part 2, image rewrite; 128-thread blocks; same rotating replicas and CUDA-graph "
                               "launch; the faster of OpenMM's capped grid (6 blocks/SM, recalled) and an uncapped grid; median of 5",
                       "speedup_of_fused_step": best / ms_step}
        elif refmod.available("cuda") and hasattr(refmod.cuda(), "constv_ref_drude_cuda_part1_single"):
            wl.reload()
            ms_ref, n_ref = time_reference_gpu_drude(env, wl, max(60, min(K, 2400)))
            ref_gpu = {"value": shard.num_atoms / (ms_ref * 1e-3), "unit": "atom-steps/s", "ms_per_step": ms_ref, "steps": n_ref,
                       "launches_per_step": 4,
                       "what": "the reference's constVDrudeLangevin.cu + constVLangevin.cu compiled by nvcc for sm_100a (oracle/_ref): Drude "
                               "part 1, part 2, hard wall, image rewrite; 128-thread blocks, uncapped grid; same replicas and graph launch",
                       "speedup_of_fused_step": ms_ref / ms_step}

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu and plain:
        rate_fn, ckind = cpu_leg()
        rate, done, threads, per_step, el = rate_fn(shard, args.cpu_seconds, replicas=wl.replicas)
        what = ("steps of the reference's constVLangevin.cu compiled for the host (part 1, part 2, image rewrite; oracle/_ref)"
                if ckind == "reference" else "fused oracle steps")
        cpu = {"value": rate, "unit": "atom-steps/s", "cores": threads, "kind": ckind,
               "sample": f"{done} full {what} rotating over {wl.replicas} replicas of the same {shard.num_atoms}-atom slab in {el:.1f} s, all host threads, injected noise pool",
               "host_cpus": os.cpu_count()}
        if ckind == "reference":
            prate, pdone, _, _, pel = cpu_oracle_rate(shard, args.cpu_seconds / 2, replicas=wl.replicas)
            cpu["port"] = {"value": prate, "unit": "atom-steps/s", "kind": "port", "cores": threads,
                           "sample": f"{pdone} single-pass oracle steps in {pel:.1f} s (a tuned CPU port, for scale)"}
    wl.close()
    del wl
    torch.cuda.empty_cache()

    # ---- behind the plugin class: CudaIntegrateConstVLangevinStepKernel::execute against the shim -----------
    if rank == 0 and world == 1 and plain and want("plugin"):
        also["plugin_path"] = run_plugin_path(args, shard.num_atoms)

    # ---- BASELINE configs[3]: the 16M-atom slab over these N GPUs -----------------------------------------------
    if plain and not slab16m and want("slab16m"):
        Rt = 8_000_000
        tiles16 = 8  # the 16M slab is always the same 8 tiles; a rank of an N-GPU job holds 8/N consecutive ones
        if tiles16 % world == 0:
            mine = [slabmod.slab_tile(Rt, tiles16, tt, args.precision, seed=SLAB_SEED + 100) for tt in
                    range(rank * tiles16 // world, (rank + 1) * tiles16 // world)]
            sh16 = slabmod.concat_slabs(mine) if len(mine) > 1 else mine[0]
            sh16.global_atom_offset = mine[0].global_atom_offset
            assert (sh16.global_atom_offset, sh16.global_atom_offset + sh16.num_real) == partition.shard_range(Rt, rank, world)
            del mine
            w16 = Workload(env, args, sh16, chains=args.chains)
            k16 = max(20, min(K, 600))
            t16 = time_steps(env, w16, k16, 5, use_graph=use_graph, repeats=args.repeats)
            also["slab_16M"] = rate_record(sh16.num_atoms * world, k16, t16, w16.algorithmic_bytes(), w16.chains, peak, "constv_fused_step_kernel",
                                           {"atoms_per_gpu": sh16.num_atoms, "n_gpus": world, "scaling": "strong", "replicas": w16.replicas})
            w16.close()
            del w16, sh16
            torch.cuda.empty_cache()

    # ---- BASELINE configs[4]: 64 independent 100k-atom slabs, batched per GPU -----------------------------------
    if plain and not slab16m and want("ensemble"):
        n_slabs = 64
        if n_slabs % world == 0:
            per_gpu = n_slabs // world
            she = slabmod.make_slab(50_000, args.precision, seed=SLAB_SEED + 200 + rank)
            we = Workload(env, args, she, M=per_gpu, chains=min(4, per_gpu), per_slab_forces=True)
            ke_ = max(20, min(K, 600))
            te = time_steps(env, we, ke_, 5, use_graph=use_graph, repeats=args.repeats)
            rec = rate_record(she.num_atoms * n_slabs, ke_, te, we.algorithmic_bytes(), we.chains, peak, "constv_fused_step_batched_kernel",
                              {"slabs": n_slabs, "slabs_per_gpu": per_gpu, "atoms_per_slab": she.num_atoms, "n_gpus": world, "replicas": we.replicas})
            also["ensemble_64x100k"] = rec
            we.close()
            del we, she
            torch.cuda.empty_cache()

    # ---- the same slab size with the constrained / polarizable steps: one kernel per step each ----
    # (own tile of one global slab per rank, six chains -- profiles/r02_ab_staged_chain_counts.jsonl; the flags --rigid-water /
    # --drude run them as the headline instead)
    if plain and not slab16m and want("rigid"):
        shw, w_atoms, w_params, w_cons = slabmod.slab_tile(R_total, world, rank, args.precision, seed=SLAB_SEED, maker=slabmod.make_water_slab)
        ww = Workload(env, args, shw, kind="settle", chains=6, water=(w_atoms, w_params, w_cons))
        # six chains on six streams: a region boundary joins all of them, so these two legs time regions of at least 120 steps
        # (their `steps` key says so); the headline keeps the driver's K
        kw_ = max(120, min(K, 1200))
        tw = time_steps(env, ww, kw_, 5, use_graph=use_graph, repeats=args.repeats)
        also["rigid_water"] = rate_record(shw.num_atoms * world, kw_, tw, ww.algorithmic_bytes(), ww.chains, peak,
                                          env.capi.last_kernel() or "constv_settle_step_tiled_kernel",
                                          {"atoms_per_gpu": shw.num_atoms, "n_gpus": world, "rigid_molecules_per_gpu": int(len(w_atoms)),
                                           "replicas": ww.replicas, "parity_note": "SETTLE restated from recall: CUDA == oracle bit for bit, "
                                           "oracle unpinned against OpenMM's solver (DESIGN.md 5)"})
        ww.close()
        del ww, shw, w_atoms, w_params, w_cons
        torch.cuda.empty_cache()
    if plain and not slab16m and want("drude"):
        shd, d_normal, d_pairs = slabmod.slab_tile(R_total, world, rank, args.precision, seed=SLAB_SEED, maker=slabmod.make_drude_slab)
        wd = Workload(env, args, shd, kind="drude", chains=6, drude=(d_normal, d_pairs))
        kd_ = max(120, min(K, 1200))
        td = time_steps(env, wd, kd_, 5, use_graph=use_graph, repeats=args.repeats)
        also["drude"] = rate_record(shd.num_atoms * world, kd_, td, wd.algorithmic_bytes(), wd.chains, peak, "constv_drude_step_balanced_kernel",
                                    {"atoms_per_gpu": shd.num_atoms, "n_gpus": world, "pairs_per_gpu": int(len(d_pairs)), "replicas": wd.replicas})
        wd.close()
        del wd, shd, d_normal, d_pairs
        torch.cuda.empty_cache()

    # ---- the partition, checked: shards == unpartitioned == oracle, bit for bit -----------------------------------
    if want("parity"):
        also["partition_parity"] = run_partition_parity(env)

    if rank == 0:
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9  # = (alg_bytes / chains) x (K x chains launches) / time
        line = {
            "metric": "atom-steps/s", "value": value, "unit": "atom-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if slab16m else "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "single" else ("f32+f64" if args.precision == "mixed" else "f64"),
            "data": "synthetic",
            "config": config,
            "variant": {"kind": kind + ("/ensemble" if args.ensemble else "") + ("/reordered" if args.reordered else ""),
                        "slabs_per_gpu": M, "pair_classes": shard.pair_classes(), "rng": "in-kernel Philox4x32-10",
                        "replicas": n_replicas,
                        "launch": t["graph"], "chains": chains,
                        "image_charge": "re-read from posq" if args.no_cache_image_charge else "side array",
                        "kernel_variant": {"variant": args.variant, "tma_block": args.tma_block, "tma_stages": args.tma_stages,
                                           "tma_ctas": args.tma_ctas, "pairs_per_thread": args.pairs_per_thread, "ctas_per_sm": args.ctas_per_sm}},
            "timing": {"repeats": t["repeats"], "statistic": "median of the repeats' K-step region times, max over ranks",
                       "us_per_step_min": 1e3 * t["ms_min"] / K, "us_per_step_max": 1e3 * t["ms_max"] / K,
                       "untimed_regions_before": t["untimed_regions"]},
            "steps_per_s": 1.0 / (ms_step * 1e-3),
            "atom_steps_per_s_per_gpu": value / world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None if args.ensemble else traffic_from_profile(shard.num_atoms, chains), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes / max(1, chains if kind != "settle_split" else 1),
                         "launches_per_step": 3 if kind == "settle_split" else chains,
                         "algorithmic_bytes_per_step": alg_bytes,
                         "note": ("a step is %d launches, one per chain (contiguous range of real atoms) on its own stream; achieved = "
                                  "algorithmic bytes per launch x launches in a region / its duration; per GPU" % chains)
                                 if chains > 1 else "one launch per step; per GPU",
                         "frac_of_8TBs_nominal": achieved / 8000.0, "kernel": kernel_name},
            "cpu_baseline": cpu,
            "reference_gpu": ref_gpu,
            "e2e": e2e,
            "gpu_launches": int(K * (3 if kind == "settle_split" else chains)),
            "gpu_launches_all_repeats": int(K * (3 if kind == "settle_split" else chains) * t["repeats"]),
            "eager_launches_in_region": int(t["eager_launches"]),
            "clocks": t["clocks"],
            "kinetic_energy_kj_mol": ke_total, "temperature_K": temperature,
            "parity_note": ("the constraint solver (SETTLE) is OpenMM's, absent here: restated from the published algorithm, parity "
                            "UNPINNED against OpenMM (DESIGN.md 5); kernel == oracle bit for bit") if water else None,
            "also": also,
        }

        def pick(d, *keys):
            for k in keys:
                d = d.get(k) if isinstance(d, dict) else None
            return d

        # the numbers of the line once more, compact, as its LAST key (a reader of the tail of a long line sees them)
        line["summary"] = {
            "us_per_step": 1e3 * ms_step, "frac_measured_peak": achieved / peak, "atom_steps_per_s": value, "n_gpus": world, "steps": K,
            "eager_launches_in_region": int(t["eager_launches"]), "partition_parity": pick(also, "partition_parity", "result"),
            "e2e_ms_per_step": pick(e2e, "ms_per_step"), "e2e_atom_steps_per_s": pick(e2e, "value"),
            "e2e_full_state_ms_per_step": pick(e2e, "full_state", "ms_per_step"),
            "e2e_full_state_pipelined_ms_per_step": pick(e2e, "full_state_pipelined", "ms_per_step"),
            "cpu_baseline_atom_steps_per_s": pick(cpu, "value"), "reference_gpu_us_per_step": (1e3 * pick(ref_gpu, "ms_per_step")) if pick(ref_gpu, "ms_per_step") else None,
            "single_launch_us": pick(also, "single_launch", "us_per_step"),
            "plugin_path_identity_forces_marginal_us": pick(also, "plugin_path", "identity", "forces", "marginal_us"),
            "plugin_path_reordered_forces_marginal_us": pick(also, "plugin_path", "reordered", "forces", "marginal_us"),
            "slab_16M_us": pick(also, "slab_16M", "us_per_step"), "slab_16M_frac": pick(also, "slab_16M", "frac_measured_peak"),
            "ensemble_64x100k_us": pick(also, "ensemble_64x100k", "us_per_step"), "ensemble_64x100k_frac": pick(also, "ensemble_64x100k", "frac_measured_peak"),
            "rigid_water_us": pick(also, "rigid_water", "us_per_step"), "rigid_water_frac": pick(also, "rigid_water", "frac_measured_peak"),
            "drude_us": pick(also, "drude", "us_per_step"), "drude_frac": pick(also, "drude", "frac_measured_peak"),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
