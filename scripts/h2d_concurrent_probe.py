"""Is the host-buffer step's slowdown at 4-8 ranks the host's PCIe / memory system or this repo's submit loop?
Under torchrun: every rank times the same pinned 6 MB host-to-device copy (200 copies enqueued back to back, no host
work between them) first ALONE (the other ranks wait at a barrier), then ALL ranks at once.  Rank 0 prints JSON lines.
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 scripts/h2d_concurrent_probe.py"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from openmm_constv_b200 import hostmem  # noqa: E402


def timed_copies(dst, bufs, stream, n=200):
    with torch.cuda.stream(stream):
        for _ in range(10):
            dst.copy_(bufs[0], non_blocking=True)
        stream.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for j in range(n):
                dst.copy_(bufs[j % 2], non_blocking=True)
            e1.record(stream)
            stream.synchronize()
            ts.append(e0.elapsed_time(e1) / n)
    ts.sort()
    return ts[2]


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = 6_000_000
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    bufs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for b in bufs:
        b.fill_(1)
    stream = torch.cuda.Stream()
    info = hostmem.describe(local)
    info.update({"rank": rank, "cpus": len(os.sched_getaffinity(0))})
    alone = None
    for r in range(world):
        if world > 1:
            dist.barrier()
        if r == rank:
            alone = timed_copies(dst, bufs, stream)
    if world > 1:
        dist.barrier()
    together = timed_copies(dst, bufs, stream)
    rec = {"rank": rank, "alone_ms_per_6MB": alone, "alone_GBps": nbytes / alone / 1e6, "together_ms_per_6MB": together,
           "together_GBps": nbytes / together / 1e6, **info}
    if world > 1:
        out = [None] * world
        dist.all_gather_object(out, rec)
        dist.barrier()
        dist.destroy_process_group()
    else:
        out = [rec]
    if rank == 0:
        for r in out:
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
