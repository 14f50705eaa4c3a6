"""Host topology probe for the e2e leg: where the GPU hangs, which NUMA nodes and cores this process may use,
and the pinned H2D bandwidth with the buffer bound to each memory node in turn (set_mempolicy through libc).
Usage: python scripts/numa_probe.py  (one GPU; prints JSON lines)."""
import ctypes as C
import glob
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402
from openmm_constv_b200 import hostmem  # noqa: E402


def main():
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    info = hostmem.describe(dev)
    info["online_nodes"] = open("/sys/devices/system/node/online").read().strip() if os.path.exists("/sys/devices/system/node/online") else None
    info["node_cpulists"] = {os.path.basename(p): open(p + "/cpulist").read().strip() for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))}
    for line in open("/proc/self/status"):
        if line.startswith(("Cpus_allowed_list", "Mems_allowed_list")):
            k, v = line.split(":")
            info[k] = v.strip()
    print(json.dumps(info), flush=True)
    nbytes = 6_000_000
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.Stream()
    nodes = sorted(int(os.path.basename(p)[4:]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    for node in [None] + nodes:
        try:
            with hostmem.prefer_node(node):
                bufs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
                for b in bufs:
                    b.fill_(1)
        except OSError as exc:
            print(json.dumps({"node": node, "error": str(exc)}), flush=True)
            continue
        where = hostmem.node_of_address(bufs[0].data_ptr())
        with torch.cuda.stream(stream):
            for _ in range(10):
                dst.copy_(bufs[0], non_blocking=True)
            best = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for j in range(200):
                    dst.copy_(bufs[j % 2], non_blocking=True)
                e1.record(stream)
                stream.synchronize()
                best.append(e0.elapsed_time(e1) / 200)
        best.sort()
        print(json.dumps({"policy_node": node, "buffer_on_node": where, "ms_per_6MB": best[2], "GBps": nbytes / best[2] / 1e6}), flush=True)


if __name__ == "__main__":
    main()
