"""Placement of pinned host buffers next to the GPU that reads them.

The host-buffer entry points (`constv_step_host*`, include/constv_b200.h) are PCIe-bound: a step uploads the real
atoms' forces (6-12 MB at 1 M atoms) and the kernels hide behind the copy.  On a two-socket host the copy only runs
at the link's speed when the pinned pages sit on the memory node the GPU's PCIe root hangs off; pages on the other
socket cross the inter-socket links, and with one process per GPU the ranks then share those links.  The reference
never meets the question (its state lives on the device and it reads back only reports, `CudaConstVKernels.cpp:139-177`);
a caller that feeds forces from the host does.

Nothing here is needed for correctness.  Every function degrades to a no-op (and says so in what it returns) when
the kernel, the container's seccomp profile or the cpuset refuses: `set_mempolicy` needs CAP_SYS_NICE under Docker's
default profile, so the first-touch route (run the allocating thread on the node's cores) is tried as well.
"""
import contextlib
import ctypes as C
import os
import platform

_MPOL_DEFAULT, _MPOL_PREFERRED = 0, 1
_MPOL_F_NODE, _MPOL_F_ADDR = 1, 2
_SYSCALLS = {"x86_64": (238, 239), "aarch64": (237, 236)}  # (set_mempolicy, get_mempolicy)
_MAXNODE = 1024


def _libc():
    return C.CDLL(None, use_errno=True)


def _syscall_numbers():
    return _SYSCALLS.get(platform.machine())


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of /sys/devices/system/node/nodeK/cpulist)."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_pci_address(device_index):
    """'0000:1b:00.0' of a CUDA device (torch's device properties; None when torch cannot tell)."""
    import torch
    p = torch.cuda.get_device_properties(device_index)
    try:
        return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except AttributeError:
        return None


def pci_numa_node(pci_address, sysfs="/sys/bus/pci/devices"):
    """Memory node of a PCI function, -1 when the platform does not say (single-node hosts, some VMs)."""
    if not pci_address:
        return -1
    try:
        with open(os.path.join(sysfs, pci_address.lower(), "numa_node")) as f:
            return int(f.read().strip())
    except (OSError, ValueError):
        return -1


def node_cpus(node, sysfs="/sys/devices/system/node"):
    try:
        with open(os.path.join(sysfs, "node%d" % node, "cpulist")) as f:
            return parse_cpulist(f.read())
    except OSError:
        return set()


def set_preferred_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) for the calling thread; node None restores MPOL_DEFAULT.
    Raises OSError when the call is refused."""
    nums = _syscall_numbers()
    if nums is None:
        raise OSError("set_mempolicy: unknown syscall number on " + platform.machine())
    libc = _libc()
    if node is None:
        rc = libc.syscall(C.c_long(nums[0]), C.c_int(_MPOL_DEFAULT), C.c_void_p(None), C.c_ulong(0))
    else:
        words = _MAXNODE // 64
        mask = (C.c_ulong * words)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(C.c_long(nums[0]), C.c_int(_MPOL_PREFERRED), mask, C.c_ulong(_MAXNODE))
    if rc != 0:
        err = C.get_errno()
        raise OSError(err, "set_mempolicy: " + os.strerror(err))


def node_of_address(address):
    """Memory node that backs the page at `address` (get_mempolicy with MPOL_F_NODE | MPOL_F_ADDR); -1 if refused."""
    nums = _syscall_numbers()
    if nums is None:
        return -1
    mode = C.c_int(-1)
    rc = _libc().syscall(C.c_long(nums[1]), C.byref(mode), C.c_void_p(None), C.c_ulong(0), C.c_void_p(address),
                         C.c_ulong(_MPOL_F_NODE | _MPOL_F_ADDR))
    return mode.value if rc == 0 else -1


@contextlib.contextmanager
def prefer_node(node):
    """Allocations (and first touches) inside the block prefer memory node `node`; None = leave the policy alone."""
    if node is None or node < 0:
        yield
        return
    set_preferred_node(node)
    try:
        yield
    finally:
        set_preferred_node(None)


@contextlib.contextmanager
def near_gpu(device_index, report=None, pci_address=None):
    """Pinned buffers allocated inside the block land on the GPU's memory node when the host allows it.

    Two routes, both undone on exit: the thread runs on the node's cores (first touch; only the cores the cpuset
    already allows), and the thread's memory policy prefers the node.  `report`, a dict, receives what happened."""
    report = {} if report is None else report
    node = pci_numa_node(pci_address if pci_address is not None else gpu_pci_address(device_index))
    report.update({"gpu_node": node, "cpu_route": False, "policy_route": False})
    if node < 0:
        yield report
        return
    before = os.sched_getaffinity(0)
    local = node_cpus(node) & before
    if local:
        try:
            os.sched_setaffinity(0, local)
            report["cpu_route"] = True
        except OSError as exc:
            report["cpu_route_error"] = str(exc)
    try:
        set_preferred_node(node)
        report["policy_route"] = True
    except OSError as exc:
        report["policy_route_error"] = str(exc)
    try:
        yield report
    finally:
        if report["policy_route"]:
            try:
                set_preferred_node(None)
            except OSError:
                pass
        if report["cpu_route"]:
            os.sched_setaffinity(0, before)


def describe(device_index):
    """What bench.py prints beside the e2e numbers."""
    addr = gpu_pci_address(device_index)
    node = pci_numa_node(addr)
    allowed = os.sched_getaffinity(0)
    return {"gpu_pci": addr, "gpu_node": node, "allowed_cpus": len(allowed),
            "allowed_cpus_on_gpu_node": len(node_cpus(node) & allowed) if node >= 0 else None}
