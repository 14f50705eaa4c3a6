"""Host-memory placement helpers (openmm_constv_b200/hostmem.py): parsing and the no-op fallbacks; no GPU needed."""
import os

from openmm_constv_b200 import hostmem


def test_parse_cpulist():
    assert hostmem.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem.parse_cpulist("") == set()
    assert hostmem.parse_cpulist("5") == {5}


def test_pci_numa_node_reads_sysfs(tmp_path):
    dev = tmp_path / "0000:1b:00.0"
    dev.mkdir()
    (dev / "numa_node").write_text("1\n")
    assert hostmem.pci_numa_node("0000:1B:00.0", sysfs=str(tmp_path)) == 1
    assert hostmem.pci_numa_node("0000:aa:00.0", sysfs=str(tmp_path)) == -1
    assert hostmem.pci_numa_node(None) == -1


def test_node_cpus(tmp_path):
    node = tmp_path / "node1"
    node.mkdir()
    (node / "cpulist").write_text("16-31\n")
    assert hostmem.node_cpus(1, sysfs=str(tmp_path)) == set(range(16, 32))
    assert hostmem.node_cpus(7, sysfs=str(tmp_path)) == set()


def test_near_gpu_without_a_node_is_a_no_op():
    before = os.sched_getaffinity(0)
    report = {}
    with hostmem.near_gpu(0, report, pci_address="ffff:ff:1f.0") as r:  # no such function: node -1
        assert r is report
        assert os.sched_getaffinity(0) == before
    assert report == {"gpu_node": -1, "cpu_route": False, "policy_route": False}
    assert os.sched_getaffinity(0) == before


def test_prefer_node_places_and_restores():
    # node 0 exists on every Linux host; a refusal (seccomp) must surface as OSError, not as a crash
    import numpy as np
    try:
        with hostmem.prefer_node(0):
            buf = np.ones(1 << 16)
    except OSError:
        return
    assert hostmem.node_of_address(buf.ctypes.data) in (-1, 0)
    with hostmem.prefer_node(None):  # "leave the policy alone"
        pass
    with hostmem.prefer_node(-1):
        pass
