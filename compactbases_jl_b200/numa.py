"""Host-side placement for the host-buffer (`*_host`) entries.

The `_host` entries stream pinned host arrays through PCIe in both directions at once; on a two-socket box every
rank that allocates (first-touches) its pinned buffers on socket 0 shares that socket's memory controllers and its
inter-socket link with all the other ranks.  ``bind_to_gpu_numa`` pins the calling process to the CPUs of the NUMA
node the GPU hangs off *before* the buffers are allocated, so that first-touch puts them next to the GPU's PCIe root
(the Julia glue does the same with ``ThreadPinning`` / ``numactl --cpunodebind``).
Nothing here touches the data path.
"""
from __future__ import annotations

import os


def _parse_cpulist(s: str) -> set[int]:
    cpus: set[int] = set()
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int) -> tuple[int | None, set[int]]:
    """(NUMA node, CPUs of that node) of a CUDA device, from sysfs; (None, {}) when the platform does not tell."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(os.path.join(base, "numa_node")) as f:
            node = int(f.read().strip())
        with open(os.path.join(base, "local_cpulist")) as f:
            cpus = _parse_cpulist(f.read())
        if node < 0:
            return None, cpus
        return node, cpus
    except Exception:
        return None, set()


def bind_to_gpu_numa(device_index: int) -> dict:
    """Restricts the calling process to the CPUs local to the GPU (intersected with the CPUs it may already use).
    Returns what was done; never raises."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        node, cpus = gpu_numa_node(device_index)
        info["numa_node"] = node
        allowed = os.sched_getaffinity(0)
        use = (cpus & allowed) if cpus else set()
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
        info["cpus"] = len(use) if use else len(allowed)
    except Exception as e:                      # containers without sysfs / affinity rights
        info["error"] = str(e)
    return info
