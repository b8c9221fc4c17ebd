"""Host-side placement for the host-buffer (end-to-end) path: one process per GPU, pinned staging buffers on the
GPU's own NUMA node.

The transform itself never touches host memory; this only matters when volumes arrive in host buffers (the
reference's DataLoader hands `filters_and_operators` CPU tensors, 10_scripts/127_*/...FLAIR.py:187-201) and the
copies over PCIe are the bottleneck.  Pinned pages are placed by first touch, so the calling thread has to run on
the GPU's node *before* it allocates them.
"""
import os
from typing import Optional

import torch

__all__ = ["gpu_numa_node", "bind_to_gpu_numa_node"]


def _sysfs_pci_dir(index: int) -> Optional[str]:
    try:
        pr = torch.cuda.get_device_properties(index)
        name = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except (AttributeError, AssertionError, RuntimeError):
        return None
    path = os.path.join("/sys/bus/pci/devices", name)
    return path if os.path.isdir(path) else None


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index: int) -> Optional[int]:
    """NUMA node the GPU hangs off, or None when sysfs does not say (single-node hosts report -1)."""
    d = _sysfs_pci_dir(index)
    if d is None:
        return None
    try:
        node = int(open(os.path.join(d, "numa_node")).read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_numa_node(index: int) -> dict:
    """Restrict this process to the CPUs local to GPU `index` (within its current affinity mask), so that pinned
    buffers allocated afterwards are first-touched on that node.  Never raises: returns what it did, e.g.
    {"node": 1, "cpus": 56, "bound": True}; {"bound": False, "why": "..."} when the topology is not visible."""
    d = _sysfs_pci_dir(index)
    if d is None:
        return {"bound": False, "why": "no sysfs entry for the device"}
    try:
        local = _parse_cpulist(open(os.path.join(d, "local_cpulist")).read())
    except (OSError, ValueError):
        return {"bound": False, "why": "local_cpulist unreadable"}
    node = gpu_numa_node(index)
    try:
        allowed = os.sched_getaffinity(0)
        target = allowed & local
        if not target:
            return {"bound": False, "node": node, "why": "no local CPU in the affinity mask"}
        if target == allowed:
            return {"bound": False, "node": node, "cpus": len(target), "why": "all allowed CPUs are already local"}
        os.sched_setaffinity(0, target)
    except (AttributeError, OSError) as e:
        return {"bound": False, "node": node, "why": str(e)}
    return {"bound": True, "node": node, "cpus": len(target)}
