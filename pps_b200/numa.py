"""NUMA placement of the host staging buffers: one process per GPU, each pinned to the CPU cores (and therefore, by
first touch, the memory) of the NUMA node its GPU hangs off.  Eight ranks pulling 190 MB each from one node's memory
is what made the round-1 end-to-end number stop scaling; with local staging every GPU reads through its own root complex.

Pure sysfs + sched_setaffinity (no libnuma in the image).  Everything is best effort: if the node cannot be determined
(no sysfs, numa_node = -1, affinity not permitted) nothing changes and the reason is returned."""
from __future__ import annotations

import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of a CUDA device (through its PCI bus id), or None."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
    except Exception:                                    # noqa: BLE001
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[device_index]) if vis else device_index
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:              # NVML prints an 8-digit domain
                bus = bus[4:]
        except Exception:                                # noqa: BLE001
            return None
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except (OSError, ValueError):
        return None


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the cores of the GPU's NUMA node (call BEFORE allocating pinned host buffers)."""
    info = {"device": device_index, "node": None, "bound": False}
    node = gpu_numa_node(device_index)
    if node is None:
        info["reason"] = "NUMA node of the device unknown"
        return info
    info["node"] = node
    try:
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target:
            info["reason"] = "no allowed core on node %d" % node
            return info
        os.sched_setaffinity(0, target)
        info["bound"], info["cores"] = True, len(target)
    except (OSError, ValueError, AttributeError) as e:
        info["reason"] = "%s: %s" % (type(e).__name__, e)
    return info
