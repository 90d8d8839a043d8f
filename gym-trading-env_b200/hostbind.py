"""Host-side placement for a one-process-per-GPU job: keep each rank's host thread — and, by first touch, the pinned
buffers it allocates afterwards — on the CPU cores / memory node its GPU hangs off.

The "hybrid" path moves every iteration's actions and results through pinned host memory; with eight ranks on a
two-socket host, buffers that landed on the far socket cross the inter-socket link on every copy.  Call
:func:`bind_host_to_gpu` FIRST in the process (before CUDA is initialised and before any pinned allocation).

No third-party code: NVML through ``pynvml`` when importable, else the PCI device's ``numa_node`` in sysfs.  Virtual
machines often expose neither; the call is then a no-op and says so in its return value.
"""
from __future__ import annotations

import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _physical_index(index):
    """CUDA ordinal -> NVML index (CUDA_VISIBLE_DEVICES may renumber the devices)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        items = [v.strip() for v in vis.split(",") if v.strip()]
        if index < len(items) and items[index].isdigit():
            return int(items[index])
    return index


def gpu_cpu_affinity(index):
    """(set of CPU ids local to GPU `index`, NUMA node or None, how it was found)."""
    phys = _physical_index(index)
    bus_id = None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        try:
            bus_id = pynvml.nvmlDeviceGetPciInfo(h).busId
            bus_id = bus_id.decode() if isinstance(bus_id, bytes) else bus_id
        except Exception:  # noqa: BLE001
            pass
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        if cpus and len(cpus) < n_cpu:
            return cpus, _numa_node_of(bus_id), "nvml"
    except Exception:  # noqa: BLE001
        pass
    node = _numa_node_of(bus_id)
    if node is not None and node >= 0:
        try:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
                return _parse_cpulist(fh.read()), node, "sysfs"
        except OSError:
            pass
    return set(), node, "none"


def _numa_node_of(bus_id):
    if not bus_id:
        return None
    # NVML prints an 8-digit PCI domain, sysfs uses 4
    parts = bus_id.lower().split(":")
    if len(parts) == 3 and len(parts[0]) == 8:
        parts[0] = parts[0][4:]
    try:
        with open("/sys/bus/pci/devices/" + ":".join(parts) + "/numa_node") as fh:
            return int(fh.read().strip())
    except (OSError, ValueError):
        return None


def bind_host_to_gpu(index, enable=None):
    """Restrict this process to the CPUs local to GPU `index` (no-op when unknown or when ``GTE_HOST_BIND=0``).
    Returns a small dict describing what was done, for logs / bench lines."""
    if enable is None:
        enable = os.environ.get("GTE_HOST_BIND", "1") != "0"
    info = {"gpu": int(index), "bound": False}
    if not enable:
        info["how"] = "disabled"
        return info
    cpus, node, how = gpu_cpu_affinity(int(index))
    info.update(how=how, numa_node=node)
    try:
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            info.update(bound=True, cpus=len(target))
        else:
            info["cpus"] = len(allowed)
    except (AttributeError, OSError) as e:  # noqa: PERF203
        info["error"] = repr(e)[:80]
    return info
