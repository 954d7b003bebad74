"""Host-side placement for one-process-per-GPU runs: bind a rank to the CPUs (and so, by first touch, the memory) of
its GPU's NUMA node.  On an 8-GPU host the page-locked frame buffers of a rank otherwise land on whatever node the
launcher started it on, and half the ranks copy across the socket interconnect."""
from __future__ import annotations

import ctypes as C
import os


def parse_cpulist(text: str) -> list[int]:
    """'0-31,64-95' -> [0..31, 64..95] (the format of sysfs cpulist files)."""
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(device: int, sysfs: str = "/sys/bus/pci/devices") -> list[int]:
    """CPUs local to CUDA device `device` (empty when the platform does not say)."""
    from . import _lib as L
    buf = C.create_string_buffer(64)
    L.check(L.load().bevk_device_pci_bus_id(int(device), buf, 64))
    bus = buf.value.decode().lower()
    for name in (bus, bus[-12:]):        # "00000000:1b:00.0" vs sysfs "0000:1b:00.0"
        path = os.path.join(sysfs, name, "local_cpulist")
        if os.path.exists(path):
            with open(path) as f:
                return parse_cpulist(f.read())
    return []


def pin_to_gpu(device: int, ranks_on_node: int = 1, slot: int = 0) -> list[int]:
    """Restrict this process to the CPUs local to `device`; with several ranks per NUMA node each takes an equal share
    (slot = its index among them).  Returns the CPU list it ended up with (unchanged affinity if nothing is known)."""
    allowed = sorted(os.sched_getaffinity(0))
    local = [c for c in gpu_local_cpus(device) if c in allowed]
    if not local:
        return allowed
    if ranks_on_node > 1:
        per = max(1, len(local) // ranks_on_node)
        share = local[slot * per:(slot + 1) * per]
        local = share or local
    os.sched_setaffinity(0, local)
    return local
