"""Host-side placement for the host<->device log stream.

A full-rate sweep log is tens of GB per call (`cmtjb_run_host` streams it to the host while the
next time chunk integrates), so the host buffer has to sit on the NUMA node the GPU's PCIe root
hangs off: a pinned buffer first touched from the far socket halves the D2H rate.  Pinned memory is
placed where the allocating thread runs, so binding the process to the GPU-local CPUs *before*
allocating is enough.  One process per GPU (torchrun) makes this a per-rank setting."""
import os
from typing import List, Optional


def _parse_cpulist(s: str) -> List[int]:
    cpus: List[int] = []
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_local_cpus(device: int = 0) -> Optional[List[int]]:
    """CPUs local to CUDA device `device` (sysfs local_cpulist of its PCI function), or None."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = torch.cuda.get_device_properties(device).pci_domain_id
        dev = torch.cuda.get_device_properties(device).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            cpus = _parse_cpulist(f.read())
        return cpus or None
    except Exception:
        pass
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        n = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        return cpus or None
    except Exception:
        return None


def bind_host_to_gpu(device: int = 0) -> Optional[List[int]]:
    """Restrict this process to the CPUs local to `device` (intersected with the CPUs it may
    already use).  Returns the CPU list now in effect, or None when nothing was changed."""
    cpus = gpu_local_cpus(device)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return None
    allowed = os.sched_getaffinity(0)
    want = sorted(set(cpus) & allowed)
    if not want or set(want) == allowed:
        return None
    try:
        os.sched_setaffinity(0, want)
    except OSError:
        return None
    return want
