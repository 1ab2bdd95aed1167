"""Host-side placement for the host-buffer pipeline (`_lib.fake_quant_host_`): on a multi-socket box the pinned
staging memory of a rank should live on the NUMA node its GPU hangs off, otherwise every H2D / D2H byte crosses
the socket interconnect and the ranks of the far socket share that link.  No arithmetic here."""
from __future__ import annotations

import os


def bind_to_gpu_numa(index: int) -> list[int]:
    """Restrict the calling process to the CPUs NVML reports as local to GPU `index` (first-touch then places
    subsequently pinned host memory on that node).  Returns the CPU list, or [] when NVML / affinity is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [w * 64 + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return []
