"""Per-tensor granularity when ONE matrix is split across GPUs (BASELINE.json config 4).

No counterpart in the reference (it never shards a matrix).  Each rank holds
a row block `[rows/N, cols]` of every matrix.  The path is:

    local range pass  ->  ONE all-reduce (MAX) over {max|absmax, -min} of all matrices  ->  local apply pass

min/max are exact and order independent, so every rank derives bit-identical (scale, zero) and the
result equals the unsharded one.  The message is 8 bytes per matrix (OPT-6.7B: 192 matrices = 1.5 KB),
i.e. latency bound; NCCL over NVLink/NVSwitch carries it.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib


def allreduce_ranges(rng: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """MAX all-reduce of the packed ranges `[max_0, -min_0, max_1, -min_1, ...]` in place.

    Works on any backend (NCCL on GPUs; gloo on CPU in the tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(rng, op=dist.ReduceOp.MAX, group=group)
    return rng


@torch.no_grad()
def pseudo_quantize_tensors_per_tensor_sharded(shards: Sequence[torch.Tensor], n_bits: int = 4, zero_point: bool = True, group: Optional[dist.ProcessGroup] = None, inplace: bool = True):
    """INT per-tensor fake-quant of many matrices whose rows are sharded over the ranks of `group`.

    `shards`: this rank's `[rows_local, cols]` block of each matrix (CUDA, fp16/bf16, contiguous, all on
    the same device).  Returns the list of fake-quantised shards (the inputs themselves when inplace).
    """
    if len(shards) == 0:
        return []
    dev = shards[0].device
    dt = shards[0].dtype
    for s in shards:
        if s.device != dev or s.dtype != dt or s.dim() != 2:
            raise ValueError("all shards must be 2-D, on one device, of one dtype")
    outs = list(shards) if inplace else [torch.empty_like(s) for s in shards]
    cfg = _lib.make_config(dt, _lib.SCHEME_INT, _lib.GROUP_PER_TENSOR, bits=n_bits, symmetric=not zero_point)
    mats = [_lib.matrix_desc(s, o, s.shape[0], s.shape[1]) for s, o in zip(shards, outs)]
    rng = torch.empty(2 * len(mats), dtype=torch.float32, device=dev)
    _lib.minmax_(mats, cfg, dev, rng)
    allreduce_ranges(rng, group)
    _lib.fake_quant_with_range_(mats, cfg, dev, rng, track_status=False)
    return outs
