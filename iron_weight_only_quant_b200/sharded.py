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


# ------------------------------------------------------------------------------------------------
# fused variant: one launch per rank, ranges exchanged inside the kernel over NVLink peer memory
# ------------------------------------------------------------------------------------------------
_symm_cache = {}


def _peer_buffers(dev: torch.device, group, n: int):
    """Symmetric (peer-mapped) exchange buffers for `n` matrices, cached per (device, group): slots float32
    [2][world][2n] and flags int32 [world][n], zeroed once; returns (_lib.Peers-ready pointer arrays, state)."""
    import ctypes

    import torch.distributed._symmetric_memory as symm_mem

    group = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    key = (dev.index, group.group_name)
    st = _symm_cache.get(key)
    if st is None or st["n"] < n:
        cap = max(n, 256)
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        slots = symm_mem.empty(2 * world * 2 * cap, dtype=torch.float32, device=dev)
        flags = symm_mem.empty(world * cap, dtype=torch.int32, device=dev)
        slots.zero_()
        flags.zero_()
        hs = symm_mem.rendezvous(slots, group)
        hf = symm_mem.rendezvous(flags, group)
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # every rank's buffers are zeroed before anybody's kernel may write into them
        st = dict(n=cap, slots=slots, flags=flags, hs=hs, hf=hf, epoch=0, world=world, rank=rank,
                  slot_ptrs=(ctypes.c_void_p * world)(*[int(x) for x in hs.buffer_ptrs]),
                  flag_ptrs=(ctypes.c_void_p * world)(*[int(x) for x in hf.buffer_ptrs]))
        _symm_cache[key] = st
    return st


@torch.no_grad()
def pseudo_quantize_tensors_per_tensor_fused(shards: Sequence[torch.Tensor], n_bits: int = 4, zero_point: bool = True, group: Optional[dist.ProcessGroup] = None, inplace: bool = True):
    """Same contract as `pseudo_quantize_tensors_per_tensor_sharded`, but ONE kernel launch per rank: the range
    pass, the exchange of {max, -min} between the ranks (peer-mapped stores over NVLink + system-scope flags inside
    the kernel, no NCCL collective) and the apply pass are fused and pipelined over L2-sized batches of matrices.
    Without an initialised process group (or world size 1) it is the single-GPU per-tensor path in one launch."""
    import ctypes

    if len(shards) == 0:
        return []
    dev, dt = shards[0].device, shards[0].dtype
    for s in shards:
        if s.device != dev or s.dtype != dt or s.dim() != 2:
            raise ValueError("all shards must be 2-D, on one device, of one dtype")
    outs = list(shards) if inplace else [torch.empty_like(s) for s in shards]
    cfg = _lib.make_config(dt, _lib.SCHEME_INT, _lib.GROUP_PER_TENSOR, bits=n_bits, symmetric=not zero_point)
    mats = [_lib.matrix_desc(s, o, s.shape[0], s.shape[1]) for s, o in zip(shards, outs)]
    peers = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        st = _peer_buffers(dev, group, len(mats))
        st["epoch"] += 1
        peers = _lib.Peers(world=st["world"], rank=st["rank"], slot_ptrs=ctypes.cast(st["slot_ptrs"], ctypes.POINTER(ctypes.c_void_p)),
                           flag_ptrs=ctypes.cast(st["flag_ptrs"], ctypes.POINTER(ctypes.c_void_p)), epoch=st["epoch"], spin_limit=0)
    _lib.per_tensor_fused_(mats, cfg, dev, peers, track_status=True)
    # the exchange waits are bounded: a peer that never arrives raises IWQ_STATUS_TIMEOUT and the kernel goes on with
    # whatever ranges it has -- never hand such weights back silently.  (One sync per call; this entry point is opt-in.)
    st = _lib.read_and_clear_status(dev)
    if st & _lib.STATUS_TIMEOUT:
        raise RuntimeError("pseudo_quantize_tensors_per_tensor_fused: a peer's ranges did not arrive (IWQ_STATUS_TIMEOUT); the outputs are invalid")
    assert st & _lib.STATUS_NAN == 0  # quant_funcs.py:40
    return outs
