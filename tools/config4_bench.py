#!/usr/bin/env python
"""BASELINE.json config 4: OPT-6.7B-shaped fp16 weights, INT w4 per-tensor (asym), every matrix row-split across
the N ranks; local range pass -> ONE NCCL all-reduce (MAX) of {max, -min} for all 192 matrices -> local apply.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config4_bench.py

Prints one JSON line (rank 0): aggregate algorithmic GB/s (4 B/weight; the two-pass scheme moves 6 B/weight),
per-phase device times (max over ranks), and a bit-exact check of one matrix against the oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_weight_only_quant_b200 import _lib, shapes, sharded  # noqa: E402
from oracle import fakequant_oracle as O  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    m = shapes.MODELS["opt-6.7b"]
    mats = m.matrices(range(m.n_layers))
    gen = torch.Generator(device=dev)
    shards = []
    for i, (o, k) in enumerate(mats):
        # IWQ_FAKE_WORLD=8 on one GPU: time the kernels on rank 0's 1/8 shards (no collective, no oracle check)
        fw = int(os.environ.get("IWQ_FAKE_WORLD", "0"))
        r = shapes.split_rows(o, fw, 0) if fw else shapes.split_rows(o, world, rank)
        gen.manual_seed(1234 + i)  # same full matrix on every rank; keep this rank's rows
        full = (torch.randn(o, k, generator=gen, device=dev) * 0.02).half()
        shards.append(full[r.start : r.stop].contiguous())
        del full
    n_local = sum(s.numel() for s in shards)
    cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, -1, bits=4, symmetric=False)
    descs = [_lib.matrix_desc(s, s, s.shape[0], s.shape[1]) for s in shards]
    rng = torch.empty(2 * len(descs), dtype=torch.float32, device=dev)
    # correctness on matrix 0 before timing (fresh data): all ranks' shards vs the unsharded oracle
    first = shards[0].clone()
    (o0,) = sharded.pseudo_quantize_tensors_per_tensor_sharded([first], n_bits=4, zero_point=True)
    gen.manual_seed(1234)
    full0 = (torch.randn(*mats[0], generator=gen, device=dev) * 0.02).half().cpu()
    ref = O.int_fake_quant(full0.view(torch.int16).numpy().view(np.uint16), "f16", 4, True, -1)["out"]
    r0 = shapes.split_rows(mats[0][0], world, rank)
    ok = bool(np.array_equal(o0.cpu().view(torch.int16).numpy().view(np.uint16), ref[r0.start : r0.stop])) or bool(os.environ.get("IWQ_FAKE_WORLD"))

    fused = bool(os.environ.get("IWQ_C4_FUSED"))

    def step(ev=None):
        if fused:  # one launch: range pass, in-kernel NVLink exchange, apply pass from L2
            if ev:
                ev[0].record()
                ev[1].record()
            peers = None
            if world > 1:
                import ctypes

                st = sharded._peer_buffers(dev, None, len(descs))
                st["epoch"] += 1
                peers = _lib.Peers(world=st["world"], rank=st["rank"], slot_ptrs=ctypes.cast(st["slot_ptrs"], ctypes.POINTER(ctypes.c_void_p)),
                                   flag_ptrs=ctypes.cast(st["flag_ptrs"], ctypes.POINTER(ctypes.c_void_p)), epoch=st["epoch"], spin_limit=0)
            _lib.per_tensor_fused_(descs, cfg, dev, peers, track_status=False)
            return
        _lib.minmax_(descs, cfg, dev, rng)
        if ev:
            ev[0].record()
        sharded.allreduce_ranges(rng)
        if ev:
            ev[1].record()
        _lib.fake_quant_with_range_(descs, cfg, dev, rng, track_status=False)

    for _ in range(3):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    iters = 10
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tot = [0.0, 0.0, 0.0]
    for _ in range(iters):
        e[0].record()
        step(ev=e[1:3])
        e[3].record()
        torch.cuda.synchronize()
        tot[0] += e[0].elapsed_time(e[1])
        tot[1] += e[1].elapsed_time(e[2])
        tot[2] += e[2].elapsed_time(e[3])
    t = torch.tensor([x / iters for x in tot] + [sum(tot) / iters, 0.0 if ok else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t[3].item()
        n_total = sum(o * k for o, k in mats)
        print(json.dumps({"config": "opt-6.7b per-tensor w4 asym, every matrix row-split across ranks" + (" (FUSED single launch)" if fused else ""), "n_gpus": world, "weights": n_total, "ms_total": ms, "ms_range_pass": t[0].item(), "ms_allreduce": t[1].item(), "ms_apply_pass": t[2].item(), "alg_GBps_aggregate": n_total * 4 / ms / 1e6, "traffic_GBps_per_gpu": n_local * 6 / ms / 1e6, "bit_exact_vs_unsharded_oracle": t[4].item() == 0.0}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
