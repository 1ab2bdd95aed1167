#!/usr/bin/env python
"""Wall-clock of `quant_wrapper.quantize_model` (the reference-facing call: quant_wrapper.py:7-84) on a 1/8 shard of a
Llama-2-70B-shaped fp16 model that already lives on the GPU (10 decoder layers, 70 nn.Linear, 17.1 GB), next to the
GPU time of the one batched launch it issues.  Host overhead = wall - GPU."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_weight_only_quant_b200 import quant_wrapper, shapes  # noqa: E402


class Args:
    w_bit, a_bit, w_group_size, w_symmetric, w_format = 4, 16, 128, False, "int"


def build(weights):
    layers = torch.nn.ModuleList()
    it = iter(weights)
    for _ in range(10):
        blk = torch.nn.Module()
        for name in ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"):
            w = next(it)
            lin = torch.nn.utils.skip_init(torch.nn.Linear, w.shape[1], w.shape[0], bias=False, device="meta")
            lin.weight = torch.nn.Parameter(w, requires_grad=False)
            setattr(blk, name, lin)
        layers.append(blk)
    m = torch.nn.Module()
    m.layers = layers
    return m


def main():
    dev = torch.device("cuda:0")
    mats = shapes.MODELS["llama-2-70b"].matrices(range(10))
    weights = [(torch.randn(o, k, device=dev, dtype=torch.float16) * 0.02) for o, k in mats]
    master = [w.clone() for w in weights]  # pristine copy: the minifloat lookups are data dependent (profiles/README.md), every
    res = {}                               # repetition starts from un-quantised weights, restored outside the timed region
    for fmt, kw in (("int4g128", {}), ("fp4g128", {"w_format": "fp4", "w_symmetric": True}), ("int4pc", {"w_group_size": -2})):
        a = Args()
        for k, v in kw.items():
            setattr(a, k, v)
        walls, gpus = [], []
        for _ in range(4):
            for w, m_ in zip(weights, master):
                w.copy_(m_)
            model = build(weights)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            quant_wrapper.quantize_model(model, a)
            e1.record()
            torch.cuda.synchronize()
            walls.append((time.perf_counter() - t0) * 1e3)
            gpus.append(e0.elapsed_time(e1))
        res[fmt] = {"wall_ms": round(min(walls), 3), "event_ms": round(min(gpus), 3)}
    print(json.dumps({"model": "llama-2-70b 1/8 shard (70 Linear, 17.1 GB) via quantize_model", "results": res}))


if __name__ == "__main__":
    main()
