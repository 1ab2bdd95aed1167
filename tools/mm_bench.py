#!/usr/bin/env python
"""Range pass of the per-tensor path alone (iwq_minmax), back to back, CUDA-event timed, for a list of knob sets.

    python tools/mm_bench.py [--shards 8] --knobs 'mm_tma=0;mm_tma=1;mm_tma=1,mm_stages=3'
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_weight_only_quant_b200 import _lib, shapes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="opt-6.7b")
    ap.add_argument("--shards", type=int, default=0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--knobs", default="")
    ap.add_argument("--one", action="store_true", help="one big matrix of the same total size instead of the model's matrices")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    m = shapes.MODELS[a.model]
    mats = m.matrices(range(m.n_layers))
    if a.shards:
        mats = [(o // a.shards, k) for o, k in mats]
    if a.one:
        tot = sum(o * k for o, k in mats)
        mats = [(tot // 8192, 8192)]
    ws = [(torch.randn(o, k, device=dev, dtype=torch.float32) * 0.02).half() for o, k in mats]
    n = sum(w.numel() for w in ws)
    cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, -1, bits=4, symmetric=False)
    descs = [_lib.matrix_desc(w, w, w.shape[0], w.shape[1]) for w in ws]
    rng = torch.empty(2 * len(descs), dtype=torch.float32, device=dev)
    L = _lib.lib()
    for ks in (a.knobs.split(";") if a.knobs else [""]):
        kv = dict(x.split("=") for x in ks.split(",")) if ks else {}
        for k_, v in {"mm_tma": 1, "mm_stages": 2, "mm_hint": 0}.items():
            L.iwq_debug_set_knob(k_.encode(), v)
        for k_, v in kv.items():
            _lib.check(L.iwq_debug_set_knob(k_.encode(), int(v)))
        for _ in range(2):
            _lib.minmax_(descs, cfg, dev, rng)
        torch.cuda.synchronize()
        torch.cuda._sleep(2_000_000)  # ~1 ms spin: the host runs ahead, so the events see device time only
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            _lib.minmax_(descs, cfg, dev, rng)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        print(json.dumps({"knobs": ks, "matrices": len(descs), "weights": n, "ms": ms, "read_GBps": n * 2 / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
