#!/usr/bin/env python
"""Kernel micro-benchmark (GPU box): algorithmic GB/s of the fused fake-quant pass on Llama-2-70B layer
shapes for a set of configurations, next to a plain torch copy of the same bytes.  Knobs come from the
environment (IWQ_VARIANT, IWQ_SEG_STAGES, IWQ_SEG_CTAS, IWQ_ROW_CTAS, IWQ_ROW_SMEM_KB), read once per
process, so `tools/sweep.sh` runs this script once per setting.

    python tools/microbench.py [--layers 2] [--iters 10] [--configs int4g128,int4pc,...]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_weight_only_quant_b200 import _lib, shapes  # noqa: E402

CONFIGS = {
    "int4g128": dict(scheme=_lib.SCHEME_INT, group=128, bits=4, symmetric=False),
    "int4g128sym": dict(scheme=_lib.SCHEME_INT, group=128, bits=4, symmetric=True),
    "int4g32": dict(scheme=_lib.SCHEME_INT, group=32, bits=4, symmetric=False),
    "int4g64": dict(scheme=_lib.SCHEME_INT, group=64, bits=4, symmetric=False),
    "int4g256": dict(scheme=_lib.SCHEME_INT, group=256, bits=4, symmetric=False),
    "int8pc": dict(scheme=_lib.SCHEME_INT, group=-2, bits=8, symmetric=False),
    "int4pc": dict(scheme=_lib.SCHEME_INT, group=-2, bits=4, symmetric=False),
    "int4pt": dict(scheme=_lib.SCHEME_INT, group=-1, bits=4, symmetric=False),
    "fp4e2m1g128": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=2, mant_bits=1, symmetric=True),
    "fp8e4m3g128": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=4, mant_bits=3, symmetric=True),
    "fp4e2m1g128asym": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=2, mant_bits=1, symmetric=False),
    "fp4e1m2g128": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=1, mant_bits=2, symmetric=True),
    "fp6e3m2g128": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=3, mant_bits=2, symmetric=True),
    "fp6e2m3g128": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=2, mant_bits=3, symmetric=True),
    "fprne_e2m1g128": dict(scheme=_lib.SCHEME_FP_RNE, group=128, exp_bits=2, mant_bits=1, symmetric=True),
    "bfp4g128": dict(scheme=_lib.SCHEME_BFP, group=128, bits=4, symmetric=True),
    "bfp6g64": dict(scheme=_lib.SCHEME_BFP, group=64, bits=6, symmetric=True),
    "int4g128_d1": dict(scheme=_lib.SCHEME_INT, group=128, bits=4, symmetric=False, quant_dim=1),
    "int4g128sym_d1": dict(scheme=_lib.SCHEME_INT, group=128, bits=4, symmetric=True, quant_dim=1),
    "int4g64_d1": dict(scheme=_lib.SCHEME_INT, group=64, bits=4, symmetric=False, quant_dim=1),
    "int4g32_d1": dict(scheme=_lib.SCHEME_INT, group=32, bits=4, symmetric=False, quant_dim=1),
    "bfp4g128_d1": dict(scheme=_lib.SCHEME_BFP, group=128, bits=4, symmetric=True, quant_dim=1),
    "int4g256_d1": dict(scheme=_lib.SCHEME_INT, group=256, bits=4, symmetric=False, quant_dim=1),
    "fp4e2m1g128_d1": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=2, mant_bits=1, symmetric=True, quant_dim=1),
    "fp4e2m1g64_d1": dict(scheme=_lib.SCHEME_FP_CODEC, group=64, exp_bits=2, mant_bits=1, symmetric=True, quant_dim=1),
    "fp8e4m3g128_d1": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=4, mant_bits=3, symmetric=True, quant_dim=1),
    "fp6e3m2g128_d1": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=3, mant_bits=2, symmetric=True, quant_dim=1),
    "fp8e4m3g64_d1": dict(scheme=_lib.SCHEME_FP_CODEC, group=64, exp_bits=4, mant_bits=3, symmetric=True, quant_dim=1),
    "fp6e3m2g64_d1": dict(scheme=_lib.SCHEME_FP_CODEC, group=64, exp_bits=3, mant_bits=2, symmetric=True, quant_dim=1),
    "fp4e2m1g128_d1_approx": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=2, mant_bits=1, symmetric=True, quant_dim=1, approx=1, hi_align_start=1, hi_align_exp_field=1, tail_pad_bits=0),
    "fp4e2m1g128_d1_double": dict(scheme=_lib.SCHEME_FP_CODEC, group=128, exp_bits=2, mant_bits=1, symmetric=True, quant_dim=1, approx=2, hi_align_start=0, hi_align_exp_field=2, tail_pad_bits=1),
    "fp8e4m3g64_d1_double": dict(scheme=_lib.SCHEME_FP_CODEC, group=64, exp_bits=4, mant_bits=3, symmetric=True, quant_dim=1, approx=2, hi_align_start=13, hi_align_exp_field=15, tail_pad_bits=-1),
    "int4pc_sym": dict(scheme=_lib.SCHEME_INT, group=-2, bits=4, symmetric=True),
    "fp4e2m1pc": dict(scheme=_lib.SCHEME_FP_CODEC, group=-2, exp_bits=2, mant_bits=1, symmetric=True),
    "int4pt_sym": dict(scheme=_lib.SCHEME_INT, group=-1, bits=4, symmetric=True),
    "int4pt_fused": dict(scheme=_lib.SCHEME_INT, group=-1, bits=4, symmetric=False, fused=True),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-2-70b")
    ap.add_argument("--layers", type=int, default=2, help="decoder layers of the model to use (0 = all)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dtype", default="f16")
    ap.add_argument("--configs", default="int4g128,int4g128sym,int4g32,int8pc,int4pt,fp4e2m1g128,fp8e4m3g128")
    ap.add_argument("--no-emit", action="store_true", help="do not emit scales/zeros")
    ap.add_argument("--fresh", action="store_true", help="no warm-up, one pass per configuration on pristine weights (profiling the data-dependent minifloat kernels)")
    ap.add_argument("--shape", default="", help="rows,cols[,count]: use `count` matrices of this shape instead of a model's layers")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.float16 if args.dtype == "f16" else torch.bfloat16
    m = shapes.MODELS[args.model]
    mats = m.matrices(range(args.layers if args.layers > 0 else m.n_layers))
    if args.shape:
        f = [int(x) for x in args.shape.split(",")]
        mats = [(f[0], f[1])] * (f[2] if len(f) > 2 else 1)
    ws = [(torch.randn(o, k, device=dev, dtype=torch.float32) * 0.02).to(dt) for o, k in mats]
    n = sum(w.numel() for w in ws)
    knobs = {k: os.environ[k] for k in os.environ if k.startswith("IWQ_")}
    res = {"knobs": knobs, "model": args.shape or args.model, "layers": args.layers if args.layers > 0 else m.n_layers, "dtype": args.dtype, "weights": n, "results": {}}

    master = [w.clone() for w in ws] if args.fresh else None

    def timeit(fn):
        if master is not None:
            for w, m_ in zip(ws, master):
                w.copy_(m_)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    # copy ceiling: same bytes read + written
    if not args.fresh:
        dst = [torch.empty_like(w) for w in ws]
        ms = timeit(lambda: [d.copy_(w) for d, w in zip(dst, ws)])
        res["results"]["torch_copy"] = {"ms": ms, "GBps": n * 4 / ms / 1e6}
        del dst
    for name in args.configs.split(","):
        c = CONFIGS[name]
        cfg = _lib.make_config(dt, c["scheme"], c["group"], bits=c.get("bits", 0), exp_bits=c.get("exp_bits", 0), mant_bits=c.get("mant_bits", 0), symmetric=c["symmetric"], quant_dim=c.get("quant_dim", 0), approx=c.get("approx", 0), hi_align_start=c.get("hi_align_start", 0), hi_align_exp_field=c.get("hi_align_exp_field", 0), tail_pad_bits=c.get("tail_pad_bits", 0))
        descs, keep = [], []
        for w in ws:
            o, k = w.shape
            G = _lib.num_segments(o, k, c["group"], c.get("quant_dim", 0))
            s = None if (args.no_emit or c["scheme"] == _lib.SCHEME_BFP) else torch.empty(G, dtype=dt, device=dev)
            z = None if (args.no_emit or c["symmetric"]) else torch.empty(G, dtype=dt, device=dev)
            keep.append((s, z))
            descs.append(_lib.matrix_desc(w, w, o, k, s, z))
        if c["group"] == -1 and c.get("fused"):

            def fn():
                _lib.per_tensor_fused_(descs, cfg, dev, None, track_status=False)
        elif c["group"] == -1:
            rng = torch.empty(2 * len(descs), dtype=torch.float32, device=dev)

            def fn():
                _lib.minmax_(descs, cfg, dev, rng)
                _lib.fake_quant_with_range_(descs, cfg, dev, rng, track_status=False)
        elif c["group"] == -2 and c.get("quant_dim", 0) == 0:
            by_cols = {}
            for d, w in zip(descs, ws):
                by_cols.setdefault(w.shape[1], []).append(d)

            def fn():
                for ds in by_cols.values():
                    _lib.fake_quant_batched_(ds, cfg, dev, track_status=False)
        else:

            def fn():
                _lib.fake_quant_batched_(descs, cfg, dev, track_status=False)

        try:
            ms = timeit(fn)
            res["results"][name] = {"ms": ms, "GBps": n * 4 / ms / 1e6}
        except Exception as e:  # noqa: BLE001
            res["results"][name] = {"error": repr(e)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
