#!/usr/bin/env python
"""In-process knob sweep (GPU box): one model's weights are generated once, then every (config, knob set) pair is
timed with CUDA events.  Knobs are set through iwq_debug_set_knob (same names as the IWQ_* environment variables).

    python tools/sweep_r2.py --model opt-6.7b --configs int4pt --knobs 'pt_slab_mb=0;pt_slab_mb=40;pt_slab_mb=40,pt_hint=1'
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_weight_only_quant_b200 import _lib, shapes  # noqa: E402
from tools.microbench import CONFIGS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-2-70b")
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--dtype", default="f16")
    ap.add_argument("--configs", default="int4g128")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--fresh", action="store_true", help="restore pristine weights before every timed pass (data-dependent kernels)")
    ap.add_argument("--knobs", default="", help="';'-separated knob sets, each 'name=value,name=value'")
    ap.add_argument("--shape", default="", help="rows,cols[,count]: `count` matrices of this shape instead of a model's layers")
    ap.add_argument("--shards", type=int, default=0, help="row-split every matrix this many ways and keep shard 0 (config 4 per-rank sizes)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    dt = torch.float16 if args.dtype == "f16" else torch.bfloat16
    m = shapes.MODELS[args.model]
    mats = m.matrices(range(args.layers if args.layers > 0 else m.n_layers))
    if args.shape:
        f = [int(x) for x in args.shape.split(",")]
        mats = [(f[0], f[1])] * (f[2] if len(f) > 2 else 1)
    if args.shards:
        mats = [(o // args.shards, k) for o, k in mats]
    ws = [(torch.randn(o, k, device=dev, dtype=torch.float32) * 0.02).to(dt) for o, k in mats]
    n = sum(w.numel() for w in ws)
    L = _lib.lib()
    defaults = {}

    master = [w.clone() for w in ws] if args.fresh else None

    def timeit(fn):
        if master is not None:
            # FRESH data for every timed pass: the minifloat lookups are data dependent (an already-quantised matrix only
            # holds grid values, which hit a handful of table entries -- no bank conflicts), so re-quantising the same
            # buffer flatters them; the restore is outside the timed region
            tot = 0.0
            for it in range(args.iters + 1):
                for w, m_ in zip(ws, master):
                    w.copy_(m_)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                if it:
                    tot += e0.elapsed_time(e1)
            return tot / args.iters
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    dst = [torch.empty_like(w) for w in ws[: max(1, len(ws) // 4)]]
    ms = timeit(lambda: [d.copy_(w) for d, w in zip(dst, ws)])
    nb = sum(d.numel() for d in dst)
    print(json.dumps({"model": args.model, "weights": n, "torch_copy_GBps": nb * 4 / ms / 1e6}), flush=True)
    del dst
    ksets = [k for k in args.knobs.split(";")] if args.knobs else [""]
    for name in args.configs.split(","):
        c = CONFIGS[name]
        cfg = _lib.make_config(dt, c["scheme"], c["group"], bits=c.get("bits", 0), exp_bits=c.get("exp_bits", 0), mant_bits=c.get("mant_bits", 0), symmetric=c["symmetric"], quant_dim=c.get("quant_dim", 0), approx=c.get("approx", 0), hi_align_start=c.get("hi_align_start", 0), hi_align_exp_field=c.get("hi_align_exp_field", 0), tail_pad_bits=c.get("tail_pad_bits", 0))
        descs, keep = [], []
        for w in ws:
            o, k = w.shape
            G = _lib.num_segments(o, k, c["group"], c.get("quant_dim", 0))
            s = None if c["scheme"] == _lib.SCHEME_BFP else torch.empty(G, dtype=dt, device=dev)
            z = None if c["symmetric"] else torch.empty(G, dtype=dt, device=dev)
            keep.append((s, z))
            descs.append(_lib.matrix_desc(w, w, o, k, s, z))
        rng = torch.empty(2 * len(descs), dtype=torch.float32, device=dev)
        if c["group"] == -1 and c.get("fused"):

            def fn():
                _lib.per_tensor_fused_(descs, cfg, dev, None, track_status=False)
        elif c["group"] == -1:
            # single-GPU per-tensor: the library orders range / apply passes itself (iwq_fake_quant_batched is not for
            # per-tensor; go through minmax + with_range for the classic order, or the per-tensor batched entry)
            def fn():
                _lib.per_tensor_batched_(descs, cfg, dev, rng, track_status=False)
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

        # rounds interleave the knob sets (clock / power drift over a long sweep would otherwise bias the later sets)
        times = {ks: [] for ks in ksets}
        errors = {}
        for rnd in range(args.rounds):
            for ks in ksets:
                if ks in errors:
                    continue
                kv = dict(x.split("=") for x in ks.split(",")) if ks else {}
                for k_, v in defaults.items():  # restore knobs touched by earlier sets
                    L.iwq_debug_set_knob(k_.encode(), v)
                for k_, v in kv.items():
                    defaults.setdefault(k_, int(os.environ.get("IWQ_" + k_.upper(), DEFAULTS.get(k_, 0))))
                    _lib.check(L.iwq_debug_set_knob(k_.encode(), int(v)))
                try:
                    times[ks].append(timeit(fn))
                except Exception as e:  # noqa: BLE001
                    errors[ks] = repr(e)
        for ks in ksets:
            if ks in errors:
                print(json.dumps({"config": name, "knobs": ks, "error": errors[ks]}), flush=True)
            else:
                ts = sorted(times[ks])
                med = ts[len(ts) // 2]
                print(json.dumps({"config": name, "knobs": ks, "ms_median": med, "GBps": n * 4 / med / 1e6, "GBps_best": n * 4 / ts[0] / 1e6, "GBps_worst": n * 4 / ts[-1] / 1e6, "rounds": len(ts), "GBps_all": [round(n * 4 / t / 1e6) for t in times[ks]]}), flush=True)
        for k_, v in defaults.items():
            L.iwq_debug_set_knob(k_.encode(), v)


DEFAULTS = {"fp_rep": 4, "fp_rep_kb": 40, "seg_stages": 3, "pt_reverse": 1, "row_tma_out": 1, "wide_seg4": 1, "fused_batch_mb": 16, "fused_lookahead": 2, "variant": -1}

if __name__ == "__main__":
    main()
