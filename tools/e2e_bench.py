#!/usr/bin/env python
"""Host-buffer pipeline throughput (algorithmic GB/s = 4 B/weight) vs chunk size for one Llama-2-70B layer
(7 pinned matrices, repeated 5 times): per-matrix calls (iwq_fake_quant_host: the ring drains after every
matrix) vs ONE batched call (iwq_fake_quant_host_batched), next to plain pinned H2D / D2H copies."""
import ctypes, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_weight_only_quant_b200 import _lib, shapes

mats = shapes.MODELS["llama-2-70b"].matrices(range(1))
hs = [(torch.randn(o, k) * 0.02).half().pin_memory() for o, k in mats]
big = max(hs, key=lambda t: t.numel())
d = torch.empty_like(big, device="cuda")
cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, 128, bits=4, symmetric=False)
st = ctypes.c_int32(0)
L = _lib.lib()
REP = 5
nb = sum(h.numel() for h in hs) * 2 * REP
res = {}
torch.cuda.synchronize()
for name, fn in (("h2d", lambda: d.copy_(big, non_blocking=True)), ("d2h", lambda: big.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); res[name + "_GBps"] = 3 * big.numel() * 2 / (time.perf_counter() - t0) / 1e9
for mb in (8, 16, 32, 64, 128):
    def per_matrix():
        for _ in range(REP):
            for h in hs:
                rc = L.iwq_fake_quant_host(ctypes.byref(cfg), h.data_ptr(), h.data_ptr(), h.shape[0], h.shape[1], 0, mb << 20, ctypes.byref(st))
                assert rc == 0
    def batched():
        _lib.fake_quant_host_(hs * REP, cfg, device=0, chunk_bytes=mb << 20)
    for name, fn in (("per_matrix", per_matrix), ("batched", batched)):
        fn()
        t0 = time.perf_counter()
        fn()
        res[f"{name}_chunk{mb}MB_alg_GBps"] = round(nb * 2 / (time.perf_counter() - t0) / 1e9, 1)
print(json.dumps(res))
