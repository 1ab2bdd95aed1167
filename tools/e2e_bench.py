#!/usr/bin/env python
"""Host-buffer pipeline (iwq_fake_quant_host) throughput vs chunk size, next to plain pinned H2D+D2H copies."""
import ctypes, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_weight_only_quant_b200 import _lib

rows, cols = 28672, 8192
h = (torch.randn(rows, cols) * 0.02).half().pin_memory()
d = torch.empty(rows, cols, dtype=torch.float16, device="cuda")
cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, 128, bits=4, symmetric=False)
st = ctypes.c_int32(0)
L = _lib.lib()
nb = h.numel() * 2
res = {}
torch.cuda.synchronize()
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); res[name + "_GBps"] = 3 * nb / (time.perf_counter() - t0) / 1e9
for mb in (8, 16, 32, 64, 128, 256):
    L.iwq_fake_quant_host(ctypes.byref(cfg), h.data_ptr(), h.data_ptr(), rows, cols, 0, mb << 20, ctypes.byref(st))
    t0 = time.perf_counter()
    for _ in range(3):
        rc = L.iwq_fake_quant_host(ctypes.byref(cfg), h.data_ptr(), h.data_ptr(), rows, cols, 0, mb << 20, ctypes.byref(st))
        assert rc == 0
    dt = (time.perf_counter() - t0) / 3
    res[f"chunk{mb}MB_alg_GBps"] = nb * 2 / dt / 1e9
print(json.dumps(res))
