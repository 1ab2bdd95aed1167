#!/usr/bin/env python
"""Per-batch timeline of pertensor_fused_kernel on CTA 0 (IWQ_FUSED_TRACE=1): when its warp 0 arrived for the batch,
when its exchange lanes had the batch's ranges, when its producer started the apply tiles, when they were done."""
import ctypes, os, sys
import numpy as np
import torch
os.environ["IWQ_FUSED_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iron_weight_only_quant_b200 import _lib, shapes
dev = torch.device("cuda:0")
m = shapes.MODELS["opt-6.7b"]
mats = m.matrices(range(m.n_layers))
shards = [(torch.randn(shapes.split_rows(o, 8, 0).stop - shapes.split_rows(o, 8, 0).start, k, device=dev) * 0.02).half() for o, k in mats]
cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, -1, bits=4, symmetric=False)
descs = [_lib.matrix_desc(s, s, s.shape[0], s.shape[1]) for s in shards]
for _ in range(3):
    _lib.per_tensor_fused_(descs, cfg, dev, None, track_status=False)
torch.cuda.synchronize()
nb = ctypes.c_int(0)
L = _lib.lib()
L.iwq_debug_fused_trace.restype = ctypes.c_longlong
off = L.iwq_debug_fused_trace(ctypes.byref(nb))
ws = _lib._scratch[("fused_ws", 0)]
tr = ws[off : off + nb.value * 32].cpu().numpy().view(np.uint64).reshape(-1, 4).astype(np.int64)
t00 = tr[tr > 0].min()
tr = (tr - t00) / 1000.0
print("batches", nb.value, "total us", tr.max())
print("  b   arrive   gathered  applyStart applyEnd | gather-arrive  start-gather  apply_dur")
for b in list(range(0, 12)) + list(range(nb.value - 4, nb.value)):
    a, g, s, e = tr[b]
    print(f"{b:3d} {a:8.1f} {g:9.1f} {s:10.1f} {e:9.1f} | {g - a:9.1f} {s - g:12.1f} {e - s:10.1f}")
d = tr[:, 1] - tr[:, 0]
print("gather - arrive: mean %.1f us, median %.1f, max %.1f" % (d.mean(), np.median(d), d.max()))
print("apply duration : mean %.1f us" % (tr[:, 3] - tr[:, 2]).mean())
print("batch period   : %.1f us" % ((tr[-1, 3] - tr[0, 3]) / (nb.value - 1)))
