"""N > 1 on real GPUs: per-tensor granularity on matrices sharded across 2 ranks, range exchanged with ONE
NCCL all-reduce (MAX).  Skipped on boxes with fewer than 2 GPUs (the gloo world-size-2 test in
tests/test_host_logic.py covers the host logic everywhere)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["IWQ_ROOT"])
from iron_weight_only_quant_b200 import shapes, sharded
from oracle import fakequant_oracle as O

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(3)
# OPT-6.7B-like shapes scaled down, every matrix row-split across the ranks (BASELINE.json config 4)
full = [(torch.randn(o, i, generator=g) * 0.02 * (k + 1)).half() for k, (o, i) in enumerate([(512, 512), (2048, 512), (512, 2048), (768, 512)])]
for zero_point in (True, False):
    shards = []
    for w in full:
        r = shapes.split_rows(w.shape[0], world, rank)
        shards.append(w[r.start:r.stop].contiguous().to(dev))
    outs = sharded.pseudo_quantize_tensors_per_tensor_sharded(shards, n_bits=4, zero_point=zero_point)
    torch.cuda.synchronize()
    for w, o in zip(full, outs):
        ref = O.int_fake_quant(w.view(torch.int16).numpy().view(np.uint16), "f16", 4, zero_point, -1)["out"]
        r = shapes.split_rows(w.shape[0], world, rank)
        got = o.cpu().view(torch.int16).numpy().view(np.uint16)
        assert np.array_equal(got, ref[r.start:r.stop]), (zero_point, w.shape)
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_per_tensor_nccl(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, IWQ_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


FUSED_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["IWQ_ROOT"])
from iron_weight_only_quant_b200 import _lib, shapes, sharded
from oracle import fakequant_oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
ok = True
for rep in range(3):  # several calls: the flags carry an epoch, the slots are double-buffered
    fulls, shards = [], []
    for i, (o, k) in enumerate([(512, 1024), (96, 264), (2048, 512), (40, 8), (1024, 4096)]):
        g = torch.Generator().manual_seed(100 * rep + i)
        full = (torch.randn(o, k, generator=g) * 0.02).half()
        r = shapes.split_rows(o, world, rank)
        fulls.append(full)
        shards.append(full[r.start:r.stop].contiguous().to(dev))
    outs = sharded.pseudo_quantize_tensors_per_tensor_fused(shards, n_bits=4, zero_point=(rep != 1))
    torch.cuda.synchronize()
    st = _lib.read_and_clear_status(dev)
    ok = ok and st == 0
    for full, o_ in zip(fulls, outs):
        ref = O.int_fake_quant(full.view(torch.int16).numpy().view(np.uint16), "f16", 4, rep != 1, -1)["out"]
        r = shapes.split_rows(full.shape[0], world, rank)
        ok = ok and bool(np.array_equal(o_.cpu().view(torch.int16).numpy().view(np.uint16), ref[r.start:r.stop]))
t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
if rank == 0:
    print("FUSED_OK" if t.item() == 0 else "FUSED_MISMATCH")
dist.destroy_process_group()
'''


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_per_tensor_peer_exchange(tmp_path):
    """iwq_fake_quant_per_tensor_fused with the matrices row-split over 2 GPUs: the ranges travel through
    peer-mapped (NVLink) stores + system-scope flags inside the kernel (torch symmetric memory supplies the
    mapped buffers), no NCCL collective; result bit-identical to the unsharded oracle."""
    import subprocess
    import sys as _sys

    script = tmp_path / "fused_worker.py"
    script.write_text(FUSED_WORKER)
    env = dict(os.environ, IWQ_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([_sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert "FUSED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
