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
