#!/usr/bin/env python
"""Benchmark of the weight fake-quant hot path (BASELINE.json metric: pseudo-quant HBM GB/s and
Llama-2-70B full-model quantise ms).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

Workload (per GPU, weak scaling): the Linear weights of a 1/8 shard of a Llama-2-70B-shaped fp16 model
(10 contiguous decoder layers = 70 matrices = 8.556 G weights = 17.1 GB), random-init on the device,
fake-quantised IN PLACE at INT w4, group 128, asymmetric, emitting scales and zero points -- exactly what
`quantize_model` does per device.  At N = 8 the union of the ranks' shards is the full 80-layer model, so
`ms_per_step` at N = 8 is the "Llama-2-70B full-model quantise ms" of the metric.  One step = one pass over
the shard = ONE persistent kernel launch (work list of 70 matrices).  Inputs (17.1 GB) exceed L2 (126 MB)
by two orders of magnitude, so no L2 flush is needed between steps.  Timing is data independent (no
data-dependent control flow), which is why re-quantising the same buffers every step is legitimate.

`value`  = algorithmic GB/s summed over ranks (4 B per weight + 4 B per group of emitted scale/zero),
           inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`    = the same metric through the host-buffer C-ABI entry point (`iwq_fake_quant_host_batched`): pinned host
           weights -> H2D -> kernel -> D2H, every byte of the shard crossing PCIe both ways each step.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pseudo_quant_hbm_gbps"
UNIT = "GB/s"
BITS, GROUP, SYMMETRIC = 4, 128, False
MODEL = "llama-2-70b"
LAYERS_PER_RANK = 10


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons sampled while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def workload_shapes(rank: int, world: int):
    from iron_weight_only_quant_b200 import shapes

    m = shapes.MODELS[MODEL]
    # each rank owns LAYERS_PER_RANK contiguous decoder layers (rank r of an 8-way balanced partition)
    layers = shapes.partition_layers(m.n_layers, 8, rank % 8)
    assert len(layers) == LAYERS_PER_RANK
    return m.matrices(layers)


def workload_config(mats_shape) -> dict:
    """The `config` object of the JSON line: the SAME for this repo's arm and for the reference arm (which times a
    bounded sample of this workload; the sample is described in its `cpu_baseline.sample`)."""
    n_weights = sum(o * k for o, k in mats_shape)
    return {
        "workload": f"{MODEL}-shaped fp16, {LAYERS_PER_RANK} decoder layers (1/8 of the model: {len(mats_shape)} matrices, {n_weights} weights, {n_weights * 2 / 1e9:.2f} GB) per GPU, INT w{BITS} g{GROUP} asym, in place, scales+zeros emitted; N=8 covers the full 80-layer model",
        "bits": BITS,
        "group": GROUP,
        "symmetric": SYMMETRIC,
        "l2": "inputs (17.1 GB per GPU) exceed L2 (126 MB); no flush needed",
        "launches_per_step": 1,
    }


def algorithmic_bytes(mats) -> int:
    n = sum(o * i for o, i in mats)
    return n * 4 + (n // GROUP) * 4  # 16-bit read + 16-bit write per weight; 16-bit scale + zero per group


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(steps: int, warmup: int, seconds_cap: float = 25.0):
    """Times oracle/torch_cpu_port.int_fake_quant (the reference's ATen op sequence) on [8192, 8192] fp16
    matrices (the 70B q_proj/o_proj shape), w4 g128 asym, with every host thread torch will use."""
    import torch

    from oracle import torch_cpu_port as P

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1234)
    w = (torch.randn(8192, 8192, generator=g) * 0.02).half()
    nbytes = algorithmic_bytes([(8192, 8192)])
    for _ in range(max(1, min(warmup, 2))):
        P.int_fake_quant(w, BITS, not SYMMETRIC, GROUP)
    times = []
    t_all = time.perf_counter()
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        P.int_fake_quant(w, BITS, not SYMMETRIC, GROUP)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > seconds_cap:
            break
    ms = 1e3 * sum(times) / len(times)
    return {
        "value": nbytes / (ms * 1e-3) / 1e9,
        "unit": UNIT,
        "cores": torch.get_num_threads(),
        "kind": "port",
        "sample": f"{len(times)} x one [8192,8192] fp16 matrix (Llama-2-70B q_proj shape, 67.1 M weights), INT w4 g128 asym, torch CPU eager ops in the reference's order (oracle/torch_cpu_port.py)",
        "ms_per_sample": ms,
        "steps_run": len(times),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_sample(args.steps, args.warmup, seconds_cap=120.0)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": cb["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": cb["steps_run"],
        "warmup": args.warmup,
        "ms_per_step": cb["ms_per_sample"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f16",
        "data": "synthetic",
        "config": workload_config(workload_shapes(0, 1)),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from iron_weight_only_quant_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    mats_shape = workload_shapes(rank, world)
    gen = torch.Generator(device=dev)
    weights, scales, zeros, descs = [], [], [], []
    for i, (o, k) in enumerate(mats_shape):
        gen.manual_seed(1234 + i)
        w = torch.empty(o, k, dtype=torch.float16, device=dev)
        # randn * 0.02 generated in chunks to bound fp32 temporaries
        step_rows = max(1, (64 << 20) // k)
        for r0 in range(0, o, step_rows):
            r1 = min(o, r0 + step_rows)
            w[r0:r1] = (torch.randn(r1 - r0, k, generator=gen, device=dev) * 0.02).half()
        G = o * k // GROUP
        s = torch.empty(G, dtype=torch.float16, device=dev)
        z = torch.empty(G, dtype=torch.float16, device=dev)
        weights.append(w)
        scales.append(s)
        zeros.append(z)
        descs.append(_lib.matrix_desc(w, w, o, k, s, z))
    cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, GROUP, bits=BITS, symmetric=SYMMETRIC)
    nbytes = algorithmic_bytes(mats_shape)
    n_weights = sum(o * k for o, k in mats_shape)

    def step():
        _lib.fake_quant_batched_(descs, cfg, dev, track_status=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    status = _lib.read_and_clear_status(dev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = nbytes * world / (ms_per_step * 1e-3) / 1e9

    # ---- end to end: pinned host weights through the host-buffer C-ABI entry point --------------------
    layer_shapes = mats_shape[: len(mats_shape) // LAYERS_PER_RANK]
    # place this rank's pinned buffers on its GPU's NUMA node (first touch after the affinity change)
    from iron_weight_only_quant_b200 import hostmem

    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(visible.split(",")[local]) if visible and visible.split(",")[local].strip().isdigit() else local
    numa_cpus = hostmem.bind_to_gpu_numa(phys)
    host = [torch.empty(o, k, dtype=torch.float16).pin_memory() for o, k in layer_shapes]
    for h, w in zip(host, weights):
        h.copy_(w)  # quantised values: content is irrelevant for timing
    def e2e_step():
        # every layer of the shard crosses PCIe both ways: ONE call, one chunk pipeline over all 70 matrices
        # (the 7 pinned buffers of a layer are passed LAYERS_PER_RANK times; 17.1 GB pinned would be wasteful)
        _lib.fake_quant_host_(host * LAYERS_PER_RANK, cfg, device=local, chunk_bytes=args.e2e_chunk_mb << 20)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    if args.skip_e2e:
        e2e_steps, e2e_s = 0, float("inf")
    else:
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = nbytes * world * e2e_steps / e2e_s / 1e9

    if rank == 0:
        peak, peak_src = _peaks()
        kernel_ms = statistics.mean(per_launch_ms)
        achieved = nbytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("bench_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        cb = cpu_sample(steps=40, warmup=1, seconds_cap=15.0) if (world == 1 and not args.skip_cpu) else None
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f16",
            "data": "synthetic",
            "config": workload_config(mats_shape),
            "pct_of_8TBps": 100.0 * value / (8000.0 * world),
            "weights_per_s": n_weights * world / (ms_per_step * 1e-3),
            "llama2_70b_full_model_ms": ms_per_step * (8.0 / world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "kernel": "seg4_tma_kernel<f16,INT_FAST,asym>", "bytes_per_launch": nbytes, "kernel_ms": kernel_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_weights * 2, "d2h_bytes_per_step": n_weights * 2, "steps": e2e_steps, "api": "iwq_fake_quant_host_batched (pinned host buffers, one 4-deep H2D/kernel/D2H chunk ring over all 70 matrices)", "numa_bound_cpus": len(numa_cpus)},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "nan_status": status,
        }
        if cb is not None:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-chunk-mb", type=int, default=64)
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave out the host-buffer leg")
    ap.add_argument("--skip-cpu", action="store_true", help="profiling runs only: leave out the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
