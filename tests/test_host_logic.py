"""CPU-side tests: the C-ABI library loads and exports every symbol include/iwq_b200.h declares, the
product path refuses to run without a GPU (no CPU fallback), layer partitioning, and the N>1 range
exchange over gloo (world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge

    ge.build()
    from iron_weight_only_quant_b200 import _lib

    return _lib


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "iwq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(iwq_[a-z_0-9]+)\s*\(", hdr))
    assert {"iwq_fake_quant", "iwq_fake_quant_batched", "iwq_minmax", "iwq_fake_quant_with_range", "iwq_fake_quant_host"} <= declared
    L = ctypes.CDLL(built.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert set(built.EXPORTS) == declared
    assert L.iwq_abi_version() == 4
    L.iwq_workspace_bytes.restype = ctypes.c_size_t
    assert L.iwq_workspace_bytes(10) == 640 + 10 * 256  # job table (64 B each, 128-aligned) + a load/store CUtensorMap pair per matrix


def test_argument_validation_without_gpu(built):
    """Validation happens before any CUDA call, so it is testable here: error codes map to the
    exception classes the reference raises."""
    L = built.lib()
    cfg = built.make_config(torch.float16, built.SCHEME_INT, 128, bits=4)
    m = built.Matrix(w=0, out=0, scales=None, zeros=None, codes=None, rows=4, cols=128)
    rc = L.iwq_fake_quant(ctypes.byref(cfg), ctypes.byref(m), None, None, None)
    with pytest.raises(ValueError):
        built.check(rc)  # null pointers
    m = built.Matrix(w=4096, out=4096, scales=None, zeros=None, codes=None, rows=4, cols=100)
    with pytest.raises(AssertionError):
        built.check(L.iwq_fake_quant(ctypes.byref(cfg), ctypes.byref(m), None, None, None))  # 100 % 128
    cfg = built.make_config(torch.float16, built.SCHEME_INT, -3, bits=4)
    m = built.Matrix(w=4096, out=4096, scales=None, zeros=None, codes=None, rows=4, cols=128)
    with pytest.raises(ValueError):
        built.check(L.iwq_fake_quant(ctypes.byref(cfg), ctypes.byref(m), None, None, None))  # Invalid w_group_size
    cfg = built.make_config(torch.float16, built.SCHEME_FP_CODEC, 128, exp_bits=5, mant_bits=2)
    with pytest.raises(RuntimeError):
        built.check(L.iwq_fake_quant(ctypes.byref(cfg), ctypes.byref(m), None, None, None))  # E5M2 on fp16
    cfg = built.make_config(torch.float16, built.SCHEME_INT, 128, bits=4)
    m = built.Matrix(w=4100, out=4100, scales=None, zeros=None, codes=None, rows=4, cols=128)
    with pytest.raises(RuntimeError):
        built.check(L.iwq_fake_quant(ctypes.byref(cfg), ctypes.byref(m), None, None, None))  # misaligned
    with pytest.raises(TypeError):
        built.make_config(torch.float32, built.SCHEME_INT, 128, bits=4)


def test_no_cpu_fallback(built):
    from iron_weight_only_quant_b200 import fp4_quantize_cpu, quant_funcs, quant_linear

    w = torch.zeros(4, 128, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="CUDA"):
        quant_funcs.pseudo_quantize_tensor(w, n_bits=4, q_group_size=128)
    with pytest.raises(RuntimeError, match="CUDA"):
        fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(w)
    lin = torch.nn.Linear(128, 4, bias=False).half()
    with pytest.raises(RuntimeError, match="CUDA"):
        quant_linear.QuantLinear.from_linear(lin)
    # nothing in the package imports the oracle
    pkg = os.path.join(ROOT, "iron_weight_only_quant_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn


def test_signatures_match_reference(built):
    """Names, positional order and defaults of the reference's public functions."""
    import inspect

    from iron_weight_only_quant_b200 import fp4_quantize_cpu, quant_funcs, quant_linear, quant_wrapper

    def sig(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values()]

    E = inspect.Parameter.empty
    assert sig(quant_funcs.pseudo_quantize_tensor) == [("tensor", E), ("n_bits", 8), ("zero_point", True), ("q_group_size", -1), ("per_tensor", False), ("inplace", False)]
    assert sig(quant_funcs.quantize_weight_per_channel_absmax) == [("w", E), ("n_bits", 8)]
    assert sig(quant_funcs.quantize_activation_per_token_absmax) == [("t", E), ("n_bits", 8)]
    assert sig(fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2) == [("tensor", E), ("group_size", 128), ("per_tensor", False), ("return_scales", False)]
    assert [n for n, _ in sig(fp4_quantize_cpu._fp_scale)] == ["tensor", "S", "M", "bias", "max_float", "min_float"]
    assert sig(quant_wrapper.quantize_model) == [("model", E), ("args", E), ("quant_mix_gate", False)]
    fl = sig(quant_linear.QuantLinear.from_linear)
    assert fl[:8] == [("linear_layer", E), ("w_bit", 4), ("w_group_size", 128), ("symmetric", False), ("mode", 0), ("weight_format", "int"), ("approximate", False), ("quant_dim", 0)]
    assert [n for n, _ in fl[8:]] == ["fp8_hi_align_start", "fp8_hi_align_exp_field", "fp8_tail_pad_bits", "double_approximate", "fp6_hi_align_start", "fp6_hi_align_exp_field", "fp6_tail_pad_bits", "fp4_hi_align_start", "fp4_hi_align_exp_field", "fp4_tail_pad_bits"]
    init = sig(quant_linear.QuantLinear.__init__)
    assert init[1:11] == [("in_features", E), ("out_features", E), ("bias", True), ("w_bit", 4), ("w_group_size", 128), ("symmetric", True), ("mode", 0), ("weight_format", "int"), ("approximate", False), ("quant_dim", 0)]
    ql = quant_linear.QuantLinear(16, 8)
    assert set(dict(ql.named_buffers()).keys()) == {"quantized"}  # the others are registered as None
    for b in ("scales", "zeros", "weight_fp4", "weight_fp6", "weight_fp8", "weight_bfp_mantissa", "weight_bfp_exponent"):
        assert hasattr(ql, b) and getattr(ql, b) is None
    assert ql.weight.shape == (8, 16) and ql.bias.shape == (8,)
    quant_linear.configure_fp_formats(fp4_exp_bits=1, fp4_mantissa_bits=2)
    assert (quant_linear.FP4_EXP_BITS, quant_linear.FP4_MANTISSA_BITS, quant_linear.FP4_EXP_BIAS) == (1, 2, 0)
    quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)
    assert quant_linear.fp_format_of("fp8") == (4, 3)


def test_layer_partition(built):
    from iron_weight_only_quant_b200 import shapes

    m = shapes.MODELS["llama-2-70b"]
    assert m.total_weights() == 68_451_041_280 and len(m.matrices(range(m.n_layers))) == 560
    assert shapes.MODELS["llama-2-7b"].total_weights() == 6_476_005_376
    assert shapes.MODELS["opt-6.7b"].total_weights() == 6_442_450_944
    assert shapes.MODELS["opt-125m"].total_weights() == 84_934_656
    for n, ws in [(80, 8), (80, 3), (32, 5), (12, 8), (3, 8)]:
        blocks = [shapes.partition_layers(n, ws, r) for r in range(ws)]
        assert [x for b in blocks for x in b] == list(range(n))  # contiguous, disjoint, complete
        assert max(len(b) for b in blocks) - min(len(b) for b in blocks) <= 1


_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.environ["IWQ_ROOT"])
from iron_weight_only_quant_b200 import shapes, sharded
from oracle import fakequant_oracle as O

dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['IWQ_PORT']}", rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
g = torch.Generator().manual_seed(7)
mats = [(torch.randn(96, 256, generator=g) * (0.02 * (i + 1))).half() for i in range(5)]
# this rank's row block of every matrix, its local {max, -min}
rng = torch.empty(10, dtype=torch.float32)
for i, w in enumerate(mats):
    rows = shapes.split_rows(96, 2, rank)
    shard = w[rows.start:rows.stop]
    b = shard.contiguous().view(torch.int16).numpy().view(np.uint16)
    mx, mn, _ = O.minmax_of(b, "f16")
    rng[2 * i], rng[2 * i + 1] = float(mx), float(-mn)
sharded.allreduce_ranges(rng)
for i, w in enumerate(mats):
    assert rng[2 * i].item() == w.float().max().item() and -rng[2 * i + 1].item() == w.float().min().item(), i
# every rank now derives the same (scale, zero): quantising its shard with the global range equals the
# matching rows of the unsharded result
for i, w in enumerate(mats):
    full = O.int_fake_quant(w.view(torch.int16).numpy().view(np.uint16), "f16", 4, True, -1)
    s, z = O.int_ranges_to_qparams(np.float32(rng[2 * i].item()), np.float32(-rng[2 * i + 1].item()), 4, True, "f16")
    rows = shapes.split_rows(96, 2, rank)
    shard = O.from_bits(w[rows.start:rows.stop].contiguous().view(torch.int16).numpy().view(np.uint16), "f16")
    out, _ = O.int_apply(shard, s, z, 4, True, "f16")
    assert np.array_equal(O.to_bits(out, "f16"), full["out"][rows.start:rows.stop]), i
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_range_allreduce_world_size_2_gloo(built, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = 29500 + (os.getpid() % 1000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), IWQ_PORT=str(port), IWQ_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_dispatch_plan_for_the_baseline_shapes(built):
    """Which kernel a shape gets (iwq_debug_plan, no CUDA needed): guards the measured routing decisions."""
    L = built.lib()
    L.iwq_debug_plan.restype = ctypes.c_int
    L.iwq_debug_plan.argtypes = [ctypes.POINTER(built.Config), ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int32)]

    def plan(rows, cols, **kw):
        scheme = kw.pop("scheme", built.SCHEME_INT)
        group = kw.pop("group")
        cfg = built.make_config(torch.float16, scheme, group, **kw)
        out = (ctypes.c_int32 * 8)()
        built.check(L.iwq_debug_plan(ctypes.byref(cfg), rows, cols, out))
        return list(out)

    SEG, ROW, COL_ASYNC, COL_REG, COL_2SWEEP, COL_DOUBLE, COL_TMA, COL_WIDE = range(8)
    # per-group: seg4 (variant 0), lanes per segment = group / 32, 4 stages of 48 KB (one CTA per SM)
    for g, lps in [(32, 1), (64, 2), (128, 4), (256, 8), (1024, 32)]:
        p = plan(8192, 8192, group=g, bits=4)
        assert p[0] == SEG and p[1] == 0 and p[2] == lps and p[3] == 4, (g, p)
    assert plan(64, 256, group=16, bits=4)[:2] == [SEG, 2]  # groups below 32 elements: one vector per lane-iteration
    assert plan(4096, 4096, group=-1, bits=4)[:3] == [SEG, 0, 1]  # per-tensor apply pass
    # per-channel: 2048 / 4096-wide rows through seg4 with 2 / 4 warps per row; others through the row kernel
    assert plan(11008, 4096, group=-2, bits=4)[:3] == [SEG, 0, 128]
    assert plan(64, 2048, group=-2, bits=8)[:3] == [SEG, 0, 64]
    p = plan(4096, 11008, group=-2, bits=4)  # 22 KB rows: 3 stages, bulk stores, 4 consumer warps
    assert p[0] == ROW and p[3] == 3 and p[4] == 4 and p[6] == 1, p
    p = plan(8192, 8192, group=-2, bits=4)  # 16 KB rows: row kernel with register stores
    assert p[0] == ROW and p[6] == 0, p
    p = plan(8192, 28672, group=-2, bits=4)  # 56 KB rows: 2 stages only
    assert p[0] == ROW and p[3] == 2 and p[6] == 0, p
    # minifloat rows: 8192-wide through seg4 (8 warps per row), longer ones with many consumer warps, no bulk stores
    assert plan(8192, 8192, group=-2, scheme=built.SCHEME_FP_CODEC, exp_bits=2, mant_bits=1, symmetric=True)[:3] == [SEG, 0, 256]
    p = plan(8192, 28672, group=-2, scheme=built.SCHEME_FP_CODEC, exp_bits=2, mant_bits=1, symmetric=True)
    assert p[0] == ROW and p[4] == 16 and p[6] == 0, p
    # quant_dim = 1
    p = plan(8192, 8192, group=128, bits=4, quant_dim=1)  # CTA-wide ring of 128 x 256 tensor-map tiles (64 KB), three stages
    assert p[0] == COL_WIDE and p[3] == 3 and p[5] == 65536 and p[7] == 256, p
    assert plan(8192, 8192, group=32, bits=4, quant_dim=1)[0] == COL_WIDE and plan(8192, 8192, group=64, bits=4, quant_dim=1)[0] == COL_WIDE
    p = plan(8192, 8192, group=256, bits=4, quant_dim=1)  # warp-private 2-D tensor-map boxes: 256 rows x 32 columns = 16 KB, one stage
    assert p[0] == COL_TMA and p[3] == 1 and p[5] == 16384 and p[7] == 32, p
    assert plan(64, 128, group=32, bits=4, quant_dim=1)[0] == COL_TMA  # smaller than one wide tile
    # minifloats with quant_dim = 1 (what the reference's scripts run): full 128 x 256 tiles always; the direct table only when it
    # fits next to three of them (FP4 on fp16; every format on bf16), else the step table
    fp = dict(scheme=built.SCHEME_FP_CODEC, symmetric=True, quant_dim=1)
    p = plan(8192, 8192, group=128, exp_bits=2, mant_bits=1, **fp)
    assert p[0] == COL_WIDE and p[3] == 3 and p[4] == 128 and p[6] == 1 and p[7] == 256, p
    p = plan(8192, 8192, group=128, exp_bits=4, mant_bits=3, **fp)
    assert p[0] == COL_WIDE and p[3] == 3 and p[4] == 128 and p[6] == 0 and p[7] == 256, p
    p = plan(8192, 8192, group=64, exp_bits=3, mant_bits=2, approx=1, hi_align_start=4, hi_align_exp_field=6, tail_pad_bits=0, **fp)
    assert p[0] == COL_WIDE and p[3] == 3 and p[4] == 128 and p[6] == 0 and p[7] == 256, p
    cfg16 = built.make_config(torch.bfloat16, built.SCHEME_FP_CODEC, 128, exp_bits=4, mant_bits=3, symmetric=True, quant_dim=1)
    out16 = (ctypes.c_int32 * 8)()
    built.check(L.iwq_debug_plan(ctypes.byref(cfg16), 8192, 8192, out16))
    assert out16[0] == COL_WIDE and out16[6] == 1 and out16[7] == 256, list(out16)
    assert plan(8192, 8192, group=512, bits=4, quant_dim=1)[0] == COL_ASYNC  # boxes are at most 256 rows
    assert plan(8192, 8192, group=-2, bits=4, quant_dim=1)[0] == COL_2SWEEP
    assert plan(8192, 8192, group=96, bits=4, quant_dim=1)[0] == COL_2SWEEP
    assert plan(8192, 8192, group=128, scheme=built.SCHEME_FP_CODEC, exp_bits=2, mant_bits=1, symmetric=True, quant_dim=1, approx=2, hi_align_start=0, hi_align_exp_field=2, tail_pad_bits=1)[0] == COL_DOUBLE
    with pytest.raises(AssertionError):
        plan(8192, 8200, group=128, bits=4)  # cols % group != 0 -> the reference's assert


def test_dropin_shims_import_and_propagate_formats(built):
    """with dropin/ in front on sys.path, `import quant_funcs / quant_wrapper / quant_linear /
    fp4_quantize_cpu` (what main.py, eval_quant_lm_eval.py and visualize_utils.py do) binds to this package, and
    configure_fp_formats called on the shim reaches the implementation module the kernels read (quant_linear.py:84-110)."""
    import importlib
    import subprocess
    import sys

    code = r"""
import sys, os
root = sys.argv[1]
sys.path.insert(0, os.path.join(root, "dropin"))
import quant_funcs, quant_wrapper, quant_linear, fp4_quantize_cpu
import iron_weight_only_quant_b200.quant_linear as impl
assert quant_funcs.pseudo_quantize_tensor.__module__ == "iron_weight_only_quant_b200.quant_funcs"
assert quant_wrapper.quantize_model.__module__ == "iron_weight_only_quant_b200.quant_wrapper"
assert quant_linear.QuantLinear is impl.QuantLinear and quant_linear.GPTQQuantLinear is impl.GPTQQuantLinear
assert fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2.__module__ == "iron_weight_only_quant_b200.fp4_quantize_cpu"
for name in ("quantize_weight_per_channel_absmax", "quantize_activation_per_token_absmax", "quantize_weight_per_tensor_absmax", "quantize_activation_per_tensor_absmax"):
    assert hasattr(quant_funcs, name), name
# the names visualize_utils.py:6 imports by value
for name in ("_fp_to_float", "_float_to_fp", "_count_fp4_values", "FP4_EXP_BITS", "FP8_MANTISSA_BITS", "MANTISSA_BITS"):
    assert hasattr(quant_linear, name), name
quant_linear.configure_fp_formats(fp4_exp_bits=1, fp4_mantissa_bits=2, fp6_exp_bits=2, fp6_mantissa_bits=3)
assert (impl.FP4_EXP_BITS, impl.FP4_MANTISSA_BITS, impl.FP4_EXP_BIAS) == (1, 2, 0), "implementation globals"
assert (quant_linear.FP4_EXP_BITS, quant_linear.FP4_MANTISSA_BITS, quant_linear.FP6_EXP_BITS, quant_linear.FP6_MANTISSA_BITS) == (1, 2, 2, 3), "shim globals"
assert impl.fp_format_of("fp4") == (1, 2) and impl.fp_format_of("fp6") == (2, 3)
quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)
assert impl.fp_format_of("fp4") == (2, 1) and quant_linear.FP6_EXP_BITS == 3
print("dropin ok")
"""
    out = subprocess.run([sys.executable, "-c", code, ROOT], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "dropin ok" in out.stdout, out.stderr[-2000:]
