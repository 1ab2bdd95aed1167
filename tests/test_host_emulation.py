"""CPU check of the product's device math (host instantiation of csrc/iwq_math.cuh + iwq_consts.h).

The build box has no GPU, so before GPU time is spent the same `__host__ __device__` functions the CUDA
kernels call are compiled for the host into a TEST-ONLY shared library (tests/host/emulate.cu) and run
against the reference's golden vectors.  This is not a fallback: the package never loads it.
"""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "emulate.cu")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

F16, BF16 = 0, 1
SCHEME_INT, SCHEME_FP_CODEC, SCHEME_FP_RNE, SCHEME_BFP = 0, 1, 2, 3
KIND_INT_FAST, KIND_INT_WIDE, KIND_FP, KIND_INT_GENERIC = 0, 1, 2, 3
FORMATS = {"e2m1": (2, 1), "e1m2": (1, 2), "e3m2": (3, 2), "e2m3": (2, 3), "e4m3": (4, 3), "e5m2": (5, 2)}


class Cfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("dtype", "scheme", "bits", "exp_bits", "mant_bits", "symmetric", "group", "variant", "quant_dim", "approx", "hi_align_start", "hi_align_exp_field", "tail_pad_bits", "code_bits")]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("emu") / "libiwq_emulate.so")
    cmd = [NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", SRC, "-o", out]
    subprocess.run(cmd, check=True, capture_output=True)
    lib = ctypes.CDLL(out)
    lib.emu_fake_quant.restype = ctypes.c_int
    lib.emu_fp_roundtrip.restype = ctypes.c_int
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def run_emu(lib, w, dtype, scheme, group, bits=4, E=0, M=0, sym=False, force_kind=-1, quant_dim=0, approx=0, hs=0, hf=0, tp=0):
    cfg = Cfg(dtype=dtype, scheme=scheme, bits=bits, exp_bits=E, mant_bits=M, symmetric=int(sym), group=group, variant=0, quant_dim=quant_dim, approx=approx, hi_align_start=hs, hi_align_exp_field=hf, tail_pad_bits=tp)
    w = np.ascontiguousarray(w, dtype=np.uint16)
    rows, cols = w.shape
    G = rows * cols // group if group > 0 else (1 if group == -1 else (cols if quant_dim == 1 else rows))
    out = np.empty_like(w)
    scales = np.empty(G, np.uint16)
    zeros = np.empty(G, np.uint16)
    rc = lib.emu_fake_quant(ctypes.byref(cfg), _ptr(w), _ptr(out), _ptr(scales), _ptr(zeros), ctypes.c_int64(rows), ctypes.c_int64(cols), force_kind)
    assert rc == 0, rc
    return out, scales, zeros


def _is_nan(bits, dt):
    if dt == "f16":
        return (bits & 0x7FFF) > 0x7C00
    return (bits & 0x7FFF) > 0x7F80


def assert_bits_equal(a, b, dt, msg=""):
    na, nb = _is_nan(a, dt), _is_nan(b, dt)
    assert np.array_equal(na, nb), msg
    assert np.array_equal(np.where(na, 0, a), np.where(nb, 0, b)), msg


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("bits", [2, 3, 4, 5, 6, 8])
def test_int_fast_math_matches_reference(emu, golden, dt, sym, bits):
    g = golden("golden_int")
    w = g[f"w_{dt}"]
    for grp in (32, 64, 128, 256, -1, -2):
        key = f"{dt}_b{bits}_{'sym' if sym else 'asym'}_g{grp}"
        out, scales, zeros = run_emu(emu, w, F16 if dt == "f16" else BF16, SCHEME_INT, grp, bits=bits, sym=sym)
        assert_bits_equal(out, g[f"ql_{key}_out"], dt, key)
        assert_bits_equal(scales, g[f"ql_{key}_scales"], dt, key)
        if not sym:
            assert_bits_equal(zeros, g[f"ql_{key}_zeros"], dt, key + " zeros")  # the sign of a zero zero-point included


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_int_generic_math_matches_reference(emu, golden, dt):
    g = golden("golden_int")
    w = g[f"w_{dt}"]
    for bits in (3, 4, 8):
        for sym in (False, True):
            key = f"{dt}_b{bits}_{'sym' if sym else 'asym'}_g128"
            out, _, _ = run_emu(emu, w, F16 if dt == "f16" else BF16, SCHEME_INT, 128, bits=bits, sym=sym, force_kind=KIND_INT_GENERIC)
            assert_bits_equal(out, g[f"ql_{key}_out"], dt, key)


def test_bf16_wide_path_matches_fast_and_reference(emu, golden):
    g = golden("golden_int")
    w = g["w_bf16"]
    for bits in (4, 6, 8):
        for sym in (False, True):
            key = f"bf16_b{bits}_{'sym' if sym else 'asym'}_g64"
            out, _, _ = run_emu(emu, w, BF16, SCHEME_INT, 64, bits=bits, sym=sym, force_kind=KIND_INT_WIDE)
            assert_bits_equal(out, g[f"ql_{key}_out"], "bf16", key)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_fp_codec_table_exhaustive(emu, golden, dt, fmt):
    E, M = FORMATS[fmt]
    if dt == "f16" and E >= 5:
        return
    g = golden("golden_fp_tables")
    bits = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    y = np.empty_like(bits)
    rc = emu.emu_fp_roundtrip(F16 if dt == "f16" else BF16, SCHEME_FP_CODEC, E, M, _ptr(bits), _ptr(y), ctypes.c_int64(65536))
    assert rc == 0
    from oracle import fakequant_oracle as O

    x = O.from_bits(bits, dt)
    fmax = O.fp_format(E, M)[3]
    ok = np.isfinite(x) & (np.abs(x) <= fmax)
    if dt == "bf16":
        ok &= (np.abs(x) >= 2.0 ** -126) | (x == 0)
    want = g[f"codec_{dt}_{fmt}_val"]
    bad = ok & (y != want)
    assert not bad.any(), [(hex(b), hex(a), hex(c)) for b, a, c in zip(bits[bad][:8], y[bad][:8], want[bad][:8])]


@pytest.mark.parametrize("fmt", ["e2m1", "e1m2", "e3m2", "e2m3", "e4m3"])
def test_fp_rne_table_exhaustive(emu, golden, fmt):
    E, M = FORMATS[fmt]
    g = golden("golden_fp_tables")
    bits = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    y = np.empty_like(bits)
    rc = emu.emu_fp_roundtrip(F16, SCHEME_FP_RNE, E, M, _ptr(bits), _ptr(y), ctypes.c_int64(65536))
    assert rc == 0
    ok = (bits & 0x7FFF) < 0x7C00
    want = g[f"fpscale_f16_{fmt}"]
    bad = ok & (y != want)
    assert not bad.any(), [(hex(b), hex(a), hex(c)) for b, a, c in zip(bits[bad][:8], y[bad][:8], want[bad][:8])]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_fp_chain_math_matches_reference(emu, golden, dt, fmt):
    E, M = FORMATS[fmt]
    if dt == "f16" and E >= 5:
        return
    g = golden("golden_fp_chain")
    w = g[f"w_{dt}"]
    for sym in (False, True):
        for grp in (32, 128, -1, -2):
            key = f"{dt}_{fmt}_{'sym' if sym else 'asym'}_g{grp}"
            out, scales, zeros = run_emu(emu, w, F16 if dt == "f16" else BF16, SCHEME_FP_CODEC, grp, E=E, M=M, sym=sym)
            assert_bits_equal(out, g[f"ql_{key}_out"], dt, key)


def test_fp_rne_chain_matches_reference(emu, golden):
    g = golden("golden_fp_rne")
    w = g["w_f16"]
    for grp in (32, 128, 512):
        out, _, _ = run_emu(emu, w, F16, SCHEME_FP_RNE, grp, E=2, M=1, sym=True)
        assert_bits_equal(out.reshape(-1, grp), g[f"e2m1_g{grp}"], "f16", str(grp))
    out, _, _ = run_emu(emu, w, F16, SCHEME_FP_RNE, -1, E=2, M=1, sym=True)
    assert_bits_equal(out.reshape(1, -1), g["e2m1_pt"], "f16", "pt")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_bfp_math_matches_reference(emu, golden, dt):
    g = golden("golden_bfp")
    w = g[f"w_{dt}"]
    for w_bit in (3, 4, 5, 6, 8, 12):
        for grp in (32, 64, 128, 512):
            out, _, _ = run_emu(emu, w, F16 if dt == "f16" else BF16, SCHEME_BFP, grp, bits=w_bit, sym=True)
            assert_bits_equal(out, g[f"{dt}_b{w_bit}_g{grp}"], dt, f"b{w_bit} g{grp}")


SINGLE_APPROX = [("e1m2_s", 1, 2, 1, 1, 0), ("e2m1_s", 2, 1, 1, 1, 0), ("e2m3_s", 2, 3, 4, 7, 2), ("e3m2_s", 3, 2, 4, 7, 2), ("e4m3_s", 4, 3, 12, 15, 1)]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_quant_dim1_math_matches_reference(emu, golden, dt):
    g = golden("golden_dim1_approx")
    w = g[f"w_{dt}"]
    d = F16 if dt == "f16" else BF16
    for grp in (32, 128):
        for sym in (False, True):
            tag = "sym" if sym else "asym"
            out, scales, _ = run_emu(emu, w, d, SCHEME_INT, grp, bits=4, sym=sym, quant_dim=1)
            assert_bits_equal(out, g[f"{dt}_int4_{tag}_g{grp}_d1_out"], dt, f"int4 {tag} g{grp}")
            assert_bits_equal(scales, g[f"{dt}_int4_{tag}_g{grp}_d1_scales"], dt, f"int4 {tag} g{grp} scales")
            for fmt, (E, M) in (("fp4", (2, 1)), ("fp8", (4, 3))):
                out, _, _ = run_emu(emu, w, d, SCHEME_FP_CODEC, grp, E=E, M=M, sym=sym, quant_dim=1)
                assert_bits_equal(out, g[f"{dt}_{fmt}_{tag}_g{grp}_d1_out"], dt, f"{fmt} {tag} g{grp}")
        out, _, _ = run_emu(emu, w, d, SCHEME_INT, grp, bits=8, sym=False, quant_dim=1)
        assert_bits_equal(out, g[f"{dt}_int8_asym_g{grp}_d1_out"], dt, f"int8 g{grp}")
        out, _, _ = run_emu(emu, w, d, SCHEME_BFP, grp, bits=4, sym=True, quant_dim=1)
        assert_bits_equal(out, g[f"{dt}_bfp4_g{grp}_d1_out"], dt, f"bfp4 g{grp}")
    out, _, _ = run_emu(emu, w, d, SCHEME_INT, -2, bits=4, sym=False, quant_dim=1)
    assert_bits_equal(out, g[f"{dt}_int4_asym_pc_d1_out"], dt, "per-channel d1")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("case", SINGLE_APPROX, ids=[c[0] for c in SINGLE_APPROX])
def test_single_approximate_math_matches_reference(emu, golden, dt, case):
    name, E, M, hs, hf, tp = case
    g = golden("golden_dim1_approx")
    w = g[f"w_{dt}"]
    d = F16 if dt == "f16" else BF16
    out, _, _ = run_emu(emu, w, d, SCHEME_FP_CODEC, 32, E=E, M=M, sym=True, approx=1, hs=hs, hf=hf, tp=tp)
    assert_bits_equal(out, g[f"{dt}_approx_{name}_g32_d0_out"], dt, "quant_dim 0")
    out, _, _ = run_emu(emu, w, d, SCHEME_FP_CODEC, 128, E=E, M=M, sym=True, approx=1, hs=hs, hf=hf, tp=tp, quant_dim=1)
    assert_bits_equal(out, g[f"{dt}_approx_{name}_g128_d1_out"], dt, "quant_dim 1")


DOUBLE_APPROX = [("e2m1_d", 2, 1, 0, 2, 1), ("e2m3_d", 2, 3, 0, 2, -1), ("e3m2_d", 3, 2, 4, 6, 0), ("e4m3_d", 4, 3, 13, 15, -1), ("e4m3_d0", 4, 3, 12, 15, 0)]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("case", DOUBLE_APPROX, ids=[c[0] for c in DOUBLE_APPROX])
def test_double_approximate_math_matches_reference(emu, golden, dt, case):
    """fp_decode_aligned_double_approx incl. the reference's int8 overflow, quant_dim = 1 (what its scripts run)"""
    name, E, M, hs, hf, tp = case
    g = golden("golden_dim1_approx")
    w = g[f"w_{dt}"]
    d = F16 if dt == "f16" else BF16
    out, _, _ = run_emu(emu, w, d, SCHEME_FP_CODEC, 128, E=E, M=M, sym=True, approx=2, hs=hs, hf=hf, tp=tp, quant_dim=1)
    assert_bits_equal(out, g[f"{dt}_approx_{name}_g128_d1_out"], dt, name)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("case", DOUBLE_APPROX, ids=[c[0] for c in DOUBLE_APPROX])
def test_double_approximate_table_decomposition(emu, dt, case):
    """The device decodes through tables (INFO-word step table, packed target logic, value table indexed by
    target | sign | code); on the CPU the same header functions are checked against the direct 4-code decode for
    EVERY 4-tuple of codes (FP4: 65 536 tuples) or 2 million random ones (FP6 / FP8)."""
    name, E, M, hs, hf, tp = case
    emu.emu_double_table_check.restype = ctypes.c_longlong
    emu.emu_double_table_check.argtypes = [ctypes.c_int] * 6 + [ctypes.c_longlong]
    bad = emu.emu_double_table_check(0 if dt == "f16" else 1, E, M, hs, hf, tp, 2_000_000)
    assert bad == 0, f"{name} {dt}: {bad} mismatching values"
