"""Pin the numpy oracle against outputs of the reference itself (tests/golden/*.npz).

CPU only.  The reference ships no tests or golden vectors for this path, so
these fixtures -- produced by importing /root/reference in the build container -- are the pin.
"""
import numpy as np
import pytest

from oracle import fakequant_oracle as O

INT_BITS = [2, 3, 4, 5, 6, 8]
GROUPS = [32, 64, 128, 256, -1, -2]
FORMATS = {"e2m1": (2, 1), "e1m2": (1, 2), "e3m2": (3, 2), "e2m3": (2, 3), "e4m3": (4, 3), "e5m2": (5, 2)}


def _nan_equal_bits(a, b, dtype):
    """bit equality, except that any NaN matches any NaN (payloads are not part of the contract)."""
    fa, fb = O.from_bits(a, dtype), O.from_bits(b, dtype)
    both_nan = np.isnan(fa) & np.isnan(fb)
    return np.array_equal(np.where(both_nan, 0, a), np.where(both_nan, 0, b))


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("bits", INT_BITS)
def test_int_matches_reference(golden, dtype, sym, bits):
    g = golden("golden_int")
    w = g[f"w_{dtype}"]
    for grp in GROUPS:
        key = f"{dtype}_b{bits}_{'sym' if sym else 'asym'}_g{grp}"
        r = O.int_fake_quant(w, dtype, bits, not sym, grp)
        # module form: weights, scales, zeros (quant_linear.py:885-956)
        assert _nan_equal_bits(r["out"], g[f"ql_{key}_out"], dtype), key
        assert _nan_equal_bits(r["scales"], g[f"ql_{key}_scales"], dtype), key
        if not sym:
            # bit-exact, the sign of a zero zero-point included (quant_funcs.py:22)
            assert _nan_equal_bits(r["zeros"], g[f"ql_{key}_zeros"], dtype), key
        # function form (quant_funcs.py:4-46) either asserts on NaN or returns the same bits
        if f"fn_{key}_assert" in g.files:
            assert r["has_nan"], key
        else:
            assert not r["has_nan"], key
            assert np.array_equal(r["out"], g[f"fn_{key}"]), key


def test_int_inplace_same_as_out_of_place(golden):
    g = golden("golden_int")
    assert np.array_equal(g["fn_inplace_f16_b4_asym_g128"], g["fn_f16_b4_asym_g128"])
    assert np.array_equal(g["fn_inplace_mutated_input"], g["fn_f16_b4_asym_g128"])


def test_int_output_never_negative_zero(golden):
    g = golden("golden_int")
    for k in g.files:
        if k.startswith("fn_") and not k.endswith("_assert") and "inplace" not in k:
            assert not (g[k] == 0x8000).any(), k


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_fp_codec_exhaustive(golden, dtype, fmt):
    """decode(encode(x)) for every 16-bit input, restricted to |x| <= fmax (the codec is only ever
    fed clamped values: quant_linear.py:747,755) and, for bf16, |x| >= 2^-126 (quirk Q5)."""
    E, M = FORMATS[fmt]
    g = golden("golden_fp_tables")
    bits = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    x = O.from_bits(bits, dtype)
    _, _, _, fmax = O.fp_format(E, M)
    ok = np.isfinite(x) & (np.abs(x) <= fmax)
    if dtype == "bf16":
        ok &= (np.abs(x) >= 2.0 ** -126) | (x == 0)
    code = O.fp_codec_encode(x, E, M, dtype)
    val = O.to_bits(O.fp_codec_decode(code, E, M), dtype)
    assert np.array_equal(code[ok], g[f"codec_{dtype}_{fmt}_code"][ok])
    assert np.array_equal(val[ok], g[f"codec_{dtype}_{fmt}_val"][ok])
    assert ok.sum() > 20000


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_log2_rounding_margin(dtype):
    """Quirk Q3 depends on whether log2(x), rounded to the 16-bit dtype, reaches the next integer k.
    That decision is "log2(x) >= midpoint(pred_D(k), k)".  Show that for every finite positive input
    the exact log2 is at least 64 fp32 ulps away from that midpoint, so any faithful fp32 log2 (Sleef,
    glibc, CUDA) -- and the closed-form per-binade thresholds the CUDA kernels use -- agree."""
    bits = np.arange(1, 0x7C00 if dtype == "f16" else 0x7F80, dtype=np.uint32).astype(np.uint16)
    x = O.from_bits(bits, dtype).astype(np.float64)
    x = x[x >= 2.0 ** -126]
    l = np.log2(x)
    k = np.floor(l) + 1.0  # the integer just above log2(x)
    kb = O.to_bits(k.astype(np.float32), dtype).astype(np.int32)
    # predecessor of k in D (towards -inf): k > 0 -> bits-1 ; k == 0 -> -min_subnormal ; k < 0 -> bits+1
    pred = O.from_bits(np.where(k > 0, kb - 1, np.where(k == 0, 0x8001, kb + 1)).astype(np.uint16), dtype).astype(np.float64)
    mid = (pred + k) / 2
    ulp32 = np.maximum(np.abs(l), 2.0 ** -126) * 2.0 ** -23
    notpow2 = np.floor(l) != l
    assert (np.abs(l - mid)[notpow2] > 64 * ulp32[notpow2]).all()
    # and the oracle's rounding agrees with the midpoint rule
    l16 = O.rD(l.astype(np.float32), dtype).astype(np.float64)
    assert np.array_equal((l16 == k)[notpow2], (l > mid)[notpow2])


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_fp_chain_matches_reference(golden, dtype, fmt):
    E, M = FORMATS[fmt]
    if dtype == "f16" and E >= 5:
        with pytest.raises(RuntimeError):
            O.fp_codec_fake_quant(np.zeros((2, 32), np.uint16), dtype, E, M, True, 32)
        return
    g = golden("golden_fp_chain")
    w = g[f"w_{dtype}"]
    for sym in (False, True):
        for grp in (32, 128, -1, -2):
            key = f"{dtype}_{fmt}_{'sym' if sym else 'asym'}_g{grp}"
            r = O.fp_codec_fake_quant(w, dtype, E, M, sym, grp)
            assert _nan_equal_bits(r["out"], g[f"ql_{key}_out"], dtype), key
            assert np.array_equal(r["scales_half"], g[f"ql_{key}_scales_half"]), key
            if not sym:
                assert np.array_equal(O.from_bits(r["zeros_half"], "f16"), O.from_bits(g[f"ql_{key}_zeros_half"], "f16")), key


@pytest.mark.parametrize("fmt", ["e2m1", "e1m2", "e3m2", "e2m3", "e4m3"])
def test_fp_rne_scale_exhaustive(golden, fmt):
    E, M = FORMATS[fmt]
    g = golden("golden_fp_tables")
    bits = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    x = O.from_bits(bits, "f16")
    bias = 2 ** (E - 1) - 1
    mf = (2 - 2 ** (-M)) * 2 ** (2 ** E - 1 - bias)
    q = O.fp_rne_scale(x, np.float32(1.0), M, bias, mf, -mf, "f16")
    ok = np.isfinite(x)
    assert np.array_equal(O.to_bits(q, "f16")[ok], g[f"fpscale_f16_{fmt}"][ok])


def test_fp_rne_matches_reference(golden):
    g = golden("golden_fp_rne")
    w = g["w_f16"]
    for grp in (32, 128, 512):
        r = O.fp_rne_fake_quant(w, 2, 1, grp)
        assert tuple(g[f"e2m1_g{grp}_shape"]) == r["out"].shape
        assert _nan_equal_bits(r["out"], g[f"e2m1_g{grp}"], "f16"), grp
    r = O.fp_rne_fake_quant(w, 2, 1, -1, per_tensor=True)
    assert tuple(g["e2m1_pt_shape"]) == r["out"].shape
    assert _nan_equal_bits(r["out"], g["e2m1_pt"], "f16")
    # the reference's own __main__ smoke input (fp4_quantize_cpu.py:74-80)
    r = O.fp_rne_fake_quant(g["main_w"], 2, 1, 128)
    assert np.array_equal(r["out"], g["main_out"])


def test_error_contract():
    w = np.zeros((4, 100), np.uint16)
    with pytest.raises(AssertionError):
        O.int_fake_quant(w, "f16", 4, True, 32)  # quant_funcs.py:11
    with pytest.raises(AssertionError):
        O.int_fake_quant(np.zeros((4,), np.uint16), "f16", 4, True, -2)  # quant_funcs.py:15
    with pytest.raises(ValueError):
        O.int_fake_quant(w, "f16", 4, True, -3)  # quant_linear.py:906
    with pytest.raises(ValueError):
        O.fp_rne_fake_quant(w, 2, 1, 32)  # fp4_quantize_cpu.py:55-56


def test_torch_cpu_port_matches_reference(golden):
    """The eager-torch port that bench.py times as the CPU baseline gives the reference's bits."""
    import torch

    from oracle import torch_cpu_port as P

    g = golden("golden_int")
    for dtype, tdt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
        w = torch.from_numpy(g[f"w_{dtype}"].view(np.int16).copy()).view(tdt)
        for bits, sym, grp in [(4, False, 128), (4, True, 64), (8, False, -2), (3, True, 32)]:
            key = f"{dtype}_b{bits}_{'sym' if sym else 'asym'}_g{grp}"
            o = P.int_fake_quant(w, bits, not sym, grp)
            assert np.array_equal(o.view(torch.int16).numpy().view(np.uint16), g[f"fn_{key}"]), key
    g = golden("golden_fp_chain")
    w = torch.from_numpy(g["w_f16"].view(np.int16).copy()).view(torch.float16)
    for fmt, (E, M) in (("e2m1", (2, 1)), ("e4m3", (4, 3))):
        for sym in (False, True):
            key = f"f16_{fmt}_{'sym' if sym else 'asym'}_g128"
            o = P.fp_codec_fake_quant(w, E, M, sym, 128)
            assert _nan_equal_bits(o.view(torch.int16).numpy().view(np.uint16), g[f"ql_{key}_out"], "f16"), key


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_bfp_matches_reference(golden, dtype):
    g = golden("golden_bfp")
    w = g[f"w_{dtype}"]
    for w_bit in (3, 4, 5, 6, 8, 12):
        for grp in (32, 64, 128, 512):
            r = O.bfp_fake_quant(w, dtype, w_bit, grp)
            assert _nan_equal_bits(r["out"], g[f"{dtype}_b{w_bit}_g{grp}"], dtype), (w_bit, grp)
    with pytest.raises(ValueError):
        O.bfp_fake_quant(w, dtype, 4, -1)


APPROX_CASES = [  # must mirror tests/golden/make_golden.py
    ("e1m2_s", 1, 2, 1, 1, 0, False),
    ("e2m1_s", 2, 1, 1, 1, 0, False),
    ("e2m1_d", 2, 1, 0, 2, 1, True),
    ("e2m3_s", 2, 3, 4, 7, 2, False),
    ("e2m3_d", 2, 3, 0, 2, -1, True),
    ("e3m2_s", 3, 2, 4, 7, 2, False),
    ("e3m2_d", 3, 2, 4, 6, 0, True),
    ("e4m3_s", 4, 3, 12, 15, 1, False),
    ("e4m3_d", 4, 3, 13, 15, -1, True),
    ("e4m3_d0", 4, 3, 12, 15, 0, True),
]


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_quant_dim1_matches_reference(golden, dtype):
    g = golden("golden_dim1_approx")
    w = g[f"w_{dtype}"]
    for grp in (32, 128):
        for sym in (False, True):
            tag = "sym" if sym else "asym"
            r = O.along_output_dim(O.int_fake_quant, w, dtype, 4, not sym, grp)
            assert _nan_equal_bits(r["out"], g[f"{dtype}_int4_{tag}_g{grp}_d1_out"], dtype), (grp, sym)
            assert _nan_equal_bits(r["scales"], g[f"{dtype}_int4_{tag}_g{grp}_d1_scales"], dtype), (grp, sym)
            for fmt, (E, M) in (("fp4", (2, 1)), ("fp8", (4, 3))):
                r = O.along_output_dim(O.fp_codec_fake_quant, w, dtype, E, M, sym, grp)
                assert _nan_equal_bits(r["out"], g[f"{dtype}_{fmt}_{tag}_g{grp}_d1_out"], dtype), (fmt, grp, sym)
                assert np.array_equal(r["scales_half"], g[f"{dtype}_{fmt}_{tag}_g{grp}_d1_scales"]), (fmt, grp, sym)
        r = O.along_output_dim(O.int_fake_quant, w, dtype, 8, True, grp)
        assert _nan_equal_bits(r["out"], g[f"{dtype}_int8_asym_g{grp}_d1_out"], dtype)
        r = O.along_output_dim(O.bfp_fake_quant, w, dtype, 4, grp)
        assert _nan_equal_bits(r["out"], g[f"{dtype}_bfp4_g{grp}_d1_out"], dtype)
    r = O.along_output_dim(O.int_fake_quant, w, dtype, 4, True, -2)
    assert _nan_equal_bits(r["out"], g[f"{dtype}_int4_asym_pc_d1_out"], dtype)


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
@pytest.mark.parametrize("case", APPROX_CASES, ids=[c[0] for c in APPROX_CASES])
def test_approximate_decode_matches_reference(golden, dtype, case):
    name, E, M, hs, hf, tp, dbl = case
    g = golden("golden_dim1_approx")
    w = g[f"w_{dtype}"]
    r = O.fp_approx_fake_quant(w, dtype, E, M, 32, hs, hf, tp, dbl)
    assert _nan_equal_bits(r["out"], g[f"{dtype}_approx_{name}_g32_d0_out"], dtype), "quant_dim 0"
    assert np.array_equal(r["scales_half"], g[f"{dtype}_approx_{name}_g32_d0_scales"])
    r = O.along_output_dim(O.fp_approx_fake_quant, w, dtype, E, M, 128, hs, hf, tp, dbl)
    assert _nan_equal_bits(r["out"], g[f"{dtype}_approx_{name}_g128_d1_out"], dtype), "quant_dim 1"
    assert np.array_equal(r["scales_half"], g[f"{dtype}_approx_{name}_g128_d1_scales"])


ODD_CASES = [((3, 5), dict(per_tensor=True), -1), ((7, 129), dict(per_tensor=True), -1), ((7, 129), dict(), -2), ((5, 36), dict(q_group_size=12), 12),
             ((1, 8), dict(per_tensor=True), -1), ((2, 3, 40), dict(q_group_size=20), 20), ((6, 1000), dict(), -2), ((1, 3), dict(), -2), ((3, 51216), dict(), -2)]


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_int_odd_shapes_match_reference(golden, dtype):
    """The oracle on shapes whose segments are not whole 16-byte vectors, against what the reference itself returns
    (quant_funcs.py:10-15 takes any shape; tests/golden/golden_odd.npz)."""
    g = golden("golden_odd")
    for ci, (shape, _kw, grp) in enumerate(ODD_CASES):
        wb = g[f"w_{dtype}_{ci}"]
        assert wb.shape == shape
        for bits, zp in ((4, True), (4, False), (8, True), (3, False)):
            ref = g[f"{dtype}_{ci}_b{bits}_{'asym' if zp else 'sym'}"]
            got = O.int_fake_quant(wb.reshape(-1, shape[-1]), dtype, bits, zp, grp)["out"].reshape(shape)
            assert _nan_equal_bits(got, ref, dtype), (shape, bits, zp)


WIDE_APPROX = [c for c in APPROX_CASES if not c[6] and c[0] not in ("e1m2_s", "e2m3_s")]


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_quant_dim1_wide_matches_reference(golden, dtype):
    """quant_dim = 1 on the [256, 264] fixture (the shape class col_wide_kernel serves): oracle == reference."""
    g = golden("golden_dim1_wide")
    w = g[f"w_{dtype}"]
    r = O.along_output_dim(O.int_fake_quant, w, dtype, 4, True, 128)
    assert _nan_equal_bits(r["out"], g[f"{dtype}_int4_asym_g128_out"], dtype)
    assert _nan_equal_bits(r["scales"], g[f"{dtype}_int4_asym_g128_scales"], dtype)
    assert np.array_equal(r["zeros"], g[f"{dtype}_int4_asym_g128_zeros"])
    r = O.along_output_dim(O.int_fake_quant, w, dtype, 3, False, 64)
    assert _nan_equal_bits(r["out"], g[f"{dtype}_int3_sym_g64_out"], dtype)
    r = O.along_output_dim(O.bfp_fake_quant, w, dtype, 4, 128)
    assert _nan_equal_bits(r["out"], g[f"{dtype}_bfp4_g128_out"], dtype)
    r = O.along_output_dim(O.fp_codec_fake_quant, w, dtype, 2, 1, True, 128)
    assert _nan_equal_bits(r["out"], g[f"{dtype}_fp4_sym_g128_out"], dtype)
    r = O.along_output_dim(O.fp_codec_fake_quant, w, dtype, 4, 3, False, 64)
    assert _nan_equal_bits(r["out"], g[f"{dtype}_fp8_asym_g64_out"], dtype)
    for name, E, M, hs, hf, tp, _ in WIDE_APPROX:
        for grp in (64, 128):
            r = O.along_output_dim(O.fp_approx_fake_quant, w, dtype, E, M, grp, hs, hf, tp, False)
            assert _nan_equal_bits(r["out"], g[f"{dtype}_approx_{name}_g{grp}_out"], dtype), (name, grp)
            assert np.array_equal(r["scales_half"], g[f"{dtype}_approx_{name}_g{grp}_scales"]), (name, grp)
