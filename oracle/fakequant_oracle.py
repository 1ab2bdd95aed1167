"""CPU oracle for the weight fake-quant hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's arithmetic, written as an explicit
"fp32 compute, round to the storage dtype after every op" model.  It is the CHECKER:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  Nothing under `iron_weight_only_quant_b200/` imports it, and the product
path raises when the CUDA library is missing -- there is no CPU fallback.

Parity pin: the reference has no tests or golden vectors of its own, so the
oracle is pinned against outputs of the reference itself, produced in the build container by
`tests/golden/make_golden.py` (imports /root/reference, CPU, torch 2.11) and committed under
`tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks every function below against them.

Conventions
-----------
* A "D array" is a numpy float32 array whose values are all exactly representable in the storage
  dtype D (fp16 or bf16).  `to_bits` / `from_bits` convert to / from the uint16 bit patterns.
* `rD(x)` rounds a float32 array to D (round-to-nearest-even) and returns a D array.
* All reference file:line citations are relative to /root/reference.
"""
from __future__ import annotations

import numpy as np

F16 = "f16"
BF16 = "bf16"

# --------------------------------------------------------------------------------------------
# storage-dtype model
# --------------------------------------------------------------------------------------------


def from_bits(bits: np.ndarray, dtype: str) -> np.ndarray:
    """uint16 bit patterns -> D array (float32 values)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint16)
    if dtype == F16:
        return bits.view(np.float16).astype(np.float32)
    if dtype == BF16:
        return (bits.astype(np.uint32) << np.uint32(16)).view(np.float32)
    raise ValueError(dtype)


def to_bits(x: np.ndarray, dtype: str) -> np.ndarray:
    """D array (or any float32 array; it is rounded) -> uint16 bit patterns."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if dtype == F16:
        with np.errstate(over="ignore"):
            return x.astype(np.float16).view(np.uint16)
    if dtype == BF16:
        u = x.view(np.uint32)
        lsb = (u >> np.uint32(16)) & np.uint32(1)
        r = (u + np.uint32(0x7FFF) + lsb) >> np.uint32(16)
        nan = np.isnan(x)
        if nan.any():
            r = np.where(nan, (u >> np.uint32(16)) | np.uint32(0x40), r)
        return r.astype(np.uint16)
    raise ValueError(dtype)


def rD(x, dtype: str) -> np.ndarray:
    """Round float32 values to the storage dtype (RNE), keep them in float32."""
    return from_bits(to_bits(np.asarray(x, dtype=np.float32), dtype), dtype)


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _div(a, b):
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return _f32(a) / _f32(b)


def _clamp(x, lo=None, hi=None):
    """torch.clamp semantics: NaN propagates."""
    x = _f32(x)
    nan = np.isnan(x)
    if lo is not None:
        x = np.maximum(x, np.float32(lo))
    if hi is not None:
        x = np.minimum(x, np.float32(hi))
    return np.where(nan, np.float32(np.nan), x)


def _group_view(w: np.ndarray, group: int):
    """Grouping of quant_funcs.py:10-15 / quant_linear.py:896-906.

    group > 0 : rows of `group` contiguous elements; -1: one row holding the whole tensor;
    -2 : one row per matrix row (also the `q_group_size=-1, per_tensor=False` default of the
    function form).
    """
    if w.ndim != 2:
        raise AssertionError("expected a 2-D tensor")
    rows, cols = w.shape
    if group > 0:
        if cols % group != 0:
            raise AssertionError("last dim not divisible by group size")
        return w.reshape(-1, group)
    if group == -1:
        return w.reshape(1, -1)
    if group == -2:
        return w.reshape(rows, -1)
    raise ValueError("Invalid w_group_size")


# --------------------------------------------------------------------------------------------
# INT min-max fake quant   (quant_funcs.py:3-46, quant_linear.py:885-956)
# --------------------------------------------------------------------------------------------


def int_params(n_bits: int, zero_point: bool):
    """max_int/min_int of quant_funcs.py:19-20 (asym) and :26-27 (sym)."""
    if zero_point:
        return 0, 2 ** n_bits - 1
    return -(2 ** (n_bits - 1)), 2 ** (n_bits - 1) - 1


def int_ranges_to_qparams(mx, mn, n_bits: int, zero_point: bool, dtype: str):
    """(max, min) or (absmax, -) per segment -> (scales, zeros) as D arrays.

    asym: quant_funcs.py:17-22 -- scales = (max-min).clamp(min=1e-5) / max_int ;
          zeros = (-round(min/scales)).clamp(min_int, max_int)
    sym : quant_funcs.py:24-29 -- scales = absmax.clamp(min=1e-5) / max_int ; zeros = 0
    Every op result is rounded to D (the tensors are D tensors in the reference).
    """
    min_int, max_int = int_params(n_bits, zero_point)
    if zero_point:
        d = rD(_f32(mx) - _f32(mn), dtype)
        d = rD(_clamp(d, lo=1e-5), dtype)
        s = rD(_div(d, np.float32(max_int)), dtype)
        z = rD(_div(mn, s), dtype)
        z = -np.rint(z)
        # torch.clamp keeps the sign of a zero (ATen CPU: min(max(x, lo), hi) returns x when x == lo): min/s in [0, 0.5]
        # -> round = +0 -> negated -0.0 -> stays -0.0; a negative value comes back as the bound +0.0
        z = np.where(z == 0, z, _clamp(z, min_int, max_int))
    else:
        d = rD(_clamp(mx, lo=1e-5), dtype)
        s = rD(_div(d, np.float32(max_int)), dtype)
        z = np.zeros_like(s)
    return s.astype(np.float32), z.astype(np.float32)


def int_apply(g: np.ndarray, s: np.ndarray, z: np.ndarray, n_bits: int, zero_point: bool, dtype: str):
    """Element stage, quant_funcs.py:36-38:  (clamp(round(w/s)+z, min_int, max_int) - z) * s."""
    min_int, max_int = int_params(n_bits, zero_point)
    q = rD(_div(g, s), dtype)
    q = np.rint(q)
    with np.errstate(invalid="ignore"):
        q = rD(q + z, dtype)
        q = _clamp(q, min_int, max_int)
        code = q
        out = rD(q - z, dtype)
        out = rD(out * s, dtype)
    return out, code


def int_fake_quant(w_bits: np.ndarray, dtype: str, n_bits: int, zero_point: bool, group: int):
    """Full INT path.  `w_bits`: 2-D uint16 array of D bit patterns.

    Returns dict(out=uint16 [rows,cols], scales=uint16 [G], zeros=uint16 [G], codes=int32 [rows,cols],
    has_nan=bool).  Mirrors quant_funcs.pseudo_quantize_tensor (quant_funcs.py:4-46) and the inline
    duplicate in QuantLinear.quantize_weight (quant_linear.py:885-956); `has_nan` is the condition
    of the assert at quant_funcs.py:40.
    """
    w = from_bits(w_bits, dtype)
    shape = w.shape
    g = _group_view(w, group)
    with np.errstate(invalid="ignore"):
        if zero_point:
            mx = _nanmax(g)
            mn = _nanmin(g)
        else:
            mx = _nanmax(np.abs(g))
            mn = None
    s, z = int_ranges_to_qparams(mx, mn, n_bits, zero_point, dtype)
    out, code = int_apply(g, s, z, n_bits, zero_point, dtype)
    return dict(
        out=to_bits(out, dtype).reshape(shape),
        scales=to_bits(s, dtype).reshape(-1),
        zeros=to_bits(z, dtype).reshape(-1),
        codes=np.where(np.isnan(code), -1, code).astype(np.int32).reshape(shape),
        has_nan=bool(np.isnan(out).any()),
    )


def _nanmax(g):
    """torch.amax(dim=1, keepdim=True): NaN propagates (np.max does too)."""
    return np.max(g, axis=1, keepdims=True).astype(np.float32)


def _nanmin(g):
    return np.min(g, axis=1, keepdims=True).astype(np.float32)


# --------------------------------------------------------------------------------------------
# minifloat codec of quant_linear.py  (quant_linear.py:126-163 encode, :213-235 decode)
# --------------------------------------------------------------------------------------------


def fp_format(exp_bits: int, mant_bits: int):
    """bias, emin, emax, fmax exactly as configure_fp_formats / the *_max_value expressions
    (quant_linear.py:84-110, :741, :803, :846): the all-ones exponent is a normal binade."""
    bias = 2 ** (exp_bits - 1) - 1
    emin = 1 - bias
    emax = (1 << exp_bits) - 1 - bias
    fmax = (1.0 + (2 ** mant_bits - 1) / (2 ** mant_bits)) * (2.0 ** emax)
    return bias, emin, emax, fmax


def _log2_D(a: np.ndarray, dtype: str) -> np.ndarray:
    """torch.log2 on a D tensor: fp32 log2, result rounded to D (quirk Q3).  The fp32 log2 is taken
    as the correctly rounded one (float64 log2 rounded to fp32); tests/test_oracle_golden.py checks
    that this reproduces the reference on every 16-bit input and that no input sits within 4 fp32
    ulps of a rounding boundary, so the choice of libm does not matter."""
    with np.errstate(divide="ignore", invalid="ignore"):
        l = np.log2(a.astype(np.float64)).astype(np.float32)
    return rD(l, dtype)


def fp_codec_encode(n: np.ndarray, exp_bits: int, mant_bits: int, dtype: str) -> np.ndarray:
    """_float_to_fp (quant_linear.py:126-163) on a D array `n`; returns uint8 codes.

    e_raw = floor(log2_D(|n|)) cast to int8 (:139), subnormal test (:142), exponent clamp (:145),
    fp32 mantissa rounding WITHOUT carry (:148-149), subnormal mantissa (:152-154), packing (:159),
    zero -> code 0 (:161).
    """
    bias, emin, emax, _ = fp_format(exp_bits, mant_bits)
    n = _f32(n)
    sign = (n < 0).astype(np.uint8)
    a = np.abs(n)
    zero = a == 0
    # :133 -- zeros are replaced by 1e-8 *in D*; in fp16 that is 0 again and log2 gives -inf; the
    # result is masked by `zero` at :161, so any finite stand-in gives the same code.
    a_safe = np.where(zero, np.float32(1.0), a)
    l = _log2_D(a_safe, dtype)
    with np.errstate(invalid="ignore"):
        e_raw = np.floor(l)
    # .to(torch.int8) at :139 -- wraps modulo 256 for |e_raw| > 127 (quirk Q5, bf16 only)
    e_raw = np.where(np.isfinite(e_raw), e_raw, 0).astype(np.int64)
    e_raw = ((e_raw + 128) % 256) - 128
    sub = e_raw < emin
    e = np.clip(e_raw, emin, emax)
    mscale = np.float32(1 << mant_bits)
    p2 = np.exp2(e.astype(np.float32)).astype(np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        m_n = np.rint((_div(a_safe, p2) - np.float32(1.0)) * mscale)
        m_n = np.clip(m_n, 0, (1 << mant_bits) - 1)
        m_s = np.rint(_div(a_safe, np.float32(2.0 ** emin)) * mscale)
        m_s = np.clip(m_s, 0, (1 << mant_bits) - 1)
    m_n = np.where(np.isnan(m_n), 0, m_n).astype(np.uint8)
    m_s = np.where(np.isnan(m_s), 0, m_s).astype(np.uint8)
    ef = np.where(sub, 0, e + bias).astype(np.uint8)
    mf = np.where(sub, m_s, m_n).astype(np.uint8)
    code = (sign << np.uint8(exp_bits + mant_bits)) | (ef << np.uint8(mant_bits)) | mf
    return np.where(zero, np.uint8(0), code).astype(np.uint8)


def fp_codec_decode(code: np.ndarray, exp_bits: int, mant_bits: int) -> np.ndarray:
    """_fp_to_float (quant_linear.py:213-235): uint8 codes -> float32 values."""
    bias, emin, _, _ = fp_format(exp_bits, mant_bits)
    code = code.astype(np.uint8)
    sign = ((code >> np.uint8(exp_bits + mant_bits)) & np.uint8(1)).astype(np.float32)
    raw_e = ((code >> np.uint8(mant_bits)) & np.uint8((1 << exp_bits) - 1)).astype(np.float32)
    mf = (code & np.uint8((1 << mant_bits) - 1)).astype(np.float32)
    v_n = (np.float32(1.0) + mf / np.float32(1 << mant_bits)) * np.exp2(raw_e - np.float32(bias))
    v_s = (mf / np.float32(1 << mant_bits)) * np.float32(2.0 ** emin)
    v = np.where(raw_e == 0, v_s, v_n).astype(np.float32)
    v = np.where(sign == 1, -v, v)  # (-1)**sign * v keeps the sign of a zero magnitude (quirk Q4)
    return np.where(code == 0, np.float32(0.0), v).astype(np.float32)


def fp_codec_roundtrip(n: np.ndarray, exp_bits: int, mant_bits: int, dtype: str) -> np.ndarray:
    """decode(encode(n)) as a D array (the `.to(weight.dtype)` at quant_linear.py:774)."""
    return rD(fp_codec_decode(fp_codec_encode(n, exp_bits, mant_bits, dtype), exp_bits, mant_bits), dtype)


def fp_codec_fake_quant(w_bits: np.ndarray, dtype: str, exp_bits: int, mant_bits: int, symmetric: bool, group: int):
    """QuantLinear.quantize_weight FP4/FP6/FP8 branch (quant_linear.py:724-883; the three branches are
    the same code with different globals).

    sym  (:744-747): S = (absmax.clamp(1e-5) / fmax).clamp(1e-5); n = clamp(w / S, +-fmax)
    asym (:749-755): mid = (max+min)*0.5; span = ((max-min)*0.5).clamp(1e-5); S = (span/fmax).clamp(1e-5)
                     n = clamp((w - mid)/S, +-fmax)
    out  (:774-779): D(decode(code)) * S  [+ mid.half().to(D)]
    Stored buffers (:761-768) are `.half()` copies (quirk Q1): `scales_half`/`zeros_half` below.
    """
    _, _, _, fmax = fp_format(exp_bits, mant_bits)
    if dtype == F16 and fmax > 65504.0:
        # torch.clamp(min=-fmax, max=fmax) on a half tensor raises for an unrepresentable bound
        raise RuntimeError("value cannot be converted to type at::Half without overflow")
    w = from_bits(w_bits, dtype)
    shape = w.shape
    g = _group_view(w, group)
    with np.errstate(invalid="ignore", over="ignore"):
        if symmetric:
            mx = rD(_clamp(_nanmax(np.abs(g)), lo=1e-5), dtype)
            S = rD(_clamp(rD(_div(mx, np.float32(fmax)), dtype), lo=1e-5), dtype)
            mid = None
            n = _clamp(rD(_div(g, S), dtype), -fmax, fmax)
        else:
            mx = _nanmax(g)
            mn = _nanmin(g)
            mid = rD(rD(mx + mn, dtype) * np.float32(0.5), dtype)
            span = rD(_clamp(rD(rD(mx - mn, dtype) * np.float32(0.5), dtype), lo=1e-5), dtype)
            S = rD(_clamp(rD(_div(span, np.float32(fmax)), dtype), lo=1e-5), dtype)
            n = _clamp(rD(_div(rD(g - mid, dtype), S), dtype), -fmax, fmax)
        code = fp_codec_encode(n, exp_bits, mant_bits, dtype)
        v = rD(fp_codec_decode(code, exp_bits, mant_bits), dtype)
        out = rD(v * S, dtype)
        if mid is not None:
            mid_half = rD(from_bits(to_bits(mid, F16), F16), dtype)  # zeros.half().to(D)
            out = rD(out + mid_half, dtype)
    res = dict(
        out=to_bits(out, dtype).reshape(shape),
        scales=to_bits(S, dtype).reshape(-1),
        scales_half=to_bits(S, F16).reshape(-1),
        codes=code.reshape(shape),
        has_nan=bool(np.isnan(out).any()),
    )
    if mid is not None:
        res["zeros"] = to_bits(mid, dtype).reshape(-1)
        res["zeros_half"] = to_bits(mid, F16).reshape(-1)
    return res


# --------------------------------------------------------------------------------------------
# ideal-RNE minifloat emulation of fp4_quantize_cpu.py  (fp4_quantize_cpu.py:37-72)
# --------------------------------------------------------------------------------------------


def fp_rne_scale(u_in: np.ndarray, S: np.ndarray, M: int, bias: int, max_float: float, min_float: float, dtype: str):
    """_fp_scale (fp4_quantize_cpu.py:37-44) on D arrays; returns the D array `tensor_q`
    (NOT multiplied back by S -- the caller does that at :72)."""
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        u = rD(_div(u_in, S), dtype)
        u = _clamp(u, min_float, max_float)
        l = _log2_D(np.abs(u), dtype)
        L = np.floor(rD(l + np.float32(bias), dtype))
        L = _clamp(L, lo=1.0)
        sc = rD(np.exp2(rD(rD(L - np.float32(M), dtype) - np.float32(bias), dtype)), dtype)
        q = np.rint(rD(_div(u, sc), dtype))
        q = rD(q * sc, dtype)
    return q


def fp_rne_fake_quant(w_bits: np.ndarray, exp_bits: int, mant_bits: int, group: int, per_tensor: bool = False, dtype: str = F16):
    """quantize_fp16_to_fp4_e1m2 (fp4_quantize_cpu.py:47-72) generalised from its hard-coded E=2,M=1
    (:61-62).  max_float (:64) = (2 - 2^-M) * 2^(2^E - 1 - bias);  S = absmax.clamp(1e-8) / max_float
    (:67-69) -- 1e-8 rounds to 0 in fp16, so an all-zero group gives S = 0 and NaNs, as in the
    reference.  Output keeps the reference's grouped shape `(-1, group)` / `(1, N)` (:57-59, :72)."""
    if w_bits.ndim != 2:
        raise ValueError("Expected a 2D tensor of shape [out_features, in_features].")
    w = from_bits(w_bits, dtype)
    if group > 0:
        if w.shape[1] % group != 0:
            raise ValueError("in_features must be divisible by group_size.")
        w = w.reshape(-1, group)
    if per_tensor:
        w = w.reshape(1, -1)
    bias = 2 ** (exp_bits - 1) - 1
    max_float = (2 - 2 ** (-mant_bits)) * 2 ** (2 ** exp_bits - 1 - bias)
    with np.errstate(invalid="ignore", over="ignore"):
        mx = rD(_clamp(_nanmax(np.abs(w)), lo=1e-8), dtype)
        S = rD(_div(mx, np.float32(max_float)), dtype)
        q = fp_rne_scale(w, S, mant_bits, bias, max_float, -max_float, dtype)
        out = rD(q * S, dtype)
    return dict(out=to_bits(out, dtype), scales=to_bits(S, dtype).reshape(-1), has_nan=bool(np.isnan(out).any()))


# --------------------------------------------------------------------------------------------
# sharded per-tensor helper (no reference counterpart: SURVEY.md section 8e)
# --------------------------------------------------------------------------------------------


def minmax_of(w_bits: np.ndarray, dtype: str):
    """(max, min, absmax) of a shard as float32 scalars -- what each rank contributes to the
    min/max all-reduce; exact and order independent, so sharded == unsharded."""
    w = from_bits(w_bits, dtype)
    return np.float32(w.max()), np.float32(w.min()), np.float32(np.abs(w).max())


# --------------------------------------------------------------------------------------------
# block floating point (shared exponent per group)   quant_linear.py:648-723, _rounding_rshift :112-123
# --------------------------------------------------------------------------------------------


def bfp_fake_quant(w_bits: np.ndarray, dtype: str, w_bit: int, group: int):
    """QuantLinear.quantize_weight, weight_format='bfp' (quant_linear.py:648-723), followed literally:
    fp16 field split (:658-661), implicit one except for exponent field 0 (:664-665), truncating right shift to
    the group's max exponent FIELD (:668-670), round-half-up right shift to `min(w_bit-1, 11)` mantissa bits
    (:673-698, _rounding_rshift :112-123), clamp (:700-701), dequantise with 2^(exp_block-15-frac) (:704-708).
    Note the reference's treatment of fp16 subnormals (exponent field 0 is used as if it were exponent -15,
    i.e. subnormal inputs count half) is reproduced, not fixed."""
    if group <= 0:
        raise ValueError("BFP 仅支持分组量化，请将 w_group_size 设为正数")
    w = from_bits(w_bits, dtype)
    shape = w.shape
    g = _group_view(w, group)
    with np.errstate(over="ignore"):
        b = g.astype(np.float16).view(np.uint16).astype(np.int64)  # .to(torch.float16).view(int16).to(int32)
    sign = (b >> 15) & 1
    exp = (b >> 10) & 0x1F
    mant = b & 0x3FF
    lead = np.where(exp == 0, 0, 1)
    m = (lead << 10) | mant
    exp_block = exp.max(axis=1, keepdims=True)
    shift = np.clip(exp_block - exp, 0, None)
    m = m >> shift
    target = min(w_bit - 1, 11)
    sd = max(0, 11 - target)
    if sd > 0:
        m = (m + (1 << (sd - 1))) >> sd
    m = np.minimum(m, (1 << target) - 1)
    frac = target - 1
    scale = np.exp2((exp_block - 15).astype(np.float32) - np.float32(frac)).astype(np.float32)
    sf = np.where(sign == 1, np.float32(-1.0), np.float32(1.0))
    with np.errstate(over="ignore"):
        out = rD(rD(m.astype(np.float32), dtype) * scale * rD(sf, dtype), dtype)
    return dict(out=to_bits(out, dtype).reshape(shape))


# --------------------------------------------------------------------------------------------
# quant_dim = 1: groups run along the OUTPUT dimension      quant_linear.py:478-483, :640-647
# --------------------------------------------------------------------------------------------


def along_output_dim(fn, w_bits: np.ndarray, *args, **kw):
    """The reference transposes the weight, quantises with groups along the (new) last dim and transposes
    back (`weight.t()` ... `_reshape_back`, quant_linear.py:640-647).  `fn` is any oracle above that takes the
    2-D bit array first; per-segment outputs (scales, ...) keep the transposed group order, like the reference."""
    r = fn(np.ascontiguousarray(w_bits.T), *args, **kw)
    r = dict(r)
    r["out"] = np.ascontiguousarray(r["out"].T)
    if "codes" in r and getattr(r["codes"], "ndim", 0) == 2:
        r["codes"] = np.ascontiguousarray(r["codes"].T)
    return r


# --------------------------------------------------------------------------------------------
# "approximate" aligned decode      quant_linear.py:237-285 (single), :288-363 (double), driver :470-632
# --------------------------------------------------------------------------------------------


def _wrap8(x):
    """two's-complement wrap to int8: the double-approximate decoder keeps its fields in torch.int8 tensors
    (quant_linear.py:311-313), so `1 << 7`, `val + offset` and left shifts overflow exactly like this."""
    return ((np.asarray(x, dtype=np.int64) + 128) % 256) - 128


def _rounding_rshift(val, shift, int8: bool = False):
    """quant_linear.py:112-123: (val + (1 << (shift-1) if shift > 0 else 0)) >> shift, element-wise.
    int8=True reproduces the overflow of the int8 tensors the double-approximate decoder passes in."""
    val = np.asarray(val, dtype=np.int64)
    shift = np.broadcast_to(np.asarray(shift, dtype=np.int64), val.shape)
    off = np.where(shift > 0, np.int64(1) << np.minimum(np.maximum(shift - 1, 0), 40), 0)
    if int8:
        return _wrap8(val + _wrap8(off)) >> np.minimum(shift, 62)
    return (val + off) >> shift


def fp_decode_aligned(code, hi_align_start, hi_align_exp_field, tail_pad_bits, exp_bits, mant_bits, dtype):
    """_fp_decode_aligned (quant_linear.py:237-285) with align_subnorm_exp_as_one=True,
    limit_align_exp_to_field=True, decode_dtype = weight dtype -- the only way the driver calls it
    (:499-509 etc.).  Returns a D array."""
    bias = 2 ** (exp_bits - 1) - 1
    code = np.asarray(code, dtype=np.uint8).astype(np.int64)
    zero = code == 0
    sign = (code >> (exp_bits + mant_bits)) & 1
    ef = (code >> mant_bits) & ((1 << exp_bits) - 1)
    mf = code & ((1 << mant_bits) - 1)
    align_exp = np.where(ef == 0, 1, ef)
    lead = np.where(ef == 0, 0, 1)
    mant_full = (lead << mant_bits) | mf
    if tail_pad_bits >= 0:
        mant_padded = mant_full << tail_pad_bits
    else:
        mant_padded = _rounding_rshift(mant_full, abs(tail_pad_bits))
    # value_normal is computed in the weight dtype (every op rounds to D)                        :268-269
    exp_unb = rD(np.where(ef == 0, 1 - bias, ef - bias).astype(np.float32), dtype)
    vn = rD(rD(mant_full.astype(np.float32), dtype) / np.float32(2.0 ** mant_bits), dtype)
    vn = rD(vn * rD(np.exp2(exp_unb), dtype), dtype)
    hi = (align_exp >= hi_align_start) & (align_exp <= hi_align_exp_field)
    shift = np.clip(hi_align_exp_field - align_exp, 0, None)
    mant_aligned = _rounding_rshift(mant_padded, shift)
    # value_hi is fp32 (.float(), python scalars)                                                :281
    vh = mant_aligned.astype(np.float32) / np.float32(2.0 ** (mant_bits + tail_pad_bits)) * np.float32(2.0 ** float(hi_align_exp_field - bias))
    v = np.where(hi, vh, vn).astype(np.float32)
    v = np.where(sign == 1, -v, v)
    v = np.where(zero, np.float32(0.0), v)
    return rD(v, dtype)  # `.to(self.weight.data.dtype)` at the call site


def fp_decode_aligned_double(code_2d, hi_align_start, hi_align_exp_field, tail_pad_bits, exp_bits, mant_bits, dtype):
    """fp_decode_aligned_double_approx (quant_linear.py:288-363) with align_subnorm_exp_as_one=True and
    handle_max_outlier=True (the driver's arguments).  `code_2d` is the [G, group] code matrix; groups of 4
    are taken along the flattened TRANSPOSE (:309, :329), the result has the shape of `code_2d`."""
    bias = 2 ** (exp_bits - 1) - 1
    code = np.asarray(code_2d, dtype=np.uint8).T.astype(np.int64)
    shape_t = code.shape
    zero = code == 0
    sign = (code >> (exp_bits + mant_bits)) & 1
    ef = (code >> mant_bits) & ((1 << exp_bits) - 1)
    mf = code & ((1 << mant_bits) - 1)
    align_exp = np.where(ef == 0, 1, ef)
    lead = np.where(ef == 0, 0, 1)
    mant_full = (lead << mant_bits) | mf
    mant_padded = _wrap8(mant_full << tail_pad_bits) if tail_pad_bits >= 0 else _rounding_rshift(mant_full, abs(tail_pad_bits), int8=True)
    if code.size % 4 != 0:
        raise ValueError("double approx requires total elements divisible by 4")
    e4 = align_exp.reshape(-1, 4)
    m4 = mant_padded.reshape(-1, 4)
    s4 = sign.reshape(-1, 4)
    z4 = zero.reshape(-1, 4)
    outl = (e4 < hi_align_start) | (e4 > hi_align_exp_field)
    cnt = outl.sum(axis=1, keepdims=True)
    gmax = e4.max(axis=1, keepdims=True)
    target = np.where(cnt <= 1, hi_align_exp_field, gmax)
    max_exp = (1 << exp_bits) - 1
    has_max = ((e4 == max_exp) & outl).any(axis=1, keepdims=True)
    target = np.where(has_max, max_exp, target)
    shift = target - e4
    right = _rounding_rshift(m4, np.clip(shift, 0, None), int8=True)
    left = _wrap8(m4 << np.clip(-shift, 0, None))
    cap = (((1 << (mant_bits + 1)) - 1) << tail_pad_bits) if tail_pad_bits >= 0 else (((1 << (mant_bits + 1)) - 1) >> abs(tail_pad_bits))
    left = np.minimum(left, cap)
    aligned = np.where(shift >= 0, right, left)
    # value: D arithmetic (mant_aligned.to(dtype) / python float * pow(dtype))                   :357-358
    hi_unb = rD((target - bias).astype(np.float32), dtype)
    v = rD(rD(aligned.astype(np.float32), dtype) / np.float32(2.0 ** (mant_bits + tail_pad_bits)), dtype)
    v = rD(v * rD(np.exp2(hi_unb), dtype), dtype)
    v = np.where(s4 == 1, -v, v)
    v = np.where(z4, np.float32(0.0), v)
    return np.ascontiguousarray(v.reshape(shape_t).T.astype(np.float32))


def fp_approx_fake_quant(w_bits: np.ndarray, dtype: str, exp_bits: int, mant_bits: int, group: int, hi_align_start: int, hi_align_exp_field: int,
                         tail_pad_bits: int, double: bool):
    """QuantLinear.quantize_weight_approximate for quant_dim = 0 (quant_linear.py:470-632): symmetric absmax
    scale (:492-494), the ordinary encoder (:495), the aligned decoder, `decoded * scales` (:606).  Which
    decoder runs follows the driver: FP4 with one exponent bit never uses the double variant (:496-509).
    Use `along_output_dim(fp_approx_fake_quant, ...)` for quant_dim = 1."""
    if group <= 0:
        raise ValueError("approximate 仅支持分组量化，w_group_size 必须 > 0")
    _, _, _, fmax = fp_format(exp_bits, mant_bits)
    w = from_bits(w_bits, dtype)
    shape = w.shape
    g = _group_view(w, group)
    with np.errstate(invalid="ignore", over="ignore"):
        mx = rD(_clamp(_nanmax(np.abs(g)), lo=1e-5), dtype)
        S = rD(_clamp(rD(_div(mx, np.float32(fmax)), dtype), lo=1e-5), dtype)
        n = _clamp(rD(_div(g, S), dtype), -fmax, fmax)
        code = fp_codec_encode(n, exp_bits, mant_bits, dtype)
        if double and not (exp_bits == 1):
            v = rD(fp_decode_aligned_double(code, hi_align_start, hi_align_exp_field, tail_pad_bits, exp_bits, mant_bits, dtype), dtype)
        else:
            v = fp_decode_aligned(code, hi_align_start, hi_align_exp_field, tail_pad_bits, exp_bits, mant_bits, dtype)
        out = rD(v * S, dtype)
    return dict(out=to_bits(out, dtype).reshape(shape), scales=to_bits(S, dtype).reshape(-1), scales_half=to_bits(S, F16).reshape(-1), codes=code.reshape(shape))
