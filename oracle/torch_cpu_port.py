"""CPU port of the reference's fake-quant path as eager torch ops -- TEST / BASELINE INFRASTRUCTURE ONLY.

Why it exists: the reference is pure Python on top of ATen; /root/reference does not exist on the GPU
box, so `bench.py`'s cpu_baseline and `--impl reference` legs time THIS port (kind "port"): the same
ATen CPU kernels in the same order as quant_funcs.py:16-38 (INT) and quant_linear.py:744-779 with the
codec of :126-163 / :213-235 (minifloat), multi-threaded by torch exactly like the reference would be.
tests/test_oracle_golden.py pins it against the reference's golden outputs.  Never imported by the
product package.
"""
from __future__ import annotations

import torch


def _grouped(t: torch.Tensor, group: int) -> torch.Tensor:
    if group > 0:
        assert t.shape[-1] % group == 0
        return t.reshape(-1, group)
    if group == -1:
        return t.reshape(1, -1)
    assert t.dim() == 2
    return t


@torch.no_grad()
def int_fake_quant(w: torch.Tensor, n_bits: int, zero_point: bool, group: int) -> torch.Tensor:
    """quant_funcs.py:4-46 as a flat function (out of place)."""
    shape = w.shape
    t = _grouped(w, group)
    if zero_point:
        hi = t.amax(dim=1, keepdim=True)
        lo = t.amin(dim=1, keepdim=True)
        qmax, qmin = 2**n_bits - 1, 0
        s = (hi - lo).clamp(min=1e-5) / qmax
        z = (-torch.round(lo / s)).clamp_(qmin, qmax)
    else:
        hi = t.abs().amax(dim=1, keepdim=True).clamp(min=1e-5)
        qmax, qmin = 2 ** (n_bits - 1) - 1, -(2 ** (n_bits - 1))
        s = hi / qmax
        z = 0
    t = (torch.clamp(torch.round(t / s) + z, qmin, qmax) - z) * s
    assert torch.isnan(t).sum() == 0
    return t.reshape(shape)


def _encode(x: torch.Tensor, E: int, M: int, bias: int) -> torch.Tensor:
    """quant_linear.py:126-163"""
    sign = (x < 0).to(torch.uint8)
    a = x.abs()
    is_zero = a == 0
    a = torch.where(is_zero, torch.full_like(a, 1e-8), a)
    emin = 1 - bias
    e = torch.floor(torch.log2(a)).to(torch.int8)
    sub = e < emin
    ec = torch.clamp(e, emin, (1 << E) - 1 - bias)
    ef = (ec + bias).to(torch.uint8)
    ms = 1 << M
    mn = torch.round((a / torch.pow(2.0, ec.float()) - 1.0) * ms).clamp(0, ms - 1).to(torch.uint8)
    msub = torch.round((a / (2.0**emin)) * ms).clamp(0, ms - 1).to(torch.uint8)
    ef = torch.where(sub, torch.zeros_like(ef), ef)
    mf = torch.where(sub, msub, mn)
    code = (sign << (E + M)) | (ef << M) | mf
    return torch.where(is_zero, torch.zeros_like(code), code)


def _decode(code: torch.Tensor, E: int, M: int, bias: int) -> torch.Tensor:
    """quant_linear.py:213-235"""
    code = code.to(torch.uint8)
    sign = ((code >> (E + M)) & 1).float()
    raw = ((code >> M) & ((1 << E) - 1)).float()
    mf = (code & ((1 << M) - 1)).float()
    vn = (1.0 + mf / float(1 << M)) * torch.pow(2.0, raw - float(bias))
    vs = (mf / float(1 << M)) * (2.0 ** (1.0 - bias))
    v = torch.where(raw == 0, vs, vn)
    v = ((-1.0) ** sign) * v
    return torch.where(code == 0, torch.zeros_like(v), v)


@torch.no_grad()
def fp_codec_fake_quant(w: torch.Tensor, E: int, M: int, symmetric: bool, group: int) -> torch.Tensor:
    """quant_linear.py:724-781 (FP4 branch; the FP6/FP8 branches are the same code)."""
    bias = 2 ** (E - 1) - 1
    fmax = (1.0 + (2**M - 1) / (2**M)) * (2.0 ** ((1 << E) - 1 - bias))
    t = _grouped(w, group)
    if symmetric:
        hi = t.abs().amax(dim=1, keepdim=True).clamp(min=1e-5)
        s = (hi / fmax).clamp(min=1e-5)
        zeros = None
        n = torch.clamp(t / s, min=-fmax, max=fmax)
    else:
        hi = t.amax(dim=1, keepdim=True)
        lo = t.amin(dim=1, keepdim=True)
        zeros = (hi + lo) * 0.5
        s = (((hi - lo) * 0.5).clamp(min=1e-5) / fmax).clamp(min=1e-5)
        n = torch.clamp((t - zeros) / s, min=-fmax, max=fmax)
    out = _decode(_encode(n, E, M, bias), E, M, bias).to(w.dtype) * s
    if zeros is not None:
        out = out + zeros.half().to(out.dtype)
    return out.to(w.dtype).reshape(w.shape)
