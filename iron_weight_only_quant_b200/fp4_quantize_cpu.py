"""Drop-in for the reference's fp4_quantize_cpu.py (ideal round-to-nearest-even minifloat emulation).

The module keeps the reference's (misleading) names: `quantize_fp16_to_fp4_e1m2` is hard-coded to
E=2, M=1 in the reference (fp4_quantize_cpu.py:61-62) and so is this one; `quantize_fp16_to_minifloat`
is the same path with (E, M) exposed.  Despite "cpu" in the file name the arithmetic runs on the GPU;
tensors must be CUDA tensors.
"""
from __future__ import annotations

import torch

from . import _lib


def _run_rne(tensor: torch.Tensor, group: int, exp_bits: int, mant_bits: int) -> torch.Tensor:
    rows, cols = tensor.shape
    out = torch.empty_like(tensor)
    cfg = _lib.make_config(torch.float16, _lib.SCHEME_FP_RNE, group, exp_bits=exp_bits, mant_bits=mant_bits, symmetric=True)
    _lib.fake_quant_(tensor, out, cfg, rows, cols, track_status=False)
    return out


def quantize_fp16_to_minifloat(tensor, exp_bits, mant_bits, group_size=128, per_tensor=False):
    """absmax-scaled RNE onto the (E, M) minifloat grid, saturating (fp4_quantize_cpu.py:37-72 with the
    hard-coded E/M made parameters).  Returns fp16 shaped like the reference: `(-1, group_size)`,
    `(1, N)` for per_tensor, the input shape when group_size <= 0."""
    if not tensor.is_cuda:
        raise RuntimeError("tensor must be on a CUDA device (no CPU fallback in iron_weight_only_quant_b200)")
    if tensor.dtype != torch.float16:
        tensor = tensor.to(torch.float16)  # fp4_quantize_cpu.py:48-49
    if tensor.dim() != 2:
        raise ValueError("Expected a 2D tensor of shape [out_features, in_features].")
    org_shape = tensor.shape
    if group_size > 0 and org_shape[1] % group_size != 0:
        raise ValueError("in_features must be divisible by group_size.")
    tensor = tensor.contiguous()
    if per_tensor:
        return _run_rne(tensor, _lib.GROUP_PER_TENSOR, exp_bits, mant_bits).reshape(1, -1)
    if group_size > 0:
        return _run_rne(tensor, int(group_size), exp_bits, mant_bits).reshape(-1, group_size)
    return _run_rne(tensor, _lib.GROUP_PER_CHANNEL, exp_bits, mant_bits)


def quantize_fp16_to_fp4_e1m2(tensor, group_size=128, per_tensor=False, return_scales=False):
    """reference fp4_quantize_cpu.py:47-72.  `return_scales` is accepted and ignored, as in the reference."""
    return quantize_fp16_to_minifloat(tensor, 2, 1, group_size=group_size, per_tensor=per_tensor)


def _fp_scale(tensor, S, M, bias, max_float, min_float):
    """reference fp4_quantize_cpu.py:37-44: `tensor / S` clamped to +-max_float and rounded (RNE, per-binade scale with the
    biased exponent clamped to >= 1) onto the (E, M) minifloat grid whose (M, bias, max_float) these are; returns
    `tensor_q`, NOT multiplied back by S.  `S` is anything that broadcasts against `tensor` (the reference passes the
    per-row `absmax / max_float`, shape [rows, 1]); fp16 CUDA tensors.  min_float must be -max_float (as in every call of
    the reference) and the triple must belong to a minifloat format with bias = 2^(E-1)-1."""
    E = None
    for e in range(1, 6):
        if 2 ** (e - 1) - 1 == bias and abs((2 - 2 ** (-M)) * 2 ** (2 ** e - 1 - bias) - max_float) < 1e-6:
            E = e
    if E is None or min_float != -max_float:
        raise NotImplementedError("_fp_scale: only (M, bias, max_float) triples of an (E, M) minifloat with min_float = -max_float")
    if tensor.dtype != torch.float16 or not tensor.is_cuda:
        raise RuntimeError("_fp_scale: fp16 CUDA tensor expected")
    if not torch.is_tensor(S):
        S = torch.tensor(float(S), dtype=torch.float16, device=tensor.device)
    S = S.to(device=tensor.device, dtype=torch.float16)
    shape = torch.broadcast_shapes(tensor.shape, S.shape)
    x = tensor.expand(shape).contiguous().reshape(-1)
    s = S.expand(shape).contiguous().reshape(-1)
    pad = x.numel() % 2
    if pad:
        x = torch.cat([x, x.new_zeros(1)])
        s = torch.cat([s, s.new_ones(1)])
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        rc = _lib.lib().iwq_fp_scale(E, int(M), x.data_ptr(), s.data_ptr(), y.data_ptr(), x.numel(), torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc)
    if pad:
        y = y[:-1]
    return y.reshape(shape)
