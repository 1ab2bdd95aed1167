"""Drop-in for the reference's quant_funcs.py: same names, positional order, defaults and return types.

`pseudo_quantize_tensor` runs as ONE fused CUDA pass (two for per-tensor granularity) instead of the
reference's ~14 ATen launches (quant_funcs.py:16-40).  Results are bit-identical to the reference for
fp16/bf16 tensors (tests/test_gpu_parity.py).
"""
from __future__ import annotations

import os

import torch

from . import _lib

# "sync": raise AssertionError like quant_funcs.py:40 (one device sync per call, as in the reference);
# "off": never read the status word (no sync).
NAN_CHECK = os.environ.get("IWQ_NAN_CHECK", "sync")


@torch.no_grad()
def pseudo_quantize_tensor(tensor, n_bits=8, zero_point=True, q_group_size=-1, per_tensor=False, inplace=False):
    """INT min-max quantize->dequantize (reference quant_funcs.py:4-46).

    Grouping (quant_funcs.py:10-15): q_group_size > 0 -> contiguous runs of that many elements along the
    last dim; per_tensor -> one range for the whole tensor; otherwise one range per row of a 2-D tensor.
    inplace=True overwrites the caller's storage and returns a view of it; inplace=False returns a new tensor.
    """
    org_tensor_shape = tensor.shape
    if q_group_size > 0:
        assert org_tensor_shape[-1] % q_group_size == 0  # quant_funcs.py:11
    if per_tensor:
        group = _lib.GROUP_PER_TENSOR
    elif q_group_size > 0:
        group = int(q_group_size)
    else:
        assert tensor.dim() == 2  # quant_funcs.py:15
        group = _lib.GROUP_PER_CHANNEL
    if not tensor.is_cuda:
        raise RuntimeError("pseudo_quantize_tensor: tensor must be on a CUDA device (no CPU fallback in iron_weight_only_quant_b200)")
    src = tensor if tensor.is_contiguous() else tensor.contiguous()
    cols = int(org_tensor_shape[-1]) if tensor.dim() > 0 else 1
    rows = src.numel() // max(cols, 1)
    if src.numel() == 0:
        return tensor if inplace else tensor.clone()
    out = src if inplace else torch.empty_like(src)
    cfg = _lib.make_config(src.dtype, _lib.SCHEME_INT, group, bits=n_bits, symmetric=not zero_point)
    check = NAN_CHECK == "sync"
    _lib.fake_quant_(src, out, cfg, rows, cols, track_status=check)
    if check:
        assert _lib.read_and_clear_status(src.device) & _lib.STATUS_NAN == 0  # quant_funcs.py:40
    return out.reshape(org_tensor_shape)


@torch.no_grad()
def quantize_weight_per_channel_absmax(w, n_bits=8):
    """reference quant_funcs.py:49-55: symmetric, one range per row."""
    return pseudo_quantize_tensor(w, n_bits=n_bits, zero_point=False, q_group_size=-1, per_tensor=False, inplace=False)


@torch.no_grad()
def quantize_activation_per_token_absmax(t, n_bits=8):
    """reference quant_funcs.py:57-62: despite the name, ASYMMETRIC per token (zero_point=True at :61)."""
    t_shape = t.shape
    t = t.view(-1, t_shape[-1])
    t = pseudo_quantize_tensor(t, n_bits=n_bits, zero_point=True, q_group_size=-1, per_tensor=False, inplace=False)
    return t.reshape(t_shape)


@torch.no_grad()
def quantize_weight_per_tensor_absmax(w, n_bits=8):
    """reference quant_funcs.py:64-70: symmetric, one range for the tensor."""
    return pseudo_quantize_tensor(w, n_bits=n_bits, zero_point=False, q_group_size=-1, per_tensor=True, inplace=False)


@torch.no_grad()
def quantize_activation_per_tensor_absmax(t, n_bits=8):
    """reference quant_funcs.py:72-77: asymmetric, one range for the tensor."""
    t_shape = t.shape
    t = t.view(-1, t_shape[-1])
    t = pseudo_quantize_tensor(t, n_bits=n_bits, zero_point=True, q_group_size=-1, per_tensor=True, inplace=False)
    return t.reshape(t_shape)


@torch.no_grad()
def quantize_to_packed_w4(tensor, n_bits=4, zero_point=True, q_group_size=128):
    """NOT in the reference (which computes its codes only to throw them away: quant_funcs.py:44-46,
    quant_linear.py:613-625) -- SURVEY section 8f row 4: the same fused pass that fake-quantises also emits the
    integer codes in a packed W4 layout, two per byte with element 2i in the low nibble (two's complement nibbles
    for the symmetric form), plus the scales and zero points `pseudo_quantize_tensor` drops.

    Returns {"codes": uint8 [rows, cols/2], "scales": [G, 1], "zeros": [G, 1] or None, "dequantized": tensor},
    G = rows*cols/q_group_size (q_group_size > 0) or rows (q_group_size = -2).  The codes are the reference's
    `clamp(round(w / s) + z)` values bit for bit (checked in tests/test_gpu_parity.py).
    """
    if not tensor.is_cuda:
        raise RuntimeError("quantize_to_packed_w4: tensor must be on a CUDA device (no CPU fallback in iron_weight_only_quant_b200)")
    assert tensor.dim() == 2 and n_bits <= 4
    rows, cols = (int(x) for x in tensor.shape)
    assert cols % 2 == 0
    group = int(q_group_size) if q_group_size > 0 else _lib.GROUP_PER_CHANNEL
    if q_group_size > 0:
        assert cols % q_group_size == 0
    src = tensor if tensor.is_contiguous() else tensor.contiguous()
    G = _lib.num_segments(rows, cols, group)
    out = torch.empty_like(src)
    codes = torch.empty(rows, cols // 2, dtype=torch.uint8, device=src.device)
    scales = torch.empty(G, 1, dtype=src.dtype, device=src.device)
    zeros = torch.empty(G, 1, dtype=src.dtype, device=src.device) if zero_point else None
    cfg = _lib.make_config(src.dtype, _lib.SCHEME_INT, group, bits=n_bits, symmetric=not zero_point, code_bits=4)
    _lib.fake_quant_(src, out, cfg, rows, cols, scales=scales, zeros=zeros, codes=codes, track_status=False)
    return {"codes": codes, "scales": scales, "zeros": zeros, "dequantized": out}
