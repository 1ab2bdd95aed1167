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
    if tensor.numel() == 0:
        return tensor if inplace else tensor.clone()
    src = tensor if tensor.is_contiguous() else tensor.contiguous()
    if src.data_ptr() % 16 != 0:
        src = src.clone()  # a view at an odd storage offset: the kernels move 16-byte vectors
    cols = int(org_tensor_shape[-1]) if tensor.dim() > 0 else 1
    rows = src.numel() // max(cols, 1)
    out = src if inplace else torch.empty_like(src)
    check = NAN_CHECK == "sync"
    if _kernel_shape_ok(rows, cols, group):
        cfg = _lib.make_config(src.dtype, _lib.SCHEME_INT, group, bits=n_bits, symmetric=not zero_point)
        _lib.fake_quant_(src, out, cfg, rows, cols, track_status=check)
    else:
        _run_odd_shape(src, out, rows, cols, group, n_bits, zero_point, check)
    if check:
        assert _lib.read_and_clear_status(src.device) & _lib.STATUS_NAN == 0  # quant_funcs.py:40
    if inplace and out is not tensor:
        tensor.copy_(out.reshape(org_tensor_shape))  # non-contiguous input: the reference mutates the caller's storage
        return tensor
    return out.reshape(org_tensor_shape)


_MAX_ROW = 51200  # longest segment the one-pass row kernel stages in shared memory


def _kernel_shape_ok(rows: int, cols: int, group: int) -> bool:
    """Shapes the vectorised kernels take directly: every segment a whole number of 16-byte vectors."""
    if group == _lib.GROUP_PER_TENSOR:
        return (rows * cols) % 8 == 0
    seg = cols if group == _lib.GROUP_PER_CHANNEL else group
    return seg % 8 == 0 and (group > 0 or seg <= _MAX_ROW)


def _run_odd_shape(src, out, rows, cols, group, n_bits, zero_point, check):
    """The reference takes ANY shape (quant_funcs.py:10-15).  Segments that are not a whole number of 16-byte vectors
    are copied into a padded [segments, seg_pad] buffer whose pad columns repeat each segment's first element (which
    changes neither its min nor its max), quantised there by the same kernels and copied back; per-channel rows too long
    for the one-pass row kernel go through the two-pass per-tensor kernels, one job per row."""
    dev, dt = src.device, src.dtype
    if group == _lib.GROUP_PER_TENSOR:
        segs = src.reshape(1, -1)
    elif group == _lib.GROUP_PER_CHANNEL:
        segs = src.reshape(rows, cols)
    else:
        segs = src.reshape(-1, group)
    n_seg, seg = segs.shape
    pad = (-seg) % 8
    work = torch.cat([segs, segs[:, :1].expand(n_seg, pad)], dim=1).contiguous() if pad else segs.clone()
    seg_pad = seg + pad
    if group == _lib.GROUP_PER_TENSOR or seg_pad > _MAX_ROW:
        cfg = _lib.make_config(dt, _lib.SCHEME_INT, _lib.GROUP_PER_TENSOR, bits=n_bits, symmetric=not zero_point)
        mats = [_lib.matrix_desc(work[i], work[i], 1, seg_pad) for i in range(n_seg)]
        rng = torch.empty(2 * n_seg, dtype=torch.float32, device=dev)
        _lib.per_tensor_batched_(mats, cfg, dev, rng, track_status=check)
    else:
        cfg = _lib.make_config(dt, _lib.SCHEME_INT, _lib.GROUP_PER_CHANNEL, bits=n_bits, symmetric=not zero_point)
        _lib.fake_quant_(work, work, cfg, n_seg, seg_pad, track_status=check)
    out.reshape(n_seg, seg).copy_(work[:, :seg])


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
