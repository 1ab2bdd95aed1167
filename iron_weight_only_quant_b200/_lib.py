"""ctypes binding of libiwq_b200.so (C ABI: include/iwq_b200.h).

This is the only place the package touches the native library.  There is NO CPU path: if the shared
library is missing, or a tensor is not on a CUDA device, the call raises.  PyTorch is used for device
memory, streams and (elsewhere) torch.distributed only; all arithmetic happens in the CUDA kernels.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IWQ_LIB", os.path.join(_HERE, "libiwq_b200.so"))

IWQ_F16, IWQ_BF16 = 0, 1
SCHEME_INT, SCHEME_FP_CODEC, SCHEME_FP_RNE, SCHEME_BFP = 0, 1, 2, 3
GROUP_PER_TENSOR, GROUP_PER_CHANNEL = -1, -2
STATUS_NAN = 1
STATUS_TIMEOUT = 2

ERR_INVALID, ERR_INDIVISIBLE, ERR_UNSUPPORTED, ERR_CUDA, ERR_ALIGN, ERR_FP16_RANGE = -1, -2, -3, -4, -5, -6

EXPORTS = (
    "iwq_abi_version",
    "iwq_last_error",
    "iwq_workspace_bytes",
    "iwq_workspace_cache",
    "iwq_fake_quant",
    "iwq_fake_quant_batched",
    "iwq_fake_quant_per_tensor_batched",
    "iwq_minmax",
    "iwq_fake_quant_with_range",
    "iwq_fp_roundtrip",
    "iwq_fp_scale",
    "iwq_selftest_division",
    "iwq_fake_quant_host",
    "iwq_fake_quant_host_batched",
    "iwq_fused_workspace_bytes",
    "iwq_debug_fused_trace",
    "iwq_debug_plan",
    "iwq_debug_set_knob",
    "iwq_fake_quant_per_tensor_fused",
)


class Config(ctypes.Structure):
    """struct iwq_config (include/iwq_b200.h)."""

    _fields_ = [(n, ctypes.c_int32) for n in ("dtype", "scheme", "bits", "exp_bits", "mant_bits", "symmetric", "group", "variant", "quant_dim", "approx", "hi_align_start", "hi_align_exp_field", "tail_pad_bits", "code_bits")]


class Matrix(ctypes.Structure):
    """struct iwq_matrix (include/iwq_b200.h)."""

    _fields_ = [
        ("w", ctypes.c_void_p),
        ("out", ctypes.c_void_p),
        ("scales", ctypes.c_void_p),
        ("zeros", ctypes.c_void_p),
        ("codes", ctypes.c_void_p),
        ("rows", ctypes.c_int64),
        ("cols", ctypes.c_int64),
    ]


class Peers(ctypes.Structure):
    """struct iwq_peers (include/iwq_b200.h)."""

    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("slot_ptrs", ctypes.POINTER(ctypes.c_void_p)), ("flag_ptrs", ctypes.POINTER(ctypes.c_void_p)), ("epoch", ctypes.c_uint32), ("spin_limit", ctypes.c_uint32)]


class HostMatrix(ctypes.Structure):
    """struct iwq_host_matrix (include/iwq_b200.h)."""

    _fields_ = [("host_w", ctypes.c_void_p), ("host_out", ctypes.c_void_p), ("rows", ctypes.c_int64), ("cols", ctypes.c_int64), ("host_scales", ctypes.c_void_p), ("host_zeros", ctypes.c_void_p)]


_lib = None
_lib_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """Load the native library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"iron_weight_only_quant_b200: native library {LIB_PATH} not found. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C iron_weight_only_quant_b200/csrc`. "
                "There is no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        L.iwq_abi_version.restype = ctypes.c_int
        L.iwq_last_error.restype = ctypes.c_char_p
        L.iwq_workspace_bytes.restype = ctypes.c_size_t
        L.iwq_workspace_bytes.argtypes = [ctypes.c_int]
        L.iwq_workspace_cache.restype = ctypes.c_int
        L.iwq_workspace_cache.argtypes = [ctypes.c_int]
        L.iwq_fake_quant.restype = ctypes.c_int
        L.iwq_fake_quant.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(Matrix), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.iwq_fake_quant_batched.restype = ctypes.c_int
        L.iwq_fake_quant_batched.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(Matrix), ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        L.iwq_fake_quant_per_tensor_batched.restype = ctypes.c_int
        L.iwq_fake_quant_per_tensor_batched.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(Matrix), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        L.iwq_minmax.restype = ctypes.c_int
        L.iwq_minmax.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(Matrix), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.iwq_fake_quant_with_range.restype = ctypes.c_int
        L.iwq_fake_quant_with_range.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(Matrix), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        L.iwq_fp_roundtrip.restype = ctypes.c_int
        L.iwq_fp_roundtrip.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.iwq_fp_scale.restype = ctypes.c_int
        L.iwq_fp_scale.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.iwq_selftest_division.restype = ctypes.c_int
        L.iwq_selftest_division.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.iwq_fake_quant_host.restype = ctypes.c_int
        L.iwq_fake_quant_host.argtypes = [ctypes.POINTER(Config), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int32)]
        L.iwq_fake_quant_host_batched.restype = ctypes.c_int
        L.iwq_fake_quant_host_batched.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(HostMatrix), ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int32)]
        L.iwq_debug_set_knob.restype = ctypes.c_int
        L.iwq_debug_set_knob.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.iwq_fused_workspace_bytes.restype = ctypes.c_size_t
        L.iwq_fused_workspace_bytes.argtypes = [ctypes.c_int]
        L.iwq_fake_quant_per_tensor_fused.restype = ctypes.c_int
        L.iwq_fake_quant_per_tensor_fused.argtypes = [ctypes.POINTER(Config), ctypes.POINTER(Matrix), ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(Peers), ctypes.c_void_p, ctypes.c_void_p]
        if L.iwq_abi_version() != 4:
            raise RuntimeError("libiwq_b200.so ABI version mismatch")
        # the job-table workspaces handed to the library are the tensors of scratch() below: nobody else writes them
        L.iwq_workspace_cache(1)
        _lib = L
    return _lib


def check(rc: int) -> None:
    """Map a negative return code to the exception class the reference raises for that condition."""
    if rc == 0:
        return
    msg = lib().iwq_last_error().decode("utf-8", "replace")
    if rc == ERR_INDIVISIBLE:
        raise AssertionError(msg)  # quant_funcs.py:11 / quant_linear.py:897 `assert ... % group == 0`
    if rc == ERR_INVALID:
        raise ValueError(msg)  # quant_linear.py:906 "Invalid w_group_size"
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_FP16_RANGE:
        raise RuntimeError(msg)  # torch.clamp on a half tensor with an unrepresentable bound
    raise RuntimeError(f"libiwq_b200 error {rc}: {msg}")


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float16:
        return IWQ_F16
    if dt == torch.bfloat16:
        return IWQ_BF16
    raise TypeError(f"iron_weight_only_quant_b200 computes in fp16/bf16 storage only, got {dt} (no CPU/fp32 fallback)")


def make_config(dtype: torch.dtype, scheme: int, group: int, *, bits: int = 0, exp_bits: int = 0, mant_bits: int = 0, symmetric: bool = False, variant: int = 0, quant_dim: int = 0, approx: int = 0, hi_align_start: int = 0, hi_align_exp_field: int = 0, tail_pad_bits: int = 0, code_bits: int = 0) -> Config:
    return Config(
        dtype=dtype_code(dtype), scheme=scheme, bits=int(bits), exp_bits=int(exp_bits), mant_bits=int(mant_bits), symmetric=int(bool(symmetric)), group=int(group), variant=int(variant),
        quant_dim=int(quant_dim), approx=int(approx), hi_align_start=int(hi_align_start), hi_align_exp_field=int(hi_align_exp_field), tail_pad_bits=int(tail_pad_bits), code_bits=int(code_bits),
    )


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device: iron_weight_only_quant_b200 has no CPU path (got device {t.device})")
    if not t.is_contiguous():
        raise ValueError(f"{what} must be contiguous")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def matrix_desc(w: torch.Tensor, out: torch.Tensor, rows: int, cols: int, scales=None, zeros=None, codes=None) -> Matrix:
    return Matrix(w=_ptr(w), out=_ptr(out), scales=_ptr(scales), zeros=_ptr(zeros), codes=_ptr(codes), rows=rows, cols=cols)


def num_segments(rows: int, cols: int, group: int, quant_dim: int = 0) -> int:
    if group > 0:
        return rows * cols // group
    if group == GROUP_PER_TENSOR:
        return 1
    return cols if quant_dim == 1 else rows


# per-device scratch: status word, per-tensor range, job-table workspace -------------------------------
_scratch = {}


def scratch(device: torch.device, n_jobs: int = 1):
    """(status int32[1], range float32[2*n], workspace uint8[...]) on `device`, cached and grown on demand."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    st = _scratch.get(key)
    need_ws = int(lib().iwq_workspace_bytes(n_jobs))
    if st is None or st[1].numel() < 2 * n_jobs or st[2].numel() < need_ws:
        status = st[0] if st is not None else torch.zeros(1, dtype=torch.int32, device=device)
        rng = torch.empty(max(2 * n_jobs, 2), dtype=torch.float32, device=device)
        ws = torch.empty(max(need_ws, 64), dtype=torch.uint8, device=device)
        if st is not None:
            lib().iwq_workspace_cache(1)  # the old workspace goes back to the allocator: forget what was uploaded to it
        st = (status, rng, ws)
        _scratch[key] = st
    return st


def fake_quant_(w: torch.Tensor, out: torch.Tensor, cfg: Config, rows: int, cols: int, scales=None, zeros=None, codes=None, track_status: bool = True):
    """Enqueue one fused fake-quant pass for one matrix on the current stream of its device."""
    _require_cuda(w, "weight")
    _require_cuda(out, "output")
    with torch.cuda.device(w.device):
        status, rng, _ = scratch(w.device)
        m = matrix_desc(w, out, rows, cols, scales, zeros, codes)
        stream = torch.cuda.current_stream(w.device).cuda_stream
        rc = lib().iwq_fake_quant(ctypes.byref(cfg), ctypes.byref(m), rng.data_ptr(), status.data_ptr() if track_status else None, stream)
    check(rc)
    return status


def fake_quant_batched_(mats: Sequence[Matrix], cfg: Config, device: torch.device, track_status: bool = True):
    """One launch for many matrices on `device` (they must share the config and segment shape)."""
    n = len(mats)
    arr = (Matrix * n)(*mats)
    with torch.cuda.device(device):
        status, _, ws = scratch(device, n)
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = lib().iwq_fake_quant_batched(ctypes.byref(cfg), arr, n, ws.data_ptr(), ws.numel(), status.data_ptr() if track_status else None, stream)
    check(rc)
    return status


def minmax_(mats: Sequence[Matrix], cfg: Config, device: torch.device, rng: torch.Tensor):
    n = len(mats)
    arr = (Matrix * n)(*mats)
    with torch.cuda.device(device):
        _, _, ws = scratch(device, n)
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = lib().iwq_minmax(ctypes.byref(cfg), arr, n, rng.data_ptr(), ws.data_ptr(), ws.numel(), stream)
    check(rc)


def fake_quant_with_range_(mats: Sequence[Matrix], cfg: Config, device: torch.device, rng: torch.Tensor, track_status: bool = True):
    n = len(mats)
    arr = (Matrix * n)(*mats)
    with torch.cuda.device(device):
        status, _, ws = scratch(device, n)
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = lib().iwq_fake_quant_with_range(ctypes.byref(cfg), arr, n, rng.data_ptr(), ws.data_ptr(), ws.numel(), status.data_ptr() if track_status else None, stream)
    check(rc)
    return status


def per_tensor_batched_(mats: Sequence[Matrix], cfg: Config, device: torch.device, rng: torch.Tensor, track_status: bool = True):
    """Per-tensor granularity for many whole matrices on one GPU: range and apply passes ordered by the library."""
    n = len(mats)
    arr = (Matrix * n)(*mats)
    with torch.cuda.device(device):
        status, _, ws = scratch(device, n)
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = lib().iwq_fake_quant_per_tensor_batched(ctypes.byref(cfg), arr, n, rng.data_ptr(), ws.data_ptr(), ws.numel(), status.data_ptr() if track_status else None, stream)
    check(rc)
    return status


def read_and_clear_status(device: torch.device) -> int:
    """Synchronising read of the device status word (then reset it)."""
    status, _, _ = scratch(device)
    v = int(status.item())
    if v:
        status.zero_()
    return v


def fake_quant_host_(tensors: Sequence[torch.Tensor], cfg: Config, device: int = 0, chunk_bytes: int = 32 << 20, scales: Optional[Sequence[torch.Tensor]] = None, zeros: Optional[Sequence[torch.Tensor]] = None, copy_only: bool = False) -> int:
    """Fake-quantise 2-D HOST tensors (pinned for full speed) in place through one H2D -> kernel -> D2H chunk
    pipeline on `device` (iwq_fake_quant_host_batched); synchronises; returns the status word (STATUS_NAN ...).
    `scales` / `zeros`: optional host tensors (one per matrix, G elements of the weight dtype) that receive the
    per-segment scales and zero points.  `copy_only` (diagnostics): the same copies without the kernels."""
    ms = []
    for i, t in enumerate(tensors):
        if t.is_cuda or t.dim() != 2 or not t.is_contiguous():
            raise ValueError("fake_quant_host_: expected contiguous 2-D host tensors")
        dtype_code(t.dtype)
        G = num_segments(t.shape[0], t.shape[1], cfg.group)
        ptrs = []
        for what, lst in (("scales", scales), ("zeros", zeros)):
            m = lst[i] if lst is not None else None
            if m is not None and (m.is_cuda or not m.is_contiguous() or m.dtype != t.dtype or m.numel() < G):
                raise ValueError(f"fake_quant_host_: {what}[{i}] must be a contiguous host tensor of the weight dtype with >= {G} elements")
            ptrs.append(None if m is None else m.data_ptr())
        ms.append(HostMatrix(t.data_ptr(), t.data_ptr(), t.shape[0], t.shape[1], ptrs[0], ptrs[1]))
    arr = (HostMatrix * len(ms))(*ms)
    st = ctypes.c_int32(0)
    if copy_only:
        check(lib().iwq_debug_set_knob(b"host_copy_only", 1))
    try:
        with torch.cuda.device(device):  # the library restores the current device itself; this keeps torch's view in step
            check(lib().iwq_fake_quant_host_batched(ctypes.byref(cfg), arr, len(ms), int(device), int(chunk_bytes), ctypes.byref(st)))
    finally:
        if copy_only:
            lib().iwq_debug_set_knob(b"host_copy_only", 0)
    return int(st.value)


def per_tensor_fused_(mats: Sequence[Matrix], cfg: Config, device: torch.device, peers: Optional[Peers] = None, track_status: bool = True):
    """ONE launch for per-tensor granularity over many matrices (range pass, in-kernel range exchange with the
    peers when the matrices are sharded across GPUs, apply pass from L2): iwq_fake_quant_per_tensor_fused."""
    n = len(mats)
    arr = (Matrix * n)(*mats)
    with torch.cuda.device(device):
        status, _, _ = scratch(device, n)
        need = lib().iwq_fused_workspace_bytes(n)
        key = ("fused_ws", device.index if device.index is not None else torch.cuda.current_device())
        ws = _scratch.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=device)
            _scratch[key] = ws
        stream = torch.cuda.current_stream(device).cuda_stream
        rc = lib().iwq_fake_quant_per_tensor_fused(ctypes.byref(cfg), arr, n, ws.data_ptr(), ws.numel(), ctypes.byref(peers) if peers is not None else None, status.data_ptr() if track_status else None, stream)
    check(rc)
    return status
