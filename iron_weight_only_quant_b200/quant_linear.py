"""Drop-in for the reference's quant_linear.py on the weight fake-quant path.

Kept from the reference (names, argument order, defaults, buffers, error behaviour):
  QuantLinear(...)                      quant_linear.py:395-468
  QuantLinear.from_linear(...)          quant_linear.py:974-1033  (aliases the Linear's weight storage and
                                                                   overwrites it in place, :1021-1029, :949)
  QuantLinear.quantize_weight()         quant_linear.py:635-958   INT (:885-956) and FP4/FP6/FP8 (:724-883)
  QuantLinear.forward(input)            quant_linear.py:960-972
  GPTQQuantLinear                       quant_linear.py:1035-1065
  configure_fp_formats + FP*_ globals   quant_linear.py:7-16, :84-110
  _float_to_fp / _fp_to_float           quant_linear.py:126-163, :213-235 (element-wise codec)

Different on purpose: the arithmetic is ONE fused CUDA pass per matrix (libiwq_b200.so) instead of ~14
(INT) / ~45 (FP) ATen launches; the constructor does not allocate and kaiming-initialise a throw-away
CPU weight when called from `from_linear` (quant_linear.py:444,460-468 costs ~12 ns per weight);
no `torch.cuda.empty_cache()` per layer (:954, :162).

Also implemented: weight_format='bfp' (block floating point, :648-723); quant_dim=1 (segments along the
output dimension, :478-483, :640-647) for every scheme; the single aligned ("approximate") decode
(_fp_decode_aligned, :237-285, driver :470-632) for both quant_dim; the double-approximate decode
(fp_decode_aligned_double_approx, :288-363) for quant_dim=1 with (out_features / group) % 4 == 0 -- the
configuration every script of the reference uses -- and for quant_dim=0 (:484-488) with group 32/64/128/256 and a
number of segments that is a multiple of 4 (otherwise the reference's groups of four wrap around: NotImplementedError).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .quant_funcs import pseudo_quantize_tensor  # noqa: F401  (re-exported like the reference, quant_linear.py:5)

MANTISSA_BITS = 12
FP4_EXP_BITS = 2
FP4_MANTISSA_BITS = 1
FP4_EXP_BIAS = 2 ** (FP4_EXP_BITS - 1) - 1
FP6_EXP_BITS = 3
FP6_MANTISSA_BITS = 2
FP6_EXP_BIAS = 2 ** (FP6_EXP_BITS - 1) - 1
FP8_EXP_BITS = 4
FP8_MANTISSA_BITS = 3
FP8_EXP_BIAS = 2 ** (FP8_EXP_BITS - 1) - 1


def configure_fp_formats(
    fp4_exp_bits: int = FP4_EXP_BITS,
    fp4_mantissa_bits: int = FP4_MANTISSA_BITS,
    fp6_exp_bits: int = FP6_EXP_BITS,
    fp6_mantissa_bits: int = FP6_MANTISSA_BITS,
    fp8_exp_bits: int = FP8_EXP_BITS,
    fp8_mantissa_bits: int = FP8_MANTISSA_BITS,
):
    """Set the (E, M) split of FP4/FP6/FP8 through module globals, exactly like quant_linear.py:84-110
    (the bias follows the exponent width).  Note the defaults are bound at import time, as in the reference."""
    global FP4_EXP_BITS, FP4_MANTISSA_BITS, FP4_EXP_BIAS
    global FP6_EXP_BITS, FP6_MANTISSA_BITS, FP6_EXP_BIAS
    global FP8_EXP_BITS, FP8_MANTISSA_BITS, FP8_EXP_BIAS

    FP4_EXP_BITS = int(fp4_exp_bits)
    FP4_MANTISSA_BITS = int(fp4_mantissa_bits)
    FP4_EXP_BIAS = 2 ** (FP4_EXP_BITS - 1) - 1

    FP6_EXP_BITS = int(fp6_exp_bits)
    FP6_MANTISSA_BITS = int(fp6_mantissa_bits)
    FP6_EXP_BIAS = 2 ** (FP6_EXP_BITS - 1) - 1

    FP8_EXP_BITS = int(fp8_exp_bits)
    FP8_MANTISSA_BITS = int(fp8_mantissa_bits)
    FP8_EXP_BIAS = 2 ** (FP8_EXP_BITS - 1) - 1


def fp_format_of(weight_format: str):
    """Current (E, M) of 'fp4' / 'fp6' / 'fp8' (reads the module globals at call time)."""
    return {
        "fp4": (FP4_EXP_BITS, FP4_MANTISSA_BITS),
        "fp6": (FP6_EXP_BITS, FP6_MANTISSA_BITS),
        "fp8": (FP8_EXP_BITS, FP8_MANTISSA_BITS),
    }[weight_format]


def _float_to_fp(x: torch.Tensor, exp_bits, mant_bits, exp_bias) -> torch.Tensor:
    """Element-wise minifloat encode with subnormals (quant_linear.py:126-163) -> uint8 codes.
    `x` must be an fp16/bf16 CUDA tensor, already clamped to the format's range (as every caller does)."""
    if exp_bias != 2 ** (exp_bits - 1) - 1:
        raise NotImplementedError("only the standard bias 2^(E-1)-1 (what configure_fp_formats produces)")
    return _codec_roundtrip(x, exp_bits, mant_bits)[1]


def _codec_roundtrip(x: torch.Tensor, exp_bits: int, mant_bits: int):
    if not x.is_cuda:
        raise RuntimeError("CUDA tensor expected (no CPU fallback in iron_weight_only_quant_b200)")
    flat = x.contiguous().reshape(-1)
    pad = flat.numel() % 2
    if pad:
        flat = torch.cat([flat, flat.new_zeros(1)])
    y = torch.empty_like(flat)
    codes = torch.empty(flat.numel(), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.lib().iwq_fp_roundtrip(
            _lib.dtype_code(x.dtype), _lib.SCHEME_FP_CODEC, int(exp_bits), int(mant_bits), flat.data_ptr(), y.data_ptr(), codes.data_ptr(), flat.numel(), torch.cuda.current_stream(x.device).cuda_stream
        )
    _lib.check(rc)
    n = x.numel()
    return y[:n].reshape(x.shape), codes[:n].reshape(x.shape)


def _fp_to_float(code: torch.Tensor, exp_bits: int = FP4_EXP_BITS, mant_bits: int = FP4_MANTISSA_BITS, exp_bias: int = FP4_EXP_BIAS) -> torch.Tensor:
    """Decode minifloat codes to fp32 (quant_linear.py:213-235).  Not on the hot path (the fused kernel
    never materialises codes); provided for the visualisation helpers: a 256-entry table gather."""
    vals = []
    for c in range(256):
        s = (c >> (exp_bits + mant_bits)) & 1
        e = (c >> mant_bits) & ((1 << exp_bits) - 1)
        m = c & ((1 << mant_bits) - 1)
        v = (m / (1 << mant_bits)) * 2.0 ** (1 - exp_bias) if e == 0 else (1.0 + m / (1 << mant_bits)) * 2.0 ** (e - exp_bias)
        vals.append(0.0 if c == 0 else (-v if s else v))
    lut = torch.tensor(vals, dtype=torch.float32, device=code.device)
    return lut[code.to(torch.uint8).long()]


def _count_fp4_values(code: torch.Tensor, exp_bits: int = FP4_EXP_BITS, mant_bits: int = FP4_MANTISSA_BITS, exp_bias: int = FP4_EXP_BIAS):
    """quant_linear.py:366-384: histogram of decoded values (imported by visualize_utils.py:6)."""
    decoded = _fp_to_float(code, exp_bits=exp_bits, mant_bits=mant_bits, exp_bias=exp_bias).reshape(-1)
    values, counts = torch.unique(decoded, sorted=True, return_counts=True)
    return values, counts


class QuantLinear(nn.Module):
    """Linear layer whose weight has been fake-quantised in place (reference quant_linear.py:395-1033)."""

    def __init__(self, in_features, out_features, bias=True, w_bit=4, w_group_size=128, symmetric=True, mode=0, weight_format: str = "int", approximate: bool = False, quant_dim: int = 0, fp8_hi_align_start: int = 12, fp8_hi_align_exp_field: int = 15, fp8_tail_pad_bits: int = 1, double_approximate: bool = False, fp6_hi_align_start: int = 4, fp6_hi_align_exp_field: int = 7, fp6_tail_pad_bits: int = 2, fp4_hi_align_start: int = 1, fp4_hi_align_exp_field: int = 1, fp4_tail_pad_bits: int = 0, _defer_init: bool = False):
        super().__init__()
        raw_format = weight_format.lower()
        # plain (non-tensor) attributes go straight into the instance dict: nn.Module.__setattr__ would run its
        # Parameter / Module / buffer checks for each of the 20 (70 % of the constructor's time; quantize_model
        # builds 560 of these for a 70B model)
        self.__dict__.update(
            in_features=in_features,
            out_features=out_features,
            w_bit=w_bit,
            w_group_size=w_group_size,
            symmetric=symmetric,
            mode=mode,
            quant_dim=quant_dim,
            weight_format="bfp" if raw_format.startswith("bfp") else raw_format,
            approximate=approximate,
            double_approximate=double_approximate,
            fp8_hi_align_start=fp8_hi_align_start,
            fp8_hi_align_exp_field=fp8_hi_align_exp_field,
            fp8_tail_pad_bits=fp8_tail_pad_bits,
            fp6_hi_align_start=fp6_hi_align_start,
            fp6_hi_align_exp_field=fp6_hi_align_exp_field,
            fp6_tail_pad_bits=fp6_tail_pad_bits,
            fp4_hi_align_start=fp4_hi_align_start,
            fp4_hi_align_exp_field=fp4_hi_align_exp_field,
            fp4_tail_pad_bits=fp4_tail_pad_bits,
        )
        if self.weight_format not in {"int", "fp4", "fp6", "fp8", "bfp"}:
            raise ValueError(f"Unsupported weight_format: {weight_format}")  # quant_linear.py:440-441

        if _defer_init:
            # from_linear replaces both parameters right away (quant_linear.py:1021-1029): skip the
            # reference's throw-away CPU allocation + kaiming_uniform_ (:444, :460-468)
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        else:
            self.weight = nn.Parameter(torch.empty(out_features, in_features))
            if bias:
                self.bias = nn.Parameter(torch.empty(out_features))
            else:
                self.register_parameter("bias", None)

        # registered buffers of the reference, in its order (quant_linear.py:451-458)
        self.register_buffer("quantized", torch.tensor(False))
        self._buffers.update(scales=None, zeros=None, weight_fp4=None, weight_fp6=None, weight_fp8=None, weight_bfp_mantissa=None, weight_bfp_exponent=None)

        if not _defer_init:
            self.reset_parameters()

    def reset_parameters(self):
        """quant_linear.py:462-468"""
        nn.init.kaiming_uniform_(self.weight, a=5**0.5)
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / fan_in**0.5
            nn.init.uniform_(self.bias, -bound, bound)

    # ---- planning (shared with quant_wrapper's batched path) ----------------------------------------
    def _plan(self):
        """Validate the configuration and return (cfg, G, wants_zeros) or None when nothing is to be done
        (w_bit >= 16, quant_linear.py:887-892).  Raises what the reference raises."""
        w = self.weight.data
        rows, cols = w.shape
        g = self.w_group_size
        qd = 1 if self.quant_dim == 1 else 0
        seg_dim = rows if qd == 1 else cols  # the dimension the groups run along (after the reference's transpose)
        if self.approximate:
            # quantize_weight_approximate (quant_linear.py:470-632): grouped, symmetric absmax scale, FP formats only
            if g <= 0:
                raise ValueError("approximate 仅支持分组量化，w_group_size 必须 > 0")  # :475-476
            assert seg_dim % g == 0  # :480 / :485
            if self.weight_format not in ("fp4", "fp6", "fp8"):
                raise NotImplementedError("approximate 目前仅支持 fp4/fp6/fp8")  # :601
            E, M = fp_format_of(self.weight_format)
            hs, hf, tp = {
                "fp4": (self.fp4_hi_align_start, self.fp4_hi_align_exp_field, self.fp4_tail_pad_bits),
                "fp6": (self.fp6_hi_align_start, self.fp6_hi_align_exp_field, self.fp6_tail_pad_bits),
                "fp8": (self.fp8_hi_align_start, self.fp8_hi_align_exp_field, self.fp8_tail_pad_bits),
            }[self.weight_format]
            double = bool(self.double_approximate) and not (self.weight_format == "fp4" and E == 1)  # :496-523
            cfg = _lib.make_config(w.dtype, _lib.SCHEME_FP_CODEC, g, exp_bits=E, mant_bits=M, symmetric=True, quant_dim=qd, approx=2 if double else 1, hi_align_start=hs, hi_align_exp_field=hf, tail_pad_bits=tp)
            return cfg, _lib.num_segments(rows, cols, g, qd), False
        if self.weight_format == "bfp" and g <= 0:
            raise ValueError("BFP 仅支持分组量化，请将 w_group_size 设为正数")  # quant_linear.py:650
        if g > 0:
            assert seg_dim % g == 0  # quant_linear.py:653 / :728 / :897 (on the transposed shape when quant_dim == 1)
        elif g not in (-1, -2):
            raise ValueError("Invalid w_group_size")  # quant_linear.py:737 / :906
        if self.weight_format == "bfp":
            # sign + min(w_bit-1, 11) mantissa bits around the group's max exponent; no scales / zeros buffers
            cfg = _lib.make_config(w.dtype, _lib.SCHEME_BFP, g, bits=self.w_bit, symmetric=True, quant_dim=qd)
            return cfg, 0, False
        if self.weight_format == "int":
            if self.w_bit >= 16:
                return None
            cfg = _lib.make_config(w.dtype, _lib.SCHEME_INT, g, bits=self.w_bit, symmetric=self.symmetric, quant_dim=qd)
        else:
            E, M = fp_format_of(self.weight_format)
            cfg = _lib.make_config(w.dtype, _lib.SCHEME_FP_CODEC, g, exp_bits=E, mant_bits=M, symmetric=self.symmetric, quant_dim=qd)
        return cfg, _lib.num_segments(rows, cols, g, qd), (not self.symmetric)

    def _alloc_qparams(self, G: int, wants_zeros: bool):
        w = self.weight.data
        if G == 0:  # BFP: the reference stores neither scales nor zeros (quant_linear.py:716-717)
            return None, None
        scales = torch.empty(G, 1, dtype=w.dtype, device=w.device)
        zeros = torch.empty(G, 1, dtype=w.dtype, device=w.device) if wants_zeros else None
        return scales, zeros

    def _finish(self, scales, zeros):
        """Store buffers the way the reference does: INT keeps the weight dtype (quant_linear.py:924-932);
        FP formats store `.half()` copies (:761-768, quirk Q1)."""
        b = self._buffers  # registered buffers: written directly (Module.__setattr__ costs ~3 us per assignment)
        if self.weight_format == "bfp":
            b["scales"] = None
            b["zeros"] = None
            b["weight_bfp_mantissa"] = None
            b["weight_bfp_exponent"] = None
        elif self.weight_format == "int":
            b["scales"] = scales
            b["zeros"] = zeros
        else:
            b["scales"] = scales.half()
            b["zeros"] = zeros.half() if zeros is not None else None
        b["weight_fp4"] = None
        b["weight_fp6"] = None
        b["weight_fp8"] = None
        self.quantized.fill_(True)

    def quantize_weight(self):
        """Quantise-dequantise `self.weight` in place (reference quant_linear.py:635-958)."""
        with torch.no_grad():
            plan = self._plan()
            if plan is None:
                self.quantized.fill_(False)
                self.weight_fp4 = None
                self.weight_fp6 = None
                self.weight_fp8 = None
                return
            cfg, G, wants_zeros = plan
            w = self.weight.data
            if not w.is_cuda:
                raise RuntimeError("QuantLinear.quantize_weight: the weight must be on a CUDA device (no CPU fallback)")
            rows, cols = w.shape
            scales, zeros = self._alloc_qparams(G, wants_zeros)
            if w.is_contiguous():
                _lib.fake_quant_(w, w, cfg, rows, cols, scales=scales, zeros=zeros, track_status=False)
            else:
                tmp = w.contiguous()
                _lib.fake_quant_(tmp, tmp, cfg, rows, cols, scales=scales, zeros=zeros, track_status=False)
                w.copy_(tmp)
            self._finish(scales, zeros)
            if self.approximate:
                self.approximate = True  # quant_linear.py:628

    def forward(self, input):
        """quant_linear.py:960-972"""
        if not self.quantized:
            return F.linear(input, self.weight, self.bias)
        original_input_shape = input.shape
        weight = self.weight.to(input.dtype)
        out = F.linear(input, weight, self.bias)
        if input.dim() > 2:
            out = out.reshape(original_input_shape[:-1] + (self.out_features,))
        return out

    @classmethod
    def _shell_from_linear(cls, linear_layer, **kw):
        """Build the module around the Linear's own storage without quantising (used by quantize_model's
        batched path and by from_linear)."""
        assert isinstance(linear_layer, nn.Linear), "Input layer must be nn.Linear"
        ql = cls(in_features=linear_layer.in_features, out_features=linear_layer.out_features, bias=linear_layer.bias is not None, _defer_init=True, **kw)
        ql.weight = nn.Parameter(linear_layer.weight.data.detach(), requires_grad=False)  # quant_linear.py:1021-1024
        if linear_layer.bias is not None:
            ql.bias = nn.Parameter(linear_layer.bias.data.detach(), requires_grad=False)  # :1025-1029
        return ql

    @classmethod
    def from_linear(
        cls,
        linear_layer,
        w_bit=4,
        w_group_size=128,
        symmetric=False,
        mode=0,
        weight_format: str = "int",
        approximate: bool = False,
        quant_dim: int = 0,
        fp8_hi_align_start: int = 12,
        fp8_hi_align_exp_field: int = 15,
        fp8_tail_pad_bits: int = 1,
        double_approximate: bool = False,
        fp6_hi_align_start: int = 4,
        fp6_hi_align_exp_field: int = 7,
        fp6_tail_pad_bits: int = 2,
        fp4_hi_align_start: int = 1,
        fp4_hi_align_exp_field: int = 1,
        fp4_tail_pad_bits: int = 0,
    ):
        """reference quant_linear.py:974-1033 (note: `symmetric` defaults to False here, True in __init__)."""
        ql = cls._shell_from_linear(
            linear_layer,
            w_bit=w_bit,
            w_group_size=w_group_size,
            symmetric=symmetric,
            mode=mode,
            weight_format=weight_format,
            approximate=approximate,
            quant_dim=quant_dim,
            fp8_hi_align_start=fp8_hi_align_start,
            fp8_hi_align_exp_field=fp8_hi_align_exp_field,
            fp8_tail_pad_bits=fp8_tail_pad_bits,
            double_approximate=double_approximate,
            fp6_hi_align_start=fp6_hi_align_start,
            fp6_hi_align_exp_field=fp6_hi_align_exp_field,
            fp6_tail_pad_bits=fp6_tail_pad_bits,
            fp4_hi_align_start=fp4_hi_align_start,
            fp4_hi_align_exp_field=fp4_hi_align_exp_field,
            fp4_tail_pad_bits=fp4_tail_pad_bits,
        )
        ql.quantize_weight()
        return ql


class GPTQQuantLinear(nn.Module):
    """Wrapper for a GPTQ-processed Linear (reference quant_linear.py:1035-1065): plain F.linear on a
    cloned weight, carrying over `scales` / `zeros` attributes when present."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor | None = None):
        super().__init__()
        assert weight.dim() == 2, "Weight tensor must be 2-dimensional"
        self.out_features, self.in_features = weight.shape
        self.weight = nn.Parameter(weight.clone().detach(), requires_grad=False)
        if bias is not None:
            self.bias = nn.Parameter(bias.clone().detach(), requires_grad=False)
        else:
            self.register_parameter("bias", None)

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return F.linear(input, self.weight, self.bias)

    @classmethod
    def from_linear(cls, linear_layer: nn.Linear) -> "GPTQQuantLinear":
        assert isinstance(linear_layer, nn.Linear), "Input layer must be nn.Linear"
        weight = linear_layer.weight.data
        bias = linear_layer.bias.data if linear_layer.bias is not None else None
        layer = cls(weight, bias)
        if hasattr(linear_layer, "scales") and linear_layer.scales is not None:
            layer.scales = linear_layer.scales.clone().detach()
        if hasattr(linear_layer, "zeros") and linear_layer.zeros is not None:
            layer.zeros = linear_layer.zeros.clone().detach()
        return layer
