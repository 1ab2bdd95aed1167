"""Drop-in for the reference's quant_wrapper.py: `quantize_model(model, args, quant_mix_gate=False)`.

Same gate, same module filter, same attributes read from `args`, same exceptions
(quant_wrapper.py:15-25, :52-73).  Different on purpose: instead of one Python-level sequence of ATen
launches plus `torch.cuda.empty_cache()` per Linear (quant_wrapper.py:52-82), all Linear weights that
live on the same device (and share a segment shape) are fake-quantised by ONE persistent work-list
launch (`iwq_fake_quant_batched`), so host overhead stays negligible next to ~5 ms of GPU time per
device for a 70B model.
"""
from __future__ import annotations

from collections import defaultdict

import torch

from . import _lib
from .quant_funcs import pseudo_quantize_tensor  # noqa: F401  (imported by the reference too, quant_wrapper.py:4)
from .quant_linear import GPTQQuantLinear, QuantLinear  # noqa: F401


def _is_target(name: str, module) -> bool:
    """quant_wrapper.py:53 -- every nn.Linear except lm_head / output_layer."""
    return isinstance(module, torch.nn.Linear) and "lm_head" not in name and "output_layer" not in name


def _set_submodule(model, name: str, new):
    parent = model
    path = name.split(".")
    for part in path[:-1]:
        parent = getattr(parent, part)
    setattr(parent, path[-1], new)  # quant_wrapper.py:76-80


class _Batcher:
    """Streams QuantLinear shells into batched launches.

    Shells are grouped by (device, dtype[, cols for per-channel]); a group is flushed -- ONE launch for all its
    matrices (two for per-tensor granularity) -- as soon as it holds `threshold` bytes of weights.  The threshold
    starts small and doubles after every flush (128 MB -> 2 GB), so the GPU starts working after the first module
    has been prepared and then always has a launch queued while Python prepares the next modules: the walker's
    host time (~50 us per module) overlaps the GPU time instead of preceding it.
    """

    FIRST, LAST = 128 << 20, 2 << 30

    def __init__(self):
        self.groups = defaultdict(list)
        self.bytes = defaultdict(int)
        self.threshold = defaultdict(lambda: self.FIRST)

    def add(self, ql):
        plan = ql._plan()
        if plan is None:  # w_bit >= 16: leave untouched (quant_linear.py:887-892)
            ql.quantized.fill_(False)
            return
        cfg, G, wants_zeros = plan
        w = ql.weight.data
        if not w.is_cuda:
            raise RuntimeError("quantize_model: weights must be on CUDA devices (no CPU fallback in iron_weight_only_quant_b200)")
        if not w.is_contiguous() or w.data_ptr() % 16 != 0:
            ql.quantize_weight()  # rare: strided / misaligned storage goes through the single-matrix call
            return
        rows, cols = w.shape
        # one launch = one configuration: the whole iwq_config is part of the key (quantize_linears_batched may be handed
        # shells that differ in bits / group / format), plus cols for per-channel (a batched launch shares the row length)
        cfg_key = tuple(getattr(cfg, f) for f, _ in cfg._fields_)
        key = (w.device, w.dtype, cols if (ql.w_group_size == _lib.GROUP_PER_CHANNEL and cfg.quant_dim == 0) else 0, cfg_key, wants_zeros)
        self.groups[key].append((ql, cfg, w, rows, cols, G, wants_zeros))
        self.bytes[key] += w.numel() * 2
        if self.bytes[key] >= self.threshold[key]:
            self.flush(key)
            self.threshold[key] = min(self.threshold[key] * 2, self.LAST)

    def flush(self, key):
        items = self.groups.pop(key, [])
        self.bytes.pop(key, None)
        if not items:
            return
        device, dtype = key[0], key[1]
        cfg = items[0][1]
        # one allocation for all scales (and one for all zero points) of the launch, handed out as views
        n_seg = sum(it[5] for it in items)
        wants_zeros = items[0][6]
        pool_s = torch.empty(n_seg, 1, dtype=dtype, device=device) if n_seg else None
        pool_z = torch.empty(n_seg, 1, dtype=dtype, device=device) if (n_seg and wants_zeros) else None
        mats, outs, off = [], [], 0
        for ql, _, w, rows, cols, G, _ in items:
            scales = pool_s[off : off + G] if G else None
            zeros = pool_z[off : off + G] if (G and pool_z is not None) else None
            off += G
            mats.append(_lib.matrix_desc(w, w, rows, cols, scales, zeros))
            outs.append((ql, scales, zeros))
        if cfg.group == _lib.GROUP_PER_TENSOR:
            rng = torch.empty(2 * len(mats), dtype=torch.float32, device=device)
            _lib.per_tensor_batched_(mats, cfg, device, rng, track_status=False)
        else:
            _lib.fake_quant_batched_(mats, cfg, device, track_status=False)
        for ql, scales, zeros in outs:
            ql._finish(scales, zeros)

    def flush_all(self):
        for key in list(self.groups):
            self.flush(key)


def quantize_linears_batched(shells):
    """Fake-quantise the weights of already-built QuantLinear shells with as few launches as possible
    (`shells`: QuantLinear modules whose `.weight` aliases the original storage)."""
    b = _Batcher()
    for ql in shells:
        b.add(ql)
    b.flush_all()


def quantize_model(model, args, quant_mix_gate=False):
    """Weight-only fake-quantisation of every nn.Linear except lm_head/output_layer (quant_wrapper.py:7-84)."""
    if (args.w_bit is not None and args.w_bit < 16) and (args.a_bit is None or args.a_bit >= 16):
        assert args.w_bit > 0 and args.w_bit < 16, "Weight bitwidth should be an integer between [1, 16] for weigth-only quantization, please check."
        w_format = getattr(args, "w_format", "int").lower()
        is_bfp = w_format.startswith("bfp")
        if is_bfp and getattr(args, "w_group_size", -1) <= 0:
            raise ValueError("BFP 量化仅支持分组量化，请设置正的 w_group_size（例如128）")  # quant_wrapper.py:19-20

        if hasattr(args, "gptq") and args.gptq:
            if w_format != "int":
                raise NotImplementedError("GPTQ 路径目前仅支持整数量化权重")  # quant_wrapper.py:24-25
            raise NotImplementedError(
                "GPTQ calibration (Hessian/Cholesky, quant_wrapper.py:26-28 -> gptq_utils.apply_gptq) is outside the "
                "bandwidth-bound fake-quant path this package implements"
            )

        shells, names = [], []
        batcher = _Batcher()
        for name, module in list(model.named_modules()):
            if _is_target(name, module):
                ql = QuantLinear._shell_from_linear(
                    module,
                    w_bit=args.w_bit,
                    w_group_size=args.w_group_size,
                    symmetric=args.w_symmetric,
                    mode=getattr(args, "mode", 0),
                    weight_format=w_format,
                    approximate=getattr(args, "approximate", False),
                    quant_dim=getattr(args, "quant_dim", 0),
                    fp8_hi_align_start=getattr(args, "fp8_hi_align_start", 12),
                    fp8_hi_align_exp_field=getattr(args, "fp8_hi_align_exp_field", 15),
                    fp8_tail_pad_bits=getattr(args, "fp8_tail_pad_bits", 1),
                    double_approximate=getattr(args, "double_approximate", False),
                    fp6_hi_align_start=getattr(args, "fp6_hi_align_start", 4),
                    fp6_hi_align_exp_field=getattr(args, "fp6_hi_align_exp_field", 7),
                    fp6_tail_pad_bits=getattr(args, "fp6_tail_pad_bits", 2),
                    fp4_hi_align_start=getattr(args, "fp4_hi_align_start", 1),
                    fp4_hi_align_exp_field=getattr(args, "fp4_hi_align_exp_field", 1),
                    fp4_tail_pad_bits=getattr(args, "fp4_tail_pad_bits", 0),
                )
                batcher.add(ql)  # launches as soon as enough weights are queued; the GPU works while we walk on
                shells.append(ql)
                names.append(name)
        batcher.flush_all()
        for name, ql in zip(names, shells):
            _set_submodule(model, name, ql)
    return model
