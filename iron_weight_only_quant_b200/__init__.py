"""B200-native (sm_100a) weight fake-quantisation: a from-scratch implementation of the hot path of
LiuTielong/Iron_weight_only_quant behind the reference's own Python signatures.

Modules mirror the reference's file names so callers can switch imports one-for-one:
  quant_funcs      pseudo_quantize_tensor and the four thin wrappers      (reference quant_funcs.py)
  fp4_quantize_cpu quantize_fp16_to_fp4_e1m2 / _fp_scale                  (reference fp4_quantize_cpu.py)
  quant_linear     QuantLinear, GPTQQuantLinear, configure_fp_formats     (reference quant_linear.py)
  quant_wrapper    quantize_model                                         (reference quant_wrapper.py)
plus `sharded` (per-tensor granularity on matrices split across GPUs, one NCCL MAX all-reduce) and
`shapes` (LLaMA/OPT Linear shapes for synthetic workloads).  All arithmetic runs in the CUDA kernels
of libiwq_b200.so; nothing here computes on the CPU.
"""
from . import _lib  # noqa: F401

__all__ = ["quant_funcs", "fp4_quantize_cpu", "quant_linear", "quant_wrapper", "sharded", "shapes"]
__version__ = "0.1.0"
