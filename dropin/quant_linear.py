"""Top-level shim so that the reference's scripts (`main.py`, `eval_quant_lm_eval.py`, `demo_activation.py`,
`visualize_utils.py`) import the B200 implementation unchanged: put this directory (and the repo root) in
front of the reference checkout on PYTHONPATH -- see INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from iron_weight_only_quant_b200.quant_linear import *  # noqa: F401,F403,E402
from iron_weight_only_quant_b200 import quant_linear as _impl  # noqa: E402

# names the reference's callers import explicitly, including underscore-prefixed ones (visualize_utils.py:6)
from iron_weight_only_quant_b200.quant_linear import (  # noqa: F401,E402
    FP4_EXP_BIAS, FP4_EXP_BITS, FP4_MANTISSA_BITS, FP6_EXP_BIAS, FP6_EXP_BITS, FP6_MANTISSA_BITS, FP8_EXP_BIAS, FP8_EXP_BITS,
    FP8_MANTISSA_BITS, MANTISSA_BITS, GPTQQuantLinear, QuantLinear, _count_fp4_values, _float_to_fp, _fp_to_float, fp_format_of,
)


def configure_fp_formats(*args, **kwargs):
    """Forward to the implementation module (whose globals the kernels read) and mirror the values here,
    because `from quant_linear import FP4_EXP_BITS` style imports bind by value (as in the reference)."""
    _impl.configure_fp_formats(*args, **kwargs)
    g = globals()
    for k in ("FP4_EXP_BITS", "FP4_MANTISSA_BITS", "FP4_EXP_BIAS", "FP6_EXP_BITS", "FP6_MANTISSA_BITS", "FP6_EXP_BIAS", "FP8_EXP_BITS", "FP8_MANTISSA_BITS", "FP8_EXP_BIAS"):
        g[k] = getattr(_impl, k)
