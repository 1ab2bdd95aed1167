"""Top-level shim so that the reference's scripts (`main.py`, `eval_quant_lm_eval.py`, `demo_activation.py`,
`visualize_utils.py`) import the B200 implementation unchanged: put this directory (and the repo root) in
front of the reference checkout on PYTHONPATH -- see INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from iron_weight_only_quant_b200.fp4_quantize_cpu import *  # noqa: F401,F403,E402
from iron_weight_only_quant_b200 import fp4_quantize_cpu as _impl  # noqa: E402

from iron_weight_only_quant_b200.fp4_quantize_cpu import _fp_scale, quantize_fp16_to_fp4_e1m2, quantize_fp16_to_minifloat  # noqa: F401,E402
