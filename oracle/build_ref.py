#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the UNMODIFIED reference modules of the fake-quant path, as a git-ignored directory that
travels to the GPU box with the repo snapshot (TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product).

The reference (LiuTielong/Iron_weight_only_quant) is a flat directory of Python scripts with no setup.py / pyproject, so
"installing" it means putting its path modules somewhere importable.  This script copies exactly the four files SURVEY.md
section 8(a) cites -- quant_funcs.py, fp4_quantize_cpu.py, quant_linear.py, quant_wrapper.py -- byte for byte from
/root/reference (or $IWQ_REFERENCE_DIR) into oracle/_ref/ and writes their sha256 next to them.  `oracle/_ref/` is listed
in .gitignore (the history stays free of reference sources) but not in .gpurunignore.

    python oracle/build_ref.py            # no-op with a message when the reference checkout is absent

`load()` below is how bench.py / the tests import them: quant_linear.py needs `matplotlib` (not installed) only for its
plotting helpers, so empty stand-in modules are registered first.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("quant_funcs.py", "fp4_quantize_cpu.py", "quant_linear.py", "quant_wrapper.py")


def build(src: str | None = None) -> bool:
    src = src or os.environ.get("IWQ_REFERENCE_DIR", "/root/reference")
    if not all(os.path.isfile(os.path.join(src, f)) for f in FILES):
        print(f"oracle/build_ref.py: no reference checkout at {src}; keeping oracle/_ref as it is")
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    sums = {}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(REF_DIR, f))
        sums[f] = hashlib.sha256(open(os.path.join(REF_DIR, f), "rb").read()).hexdigest()
    json.dump({"source": src, "sha256": sums}, open(os.path.join(REF_DIR, "MANIFEST.json"), "w"), indent=1)
    return True


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in FILES)


_loaded = {}


def load(name: str = "quant_funcs"):
    """Import `name` ('quant_funcs', 'quant_linear', 'fp4_quantize_cpu', 'quant_wrapper') from oracle/_ref under a private
    module name (so that it can never shadow, or be shadowed by, the product's same-named drop-in modules)."""
    if name in _loaded:
        return _loaded[name]
    if not available():
        raise FileNotFoundError("oracle/_ref is empty: run `python oracle/build_ref.py` where /root/reference exists")
    for stub in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    # the reference modules import each other by their top-level names (quant_linear.py: `from quant_funcs import ...`)
    saved = {m: sys.modules.get(m) for m in ("quant_funcs", "quant_linear", "fp4_quantize_cpu", "quant_wrapper", "gptq_utils")}
    sys.path.insert(0, REF_DIR)
    try:
        for m in saved:
            sys.modules.pop(m, None)
        if name == "quant_wrapper":
            sys.modules["gptq_utils"] = types.ModuleType("gptq_utils")  # GPTQ branch only; not on the path
            sys.modules["gptq_utils"].apply_gptq = None
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REF_DIR)
        for m, old in saved.items():
            cur = sys.modules.pop(m, None)
            if m == name:
                mod = cur
            if old is not None:
                sys.modules[m] = old
    _loaded[name] = mod
    return mod


if __name__ == "__main__":
    ok = build(sys.argv[1] if len(sys.argv) > 1 else None)
    print("oracle/_ref:", "written" if ok else ("present" if available() else "absent"))
