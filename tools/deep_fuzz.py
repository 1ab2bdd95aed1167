#!/usr/bin/env python
"""One-off deeper run of the seeded fuzz tests with seeds the committed suite does not use (GPU box):
    python tools/deep_fuzz.py 100 140"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_fuzz as F  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    for fn in (F.test_random_configurations_match_oracle, F.test_random_column_configurations_wide_tiles):
        try:
            fn(seed)
        except Exception:  # noqa: BLE001
            bad += 1
            print(f"FAIL seed {seed} {fn.__name__}")
            traceback.print_exc(limit=2)
print(f"seeds {lo}..{hi - 1}: {bad} failures")
sys.exit(1 if bad else 0)
