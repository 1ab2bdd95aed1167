#!/usr/bin/env python
"""Opcode histogram of one kernel's SASS (cuobjdump -sass on an object file of csrc/build).

    python tools/sass_hist.py iron_weight_only_quant_b200/csrc/build/iwq_inst_f16_int_fast.o 'seg4_tma_kernelILi0ELi0ELb0ELb0ELi0E'
"""
import collections
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", out)
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        ops = collections.Counter()
        n = 0
        for line in b.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(2)] += 1
                n += 1
        demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        print(f"== {demangled}\n   {n} SASS instructions")
        fam = collections.Counter()
        for op, c in ops.items():
            fam[op.split(".")[0]] += c
        print("   by mnemonic:", ", ".join(f"{k} {v}" for k, v in fam.most_common()))
        keys = ["UBLKCP", "UTMALDG", "UTMASTG", "UTMACMDFLUSH", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "FFMA2", "FMUL2", "FADD2", "HFMA2", "HADD2", "HMNMX2", "HMUL2", "SHFL", "PRMT", "MUFU", "BAR"]
        print("   of interest:")
        for k in keys:
            hits = {op: c for op, c in ops.items() if op.split(".")[0] == k}
            if hits:
                print(f"     {k:14s} {sum(hits.values()):4d}   " + ", ".join(f"{o} x{c}" for o, c in sorted(hits.items(), key=lambda x: -x[1])[:6]))
        return
    print("no function matching", pat)
    for b in blocks[1:]:
        print("  ", b.split("\n", 1)[0].strip()[:150])


if __name__ == "__main__":
    main()
