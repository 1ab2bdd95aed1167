#!/usr/bin/env python
"""Mint the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF on CPU.

Run in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own modules unmodified (`quant_funcs`, `fp4_quantize_cpu`, `quant_linear`;
`matplotlib` is stubbed because it is not installed and only used for plotting) and stores inputs and
outputs as uint16/uint8 bit patterns in compressed .npz files.  Nothing here is product code; the
fixtures pin `oracle/fakequant_oracle.py` (tests/test_oracle_golden.py) and the CUDA path
(tests/test_gpu_parity.py).  Recorded with torch 2.11.0+cu128 (CPU kernels), 8 threads.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("IWQ_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)

import fp4_quantize_cpu  # noqa: E402
import quant_funcs  # noqa: E402
import quant_linear  # noqa: E402

TORCH_DT = {"f16": torch.float16, "bf16": torch.bfloat16}
INT_BITS = [2, 3, 4, 5, 6, 8]
GROUPS = [32, 64, 128, 256, -1, -2]
FP_FORMATS = {  # name -> (weight_format, E, M)
    "e2m1": ("fp4", 2, 1),
    "e1m2": ("fp4", 1, 2),
    "e3m2": ("fp6", 3, 2),
    "e2m3": ("fp6", 2, 3),
    "e4m3": ("fp8", 4, 3),
    "e5m2": ("fp8", 5, 2),  # bf16 only: fmax overflows fp16 and the reference raises
}


def bits_of(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def build_input(seed: int = 0, rows: int = 32, cols: int = 512) -> torch.Tensor:
    """fp32 master matrix: random rows plus the adversarial rows of SURVEY.md section 8c(3)."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(rows, cols, generator=g) * 0.02
    w[16] = 0.0113
    w[17] = 0.0
    w[18] = torch.randn(cols, generator=g).abs() * 0.05 + 0.01
    w[19] = -(torch.randn(cols, generator=g).abs() * 0.05 + 0.01)
    w[20] = torch.randn(cols, generator=g) * 3000.0
    w[20, 7] = 30000.0
    w[21] = torch.randn(cols, generator=g) * 1e-6
    ties = (torch.arange(cols) % 16).float() - 7.5  # -7.5 .. 7.5 -> w/s lands on x.5 for w4 asym
    w[22] = ties * 2.0 ** -6
    w[23] = torch.randn(cols, generator=g) * 0.02
    w[23, 100] = 2.0
    w[23, 300] = -1.5
    w[24] = torch.randn(cols, generator=g) * 0.02
    w[24, ::3] = 0.0
    w[24, 1::6] = -0.0
    w[25] = torch.randn(cols, generator=g) * 0.02 + 1.0
    for r, sc in zip(range(26, 32), (1e-3, 1e-1, 1.0, 10.0, 100.0, 1e-4)):
        w[r] = torch.randn(cols, generator=g) * sc
    return w


def gen_int(out: dict):
    master = build_input()
    for dn, dt in TORCH_DT.items():
        w = master.to(dt)
        out[f"w_{dn}"] = bits_of(w)
        for bits in INT_BITS:
            for sym in (False, True):
                for grp in GROUPS:
                    key = f"{dn}_b{bits}_{'sym' if sym else 'asym'}_g{grp}"
                    # function form (quant_funcs.py:4)
                    kw = dict(n_bits=bits, zero_point=not sym, inplace=False)
                    if grp > 0:
                        kw.update(q_group_size=grp)
                    elif grp == -1:
                        kw.update(per_tensor=True)
                    try:
                        o = quant_funcs.pseudo_quantize_tensor(w.clone(), **kw)
                        out[f"fn_{key}"] = bits_of(o)
                    except AssertionError:
                        out[f"fn_{key}_assert"] = np.array([1], dtype=np.uint8)
                    # module form (quant_linear.py:885-956)
                    lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False)
                    lin.weight.data = w.clone()
                    ql = quant_linear.QuantLinear.from_linear(lin, w_bit=bits, w_group_size=grp, symmetric=sym, weight_format="int")
                    out[f"ql_{key}_out"] = bits_of(ql.weight.data)
                    out[f"ql_{key}_scales"] = bits_of(ql.scales.reshape(-1))
                    if ql.zeros is not None:
                        out[f"ql_{key}_zeros"] = bits_of(ql.zeros.reshape(-1))
    # in-place form equals out-of-place (SURVEY 8a edge semantics)
    w = master.to(torch.float16)
    t = w.clone()
    r = quant_funcs.pseudo_quantize_tensor(t, n_bits=4, zero_point=True, q_group_size=128, inplace=True)
    out["fn_inplace_f16_b4_asym_g128"] = bits_of(r)
    out["fn_inplace_mutated_input"] = bits_of(t)


def gen_fp_chain(out: dict):
    master = build_input(seed=1)
    for dn, dt in TORCH_DT.items():
        w = master.to(dt)
        out[f"w_{dn}"] = bits_of(w)
        for fname, (wf, E, M) in FP_FORMATS.items():
            if dn == "f16" and E >= 5:
                continue
            kw = {f"{wf}_exp_bits": E, f"{wf}_mantissa_bits": M}
            quant_linear.configure_fp_formats(**kw)
            for sym in (False, True):
                for grp in (32, 128, -1, -2):
                    key = f"{dn}_{fname}_{'sym' if sym else 'asym'}_g{grp}"
                    lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False)
                    lin.weight.data = w.clone()
                    ql = quant_linear.QuantLinear.from_linear(lin, w_bit=4, w_group_size=grp, symmetric=sym, weight_format=wf)
                    out[f"ql_{key}_out"] = bits_of(ql.weight.data)
                    out[f"ql_{key}_scales_half"] = bits_of(ql.scales.reshape(-1))
                    if ql.zeros is not None:
                        out[f"ql_{key}_zeros_half"] = bits_of(ql.zeros.reshape(-1))
        quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)


def gen_fp_tables(out: dict):
    allbits = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16)
    for dn, dt in TORCH_DT.items():
        x = allbits.view(dt)
        for fname, (_, E, M) in FP_FORMATS.items():
            bias = 2 ** (E - 1) - 1
            code = quant_linear._float_to_fp(x, E, M, bias)
            val = quant_linear._fp_to_float(code, E, M, bias)
            out[f"codec_{dn}_{fname}_code"] = code.numpy().astype(np.uint8)
            out[f"codec_{dn}_{fname}_val"] = bits_of(val.to(dt))
    # _fp_scale with S = 1 (fp4_quantize_cpu.py:37-44), fp16 only (the function casts to fp16)
    x = allbits.view(torch.float16)
    for fname, (_, E, M) in FP_FORMATS.items():
        if E >= 5:
            continue
        bias = 2 ** (E - 1) - 1
        mf = (2 - 2 ** (-M)) * 2 ** (2 ** E - 1 - bias)
        S = torch.ones(1, 1, dtype=torch.float16)
        q = fp4_quantize_cpu._fp_scale(x.reshape(1, -1), S, M, bias, mf, -mf)
        out[f"fpscale_f16_{fname}"] = bits_of(q.reshape(-1))


def gen_fp_rne(out: dict):
    master = build_input(seed=2)
    w = master.to(torch.float16)
    out["w_f16"] = bits_of(w)
    for grp in (32, 128, 512):
        o = fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(w.clone(), group_size=grp)
        out[f"e2m1_g{grp}"] = bits_of(o)
        out[f"e2m1_g{grp}_shape"] = np.array(o.shape, dtype=np.int64)
    o = fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(w.clone(), group_size=-1, per_tensor=True)
    out["e2m1_pt"] = bits_of(o)
    out["e2m1_pt_shape"] = np.array(o.shape, dtype=np.int64)
    # the reference's only "known-answer" print: fp4_quantize_cpu.py:74-80 (seed 0, randn(64,128))
    torch.manual_seed(0)
    wm = torch.randn(64, 128, dtype=torch.float16)
    out["main_w"] = bits_of(wm)
    out["main_out"] = bits_of(fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(wm, group_size=128))


def gen_bfp(out: dict):
    """weight_format='bfp' (quant_linear.py:648-723) on the INT input matrix."""
    master = build_input()
    for dn, dt in TORCH_DT.items():
        w = master.to(dt)
        out[f"w_{dn}"] = bits_of(w)
        for w_bit in (3, 4, 5, 6, 8, 12):
            for grp in (32, 64, 128, 512):
                lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False)
                lin.weight.data = w.clone()
                ql = quant_linear.QuantLinear.from_linear(lin, w_bit=w_bit, w_group_size=grp, symmetric=True, weight_format="bfp")
                out[f"{dn}_b{w_bit}_g{grp}"] = bits_of(ql.weight.data)


def build_input_tall(seed: int = 3, rows: int = 512, cols: int = 64) -> torch.Tensor:
    """[512, 64] master for quant_dim = 1 (groups run down the 512 rows): random columns plus adversarial ones."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(rows, cols, generator=g) * 0.02
    w[:, 40] = 0.0113
    w[:, 41] = 0.0
    w[:, 42] = torch.randn(rows, generator=g).abs() * 0.05 + 0.01
    w[:, 43] = torch.randn(rows, generator=g) * 1e-6
    w[:, 44] = torch.randn(rows, generator=g) * 30.0
    w[::3, 45] = 0.0
    w[1::6, 45] = -0.0
    w[100, 46] = 2.0
    w[:, 47] = torch.randn(rows, generator=g) * 1e-3
    return w


APPROX_CASES = [  # (name, weight_format, E, M, hi_align_start, hi_align_exp_field, tail_pad_bits, double)  -- from scripts/opt-6.7b/scripts_4.sh
    ("e1m2_s", "fp4", 1, 2, 1, 1, 0, False),
    ("e2m1_s", "fp4", 2, 1, 1, 1, 0, False),
    ("e2m1_d", "fp4", 2, 1, 0, 2, 1, True),
    ("e2m3_s", "fp6", 2, 3, 4, 7, 2, False),
    ("e2m3_d", "fp6", 2, 3, 0, 2, -1, True),
    ("e3m2_s", "fp6", 3, 2, 4, 7, 2, False),
    ("e3m2_d", "fp6", 3, 2, 4, 6, 0, True),
    ("e4m3_s", "fp8", 4, 3, 12, 15, 1, False),
    ("e4m3_d", "fp8", 4, 3, 13, 15, -1, True),
    ("e4m3_d0", "fp8", 4, 3, 12, 15, 0, True),
]


def gen_dim1_approx(out: dict):
    """quant_dim = 1 for every scheme, and the approximate / double-approximate decode (quant_dim 0 and 1)."""
    master = build_input_tall()
    for dn, dt in TORCH_DT.items():
        w = master.to(dt)
        out[f"w_{dn}"] = bits_of(w)

        def run(key, **kw):
            lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False)
            lin.weight.data = w.clone()
            ql = quant_linear.QuantLinear.from_linear(lin, **kw)
            out[f"{dn}_{key}_out"] = bits_of(ql.weight.data)
            if ql.scales is not None:
                out[f"{dn}_{key}_scales"] = bits_of(ql.scales.reshape(-1))
            if ql.zeros is not None:
                out[f"{dn}_{key}_zeros"] = bits_of(ql.zeros.reshape(-1))

        for grp in (32, 128):
            for sym in (False, True):
                tag = "sym" if sym else "asym"
                run(f"int4_{tag}_g{grp}_d1", w_bit=4, w_group_size=grp, symmetric=sym, weight_format="int", quant_dim=1)
                quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)
                run(f"fp4_{tag}_g{grp}_d1", w_bit=4, w_group_size=grp, symmetric=sym, weight_format="fp4", quant_dim=1)
                run(f"fp8_{tag}_g{grp}_d1", w_bit=8, w_group_size=grp, symmetric=sym, weight_format="fp8", quant_dim=1)
            run(f"int8_asym_g{grp}_d1", w_bit=8, w_group_size=grp, symmetric=False, weight_format="int", quant_dim=1)
            run(f"bfp4_g{grp}_d1", w_bit=4, w_group_size=grp, symmetric=True, weight_format="bfp", quant_dim=1)
        run("int4_asym_pc_d1", w_bit=4, w_group_size=-2, symmetric=False, weight_format="int", quant_dim=1)
        for name, wf, E, M, hs, hf, tp, dbl in APPROX_CASES:
            quant_linear.configure_fp_formats(**{f"{wf}_exp_bits": E, f"{wf}_mantissa_bits": M})
            kw = {f"{wf}_hi_align_start": hs, f"{wf}_hi_align_exp_field": hf, f"{wf}_tail_pad_bits": tp}
            for qd in (0, 1):
                grp = 32 if qd == 0 else 128
                run(f"approx_{name}_g{grp}_d{qd}", w_bit=4, w_group_size=grp, symmetric=True, weight_format=wf, approximate=True, double_approximate=dbl, quant_dim=qd, **kw)
            quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)


def gen_dim1_wide(out: dict):
    """quant_dim = 1 on a [256, 264] matrix -- two 128-row tiles and a ragged second column tile of col_wide_kernel (round 2) --
    through QuantLinear.from_linear exactly as main.py builds it: INT, BFP, exact and single-approximate FP at group 64 / 128."""
    g = torch.Generator().manual_seed(11)
    master = torch.randn(256, 264, generator=g) * 0.02
    master[:, 200] = 0.0113
    master[:, 201] = 0.0
    master[:, 202] = torch.randn(256, generator=g).abs() * 0.05 + 0.01
    master[:, 203] = torch.randn(256, generator=g) * 1e-6
    master[::3, 260] = 0.0
    master[1::6, 260] = -0.0
    master[100, 261] = 2.0
    for dn, dt in TORCH_DT.items():
        w = master.to(dt)
        out[f"w_{dn}"] = bits_of(w)

        def run(key, **kw):
            lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False)
            lin.weight.data = w.clone()
            ql = quant_linear.QuantLinear.from_linear(lin, **kw)
            out[f"{dn}_{key}_out"] = bits_of(ql.weight.data)
            if ql.scales is not None:
                out[f"{dn}_{key}_scales"] = bits_of(ql.scales.reshape(-1))
            if ql.zeros is not None:
                out[f"{dn}_{key}_zeros"] = bits_of(ql.zeros.reshape(-1))

        quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)
        run("int4_asym_g128", w_bit=4, w_group_size=128, symmetric=False, weight_format="int", quant_dim=1)
        run("int3_sym_g64", w_bit=3, w_group_size=64, symmetric=True, weight_format="int", quant_dim=1)
        run("bfp4_g128", w_bit=4, w_group_size=128, symmetric=True, weight_format="bfp", quant_dim=1)
        run("fp4_sym_g128", w_bit=4, w_group_size=128, symmetric=True, weight_format="fp4", quant_dim=1)
        run("fp8_asym_g64", w_bit=8, w_group_size=64, symmetric=False, weight_format="fp8", quant_dim=1)
        for name, wf, E, M, hs, hf, tp, dbl in APPROX_CASES:
            if dbl or name in ("e1m2_s", "e2m3_s"):
                continue
            quant_linear.configure_fp_formats(**{f"{wf}_exp_bits": E, f"{wf}_mantissa_bits": M})
            kw = {f"{wf}_hi_align_start": hs, f"{wf}_hi_align_exp_field": hf, f"{wf}_tail_pad_bits": tp}
            for grp in (64, 128):
                run(f"approx_{name}_g{grp}", w_bit=4, w_group_size=grp, symmetric=True, weight_format=wf, approximate=True, double_approximate=False, quant_dim=1, **kw)
            quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)


ODD_CASES = [  # (shape, kwargs of pseudo_quantize_tensor): segments that are not whole 16-byte vectors, 1-row / 3-D inputs
    ((3, 5), dict(per_tensor=True)), ((7, 129), dict(per_tensor=True)), ((7, 129), dict()), ((5, 36), dict(q_group_size=12)),
    ((1, 8), dict(per_tensor=True)), ((2, 3, 40), dict(q_group_size=20)), ((6, 1000), dict()), ((1, 3), dict()), ((3, 51216), dict()),
]


def gen_odd(out: dict):
    """quant_funcs.pseudo_quantize_tensor (quant_funcs.py:4-46) on shapes outside the vectorised kernels' direct reach (round 2:
    the advisor's [3,5] / [7,129] per-tensor cases and friends; the last one is a per-channel row longer than any ring stage)."""
    for dn, dt in TORCH_DT.items():
        for ci, (shape, kw) in enumerate(ODD_CASES):
            g = torch.Generator().manual_seed(500 + ci)
            n = int(np.prod(shape))
            w = (torch.randn(n, generator=g) * 0.02).reshape(shape).to(dt)
            out[f"w_{dn}_{ci}"] = bits_of(w)
            for bits, zp in ((4, True), (4, False), (8, True), (3, False)):
                y = quant_funcs.pseudo_quantize_tensor(w.clone(), n_bits=bits, zero_point=zp, **kw)
                out[f"{dn}_{ci}_b{bits}_{'asym' if zp else 'sym'}"] = bits_of(y)


def main():
    torch.set_num_threads(8)
    only = sys.argv[1:]
    for name, fn in (("golden_dim1_wide", gen_dim1_wide), ("golden_odd", gen_odd), ("golden_int", gen_int), ("golden_fp_chain", gen_fp_chain), ("golden_fp_tables", gen_fp_tables), ("golden_fp_rne", gen_fp_rne), ("golden_bfp", gen_bfp), ("golden_dim1_approx", gen_dim1_approx)):
        if only and name not in only:
            continue
        d = {}
        fn(d)
        d["_meta"] = np.array([f"torch {torch.__version__}; reference at {REF}; cpu"], dtype="U")
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(f"{name}: {len(d)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
