"""Seeded randomised parity sweep of the CUDA path against the numpy oracle: random shapes, granularities, bit
widths, formats, dtypes and both segment directions, on inputs that mix a normal bulk with the adversarial values
the reference's arithmetic is sensitive to (constant / all-positive / all-negative segments, +-max, subnormals,
exact .5 ties after scaling, outliers x100, signed zeros).  Bit-exact throughout.

Run on the B200 box:  python -m pytest tests/test_gpu_fuzz.py -m gpu -q
"""
import numpy as np
import pytest
import torch

from oracle import fakequant_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from iron_weight_only_quant_b200 import _lib

DEV = "cuda:0"
TDT = {"f16": torch.float16, "bf16": torch.bfloat16}


def bits(t):
    return t.detach().contiguous().cpu().view(torch.int16).numpy().view(np.uint16).copy()


def adversarial(rng: np.random.Generator, rows: int, cols: int, dt: str) -> torch.Tensor:
    w = rng.standard_normal((rows, cols)).astype(np.float32) * 0.02
    flat = w.reshape(-1, 8)
    n = flat.shape[0]
    for kind in range(8):
        idx = rng.integers(0, n, size=max(1, n // 40))
        if kind == 0:
            flat[idx] = 0.0131                      # constant runs
        elif kind == 1:
            flat[idx] = np.abs(flat[idx]) + 0.001   # all positive
        elif kind == 2:
            flat[idx] = -np.abs(flat[idx]) - 0.001  # all negative
        elif kind == 3:
            # large magnitudes of one sign per run (max - min stays finite: overflow to inf is a documented
            # divergence of NaN patterns)
            flat[idx, 0] = 30000.0  # (beyond the fp16 range, bf16 BFP differs from the reference: DESIGN.md section 2)
        elif kind == 4:
            flat[idx] *= 1e-5                       # fp16 subnormals / tiny bf16
        elif kind == 5:
            flat[idx, 3] *= 100.0                   # outliers
        elif kind == 6:
            flat[idx, 1] = -0.0
            flat[idx, 2] = 0.0
        else:
            small = np.abs(flat[idx]) < 1.0
            flat[idx] = np.where(small, np.round(flat[idx] * 64.0) / 64.0 + 1.0 / 128.0, flat[idx])  # exact ties after power-of-two scaling
    return torch.from_numpy(w).to(TDT[dt])


def same(a, b, dt, msg):
    top = 0x7C00 if dt == "f16" else 0x7F80
    na, nb = (a & 0x7FFF) > top, (b & 0x7FFF) > top
    assert np.array_equal(na, nb), f"NaN pattern differs: {msg}: {int(na.sum())} vs {int(nb.sum())} NaNs; first at {np.argwhere(na != nb)[:4].tolist()}"
    bad = np.where(na, 0, a) != np.where(nb, 0, b)
    assert not bad.any(), f"{msg}: {int(bad.sum())} of {a.size} differ; first at {np.argwhere(bad)[:4].tolist()}"


FP_FORMATS = [(2, 1), (1, 2), (3, 2), (2, 3), (4, 3)]


def make_case(seed: int, case: int):
    """Deterministic in (seed, case) alone, so that a failure can be replayed on the CPU emulator."""
    rng = np.random.default_rng([1000 + seed, case])
    dt = ["f16", "bf16"][int(rng.integers(0, 2))]
    scheme = ["int", "int", "fp", "bfp"][int(rng.integers(0, 4))]
    qdim = int(rng.integers(0, 2))
    grp = int(rng.choice([32, 64, 128, 256, -2] if scheme != "bfp" else [32, 64, 128, 256]))
    g_abs = grp if grp > 0 else 8
    # the grouped dimension is a multiple of the group, the other a multiple of 8 (16-byte vectors)
    a = g_abs * int(rng.integers(1, 5))
    b = 8 * int(rng.integers(1, 70))
    rows, cols = (a, b) if qdim == 1 else (b, a)
    if qdim == 0 and grp == -2:
        cols = int(rng.choice([768, 1032, 2048, 4096, 5120, 11008]))
        rows = int(rng.integers(1, 20))
    w = adversarial(rng, rows, cols, dt)
    sym = bool(rng.integers(0, 2))
    if scheme == "int":
        nb = int(rng.choice([2, 3, 4, 5, 6, 8]))
        kw = dict(scheme=_lib.SCHEME_INT, bits=nb, symmetric=sym)
        fn, args = O.int_fake_quant, (dt, nb, not sym, grp)
        tag = f"seed{seed}/{case} {dt} int{nb} sym={sym} g{grp} d{qdim} [{rows},{cols}]"
    elif scheme == "fp":
        E, M = FP_FORMATS[int(rng.integers(0, len(FP_FORMATS)))]
        kw = dict(scheme=_lib.SCHEME_FP_CODEC, exp_bits=E, mant_bits=M, symmetric=sym)
        fn, args = O.fp_codec_fake_quant, (dt, E, M, sym, grp)
        tag = f"seed{seed}/{case} {dt} e{E}m{M} sym={sym} g{grp} d{qdim} [{rows},{cols}]"
    else:
        nb = int(rng.choice([4, 5, 6, 8]))
        kw = dict(scheme=_lib.SCHEME_BFP, bits=nb, symmetric=True)
        fn, args = O.bfp_fake_quant, (dt, nb, grp)
        tag = f"seed{seed}/{case} {dt} bfp{nb} g{grp} d{qdim} [{rows},{cols}]"
    return dict(dt=dt, scheme=scheme, qdim=qdim, grp=grp, rows=rows, cols=cols, w=w, sym=sym, kw=kw, fn=fn, args=args, tag=tag)


@pytest.mark.parametrize("seed", range(6))
def test_random_configurations_match_oracle(seed):
    for case in range(14):
        c = make_case(seed, case)
        dt, rows, cols, grp, qdim, scheme, tag = c["dt"], c["rows"], c["cols"], c["grp"], c["qdim"], c["scheme"], c["tag"]
        wb = bits(c["w"])
        G = _lib.num_segments(rows, cols, grp, qdim)
        scales = torch.zeros(G, dtype=TDT[dt], device=DEV)
        zeros = torch.zeros(G, dtype=TDT[dt], device=DEV)
        wd = c["w"].to(DEV)
        out = torch.empty_like(wd)
        kw = dict(c["kw"])
        cfg = _lib.make_config(TDT[dt], kw.pop("scheme"), grp, quant_dim=qdim, **kw)
        ref = O.along_output_dim(c["fn"], wb, *c["args"]) if qdim == 1 else c["fn"](wb, *c["args"])
        emit_z = scheme != "bfp" and not c["sym"]
        _lib.fake_quant_(wd, out, cfg, rows, cols, scales=None if scheme == "bfp" else scales, zeros=zeros if emit_z else None, track_status=False)
        same(bits(out), ref["out"], dt, tag)
        if scheme != "bfp":
            same(bits(scales), ref["scales"], dt, tag + " scales")
        if emit_z and "zeros" in ref:
            same(bits(zeros), ref["zeros"], dt, tag + " zeros")  # bit-exact, the sign of a zero zero-point included


def _guarded(nbytes: int, guard: int = 4096):
    """uint8 arena with `guard` canary bytes on both sides of an `nbytes` payload (16-byte aligned payload)."""
    pad = (16 - nbytes % 16) % 16
    arena = torch.full((guard + nbytes + pad + guard,), 0xA5, dtype=torch.uint8, device=DEV)
    return arena, arena[guard : guard + nbytes]


def _canaries_intact(arena, nbytes: int, guard: int = 4096) -> bool:
    return bool((arena[:guard] == 0xA5).all().item()) and bool((arena[guard + nbytes + (16 - nbytes % 16) % 16 :] == 0xA5).all().item())


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_no_out_of_bounds_writes(dt):
    """compute-sanitizer is not available on the GPU pool, so: every output (weights, scales, zero points, codes)
    sits between canary zones; ragged tiles, odd row counts and every kernel family must leave them untouched."""
    esz = 2
    cases = [
        # rows, cols, group, quant_dim, scheme kwargs
        (3, 200, 8, 0, dict(scheme="int", bits=4)), (1025, 136, 136, 0, dict(scheme="int", bits=4)), (7, 4096, -2, 0, dict(scheme="int", bits=8)),
        (5, 11008, -2, 0, dict(scheme="int", bits=4)), (9, 2048, 2048, 0, dict(scheme="int", bits=4)), (129, 384, 128, 0, dict(scheme="fp", E=2, M=1)),
        (65, 1024, -1, 0, dict(scheme="int", bits=4)), (384, 264, -2, 1, dict(scheme="int", bits=4)), (256, 520, 128, 1, dict(scheme="int", bits=4)),
        (512, 72, 32, 1, dict(scheme="fp", E=4, M=3)), (512, 264, 128, 1, dict(scheme="fp", E=2, M=1, approx=2)), (33, 512, 64, 0, dict(scheme="bfp", bits=6)),
    ]
    for rows, cols, grp, qd, kw in cases:
        n = rows * cols
        G = _lib.num_segments(rows, cols, grp, qd)
        g = torch.Generator().manual_seed(rows + cols)
        w = (torch.randn(rows, cols, generator=g) * 0.02).to(TDT[dt])
        a_w, p_w = _guarded(n * esz)
        a_s, p_s = _guarded(G * esz)
        a_z, p_z = _guarded(G * esz)
        a_c, p_c = _guarded(n)
        wv = p_w.view(TDT[dt]).view(rows, cols)
        wv.copy_(w.to(DEV))
        sym = kw.get("sym", False)
        if kw["scheme"] == "int":
            cfg = _lib.make_config(TDT[dt], _lib.SCHEME_INT, grp, bits=kw["bits"], symmetric=sym, quant_dim=qd)
        elif kw["scheme"] == "fp":
            ap = kw.get("approx", 0)
            cfg = _lib.make_config(TDT[dt], _lib.SCHEME_FP_CODEC, grp, exp_bits=kw["E"], mant_bits=kw["M"], symmetric=bool(ap) or sym, quant_dim=qd, approx=ap, hi_align_start=0, hi_align_exp_field=2, tail_pad_bits=1)
        else:
            cfg = _lib.make_config(TDT[dt], _lib.SCHEME_BFP, grp, bits=kw["bits"], symmetric=True, quant_dim=qd)
        emits_q = kw["scheme"] != "bfp"
        emits_z = emits_q and not cfg.symmetric
        want_codes = qd == 0 and kw["scheme"] == "int"
        _lib.fake_quant_(wv, wv, cfg, rows, cols, scales=p_s.view(TDT[dt]) if emits_q else None, zeros=p_z.view(TDT[dt]) if emits_z else None, codes=p_c if want_codes else None, track_status=False)
        torch.cuda.synchronize()
        tag = f"{dt} [{rows},{cols}] g{grp} d{qd} {kw}"
        assert _canaries_intact(a_w, n * esz), tag + ": weight canary"
        assert _canaries_intact(a_s, G * esz), tag + ": scales canary"
        assert _canaries_intact(a_z, G * esz), tag + ": zeros canary"
        assert _canaries_intact(a_c, n), tag + ": codes canary"
        if not emits_z:
            assert bool((p_z == 0xA5).all().item()), tag + ": zeros written although not requested"


ALT_WORKER = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.environ["IWQ_ROOT"])
from iron_weight_only_quant_b200 import _lib
from oracle import fakequant_oracle as O
DEV = "cuda:0"
def bits(t): return t.detach().contiguous().cpu().view(torch.int16).numpy().view(np.uint16).copy()
bad = []
cases = [  # rows, cols, group, quant_dim, bits, sym
    (256, 520, 128, 1, 4, False), (512, 264, 64, 1, 4, True), (1024, 136, 256, 1, 3, False), (64, 72, 32, 1, 4, False),
    (9, 4096, -2, 0, 4, False), (5, 8192, -2, 0, 8, True), (7, 2048, -2, 0, 4, False), (33, 11008, -2, 0, 4, False), (6, 5120, -2, 0, 4, True),
    (3, 8192, 4096, 0, 4, False), (64, 1024, 128, 0, 4, False),
]
for rows, cols, grp, qd, nb, sym in cases:
    g = torch.Generator().manual_seed(rows * 3 + cols)
    w = (torch.randn(rows, cols, generator=g) * 0.02).half()
    ref = O.along_output_dim(O.int_fake_quant, bits(w), "f16", nb, not sym, grp) if qd else O.int_fake_quant(bits(w), "f16", nb, not sym, grp)
    wd = w.to(DEV)
    G = _lib.num_segments(rows, cols, grp, qd)
    sc = torch.zeros(G, dtype=torch.float16, device=DEV)
    cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, grp, bits=nb, symmetric=sym, quant_dim=qd)
    _lib.fake_quant_(wd, wd, cfg, rows, cols, scales=sc, track_status=False)
    if not (np.array_equal(bits(wd), ref["out"]) and np.array_equal(bits(sc), ref["scales"])):
        bad.append((rows, cols, grp, qd, nb, sym))
print("ALT_OK" if not bad else f"ALT_MISMATCH {bad}")
'''


@pytest.mark.parametrize("knobs", ["IWQ_COL_MODE=1", "IWQ_COL_TILE=-1", "IWQ_COL_STAGES=2", "IWQ_WIDE_SEG4=0", "IWQ_WIDE_SEG4=2", "IWQ_ROW_TMA_OUT=0", "IWQ_ROW_TMA_OUT=2", "IWQ_VARIANT=2", "IWQ_SEG_STAGES=4"])
def test_alternate_kernel_variants_match_oracle(knobs, tmp_path):
    """The tuning knobs select other kernels / tile shapes (register-resident and two-sweep column kernels, row kernel
    with and without bulk stores, multi-warp segments on or off ...); they are read once per process, so each setting
    runs in a subprocess.  Every variant must be bit-exact like the default."""
    import os
    import subprocess
    import sys

    script = tmp_path / "alt_worker.py"
    script.write_text(ALT_WORKER)
    env = dict(os.environ, IWQ_ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for kv in knobs.split():
        k, v = kv.split("=")
        env[k] = v
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert "ALT_OK" in out.stdout, knobs + ": " + out.stdout[-1500:] + out.stderr[-1500:]


@pytest.mark.parametrize("seed", range(4))
def test_random_column_configurations_wide_tiles(seed):
    """quant_dim = 1 on shapes that reach col_wide_kernel (rows >= 128, cols >= 256): several 128-row tiles, ragged last
    column tile, row counts that are not a multiple of the tile (rows = group x odd), every scheme / dtype, adversarial
    values; scales and zero points in the reference's column-major segment order."""
    rng = np.random.default_rng([2000 + seed])
    for case in range(8):
        dt = ["f16", "bf16"][int(rng.integers(0, 2))]
        scheme = ["int", "fp", "bfp", "int"][int(rng.integers(0, 4))]
        grp = int(rng.choice([32, 64, 128]))
        rows = grp * int(rng.integers(max(1, 128 // grp), 40 if grp == 32 else 13))
        rows = max(rows, 128)
        cols = 8 * int(rng.integers(32, 170))
        w = adversarial(rng, rows, cols, dt)
        sym = bool(rng.integers(0, 2))
        if scheme == "int":
            nb = int(rng.choice([2, 3, 4, 5, 8]))
            kw = dict(bits=nb, symmetric=sym)
            sid, fn, args = _lib.SCHEME_INT, O.int_fake_quant, (dt, nb, not sym, grp)
        elif scheme == "fp":
            E, M = FP_FORMATS[int(rng.integers(0, len(FP_FORMATS)))]
            kw = dict(exp_bits=E, mant_bits=M, symmetric=sym)
            sid, fn, args = _lib.SCHEME_FP_CODEC, O.fp_codec_fake_quant, (dt, E, M, sym, grp)
        else:
            nb = int(rng.choice([4, 6, 8]))
            kw = dict(bits=nb, symmetric=True)
            sid, fn, args = _lib.SCHEME_BFP, O.bfp_fake_quant, (dt, nb, grp)
        tag = f"seed{seed}/{case} {dt} {scheme} {kw} g{grp} d1 [{rows},{cols}]"
        ref = O.along_output_dim(fn, bits(w), *args)
        G = _lib.num_segments(rows, cols, grp, 1)
        scales = torch.zeros(G, dtype=TDT[dt], device=DEV)
        zeros = torch.zeros(G, dtype=TDT[dt], device=DEV)
        emit_z = scheme != "bfp" and not sym
        # in place and batched with a second matrix of another shape (tensor maps per matrix behind the job table)
        wd = w.to(DEV)
        w2 = adversarial(rng, 256, 264, dt)
        ref2 = O.along_output_dim(fn, bits(w2), *args)
        w2d = w2.to(DEV)
        cfg = _lib.make_config(TDT[dt], sid, grp, quant_dim=1, **kw)
        mats = [_lib.matrix_desc(wd, wd, rows, cols, None if scheme == "bfp" else scales, zeros if emit_z else None), _lib.matrix_desc(w2d, w2d, 256, 264)]
        _lib.fake_quant_batched_(mats, cfg, torch.device(DEV), track_status=False)
        same(bits(wd), ref["out"], dt, tag)
        same(bits(w2d), ref2["out"], dt, tag + " (second matrix of the batch)")
        if scheme != "bfp":
            same(bits(scales), ref["scales"], dt, tag + " scales")
        if emit_z and "zeros" in ref:
            same(bits(zeros), ref["zeros"], dt, tag + " zeros")
