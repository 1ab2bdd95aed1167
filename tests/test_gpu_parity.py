"""Parity of the CUDA path (through the C ABI) against (a) the reference's own outputs committed as
golden fixtures and (b) the numpy oracle on seeded inputs.  Bit-exact throughout: the path is
integer/bit-level arithmetic on 16-bit storage; the only tolerance is "any NaN equals any NaN".

Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import fakequant_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from iron_weight_only_quant_b200 import _lib, fp4_quantize_cpu, quant_funcs, quant_linear, quant_wrapper, shapes, sharded

DEV = "cuda:0"
TDT = {"f16": torch.float16, "bf16": torch.bfloat16}
FORMATS = {"e2m1": ("fp4", 2, 1), "e1m2": ("fp4", 1, 2), "e3m2": ("fp6", 3, 2), "e2m3": ("fp6", 2, 3), "e4m3": ("fp8", 4, 3), "e5m2": ("fp8", 5, 2)}
INT_BITS = [2, 3, 4, 5, 6, 8]
GROUPS = [32, 64, 128, 256, -1, -2]


def bits(t: torch.Tensor) -> np.ndarray:
    return t.detach().contiguous().cpu().view(torch.int16).numpy().view(np.uint16).copy()


def tensor_of(b: np.ndarray, dt: str, device=DEV) -> torch.Tensor:
    return torch.from_numpy(b.view(np.int16).copy()).view(TDT[dt]).to(device)


def assert_bits(a: np.ndarray, b: np.ndarray, dt: str, msg=""):
    top = 0x7C00 if dt == "f16" else 0x7F80
    na, nb = (a & 0x7FFF) > top, (b & 0x7FFF) > top
    assert np.array_equal(na, nb), f"NaN pattern differs: {msg}"
    bad = np.where(na, 0, a) != np.where(nb, 0, b)
    assert not bad.any(), f"{msg}: {int(bad.sum())} of {a.size} differ; first at {np.argwhere(bad)[:4].tolist()}"


def fn_kwargs(bits_, sym, grp):
    kw = dict(n_bits=bits_, zero_point=not sym)
    if grp > 0:
        kw["q_group_size"] = grp
    elif grp == -1:
        kw["per_tensor"] = True
    return kw


# ------------------------------------------------------------------------------------------------
# (a) against the reference's own outputs
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("nb", INT_BITS)
def test_int_matches_reference_golden(golden, dt, sym, nb):
    g = golden("golden_int")
    w = tensor_of(g[f"w_{dt}"], dt)
    for grp in GROUPS:
        key = f"{dt}_b{nb}_{'sym' if sym else 'asym'}_g{grp}"
        # module form: in place on the Linear's storage, scales / zeros buffers (quant_linear.py:885-956)
        lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, device=DEV, dtype=TDT[dt])
        lin.weight.data.copy_(w)
        ptr = lin.weight.data.data_ptr()
        ql = quant_linear.QuantLinear.from_linear(lin, w_bit=nb, w_group_size=grp, symmetric=sym, weight_format="int")
        assert ql.weight.data.data_ptr() == ptr and not ql.weight.requires_grad
        assert_bits(bits(ql.weight.data), g[f"ql_{key}_out"], dt, key)
        assert_bits(bits(ql.scales.reshape(-1)), g[f"ql_{key}_scales"], dt, key + " scales")
        assert ql.scales.dtype == TDT[dt] and ql.scales.shape[1] == 1
        if sym:
            assert ql.zeros is None
        else:
            assert_bits(bits(ql.zeros.reshape(-1)), g[f"ql_{key}_zeros"], dt, key + " zeros")  # sign of a zero zero-point included
        # function form (quant_funcs.py:4-46)
        if f"fn_{key}_assert" in g.files:
            with pytest.raises(AssertionError):
                quant_funcs.pseudo_quantize_tensor(w.clone(), **fn_kwargs(nb, sym, grp))
        else:
            src = w.clone()
            o = quant_funcs.pseudo_quantize_tensor(src, **fn_kwargs(nb, sym, grp))
            assert o.data_ptr() != src.data_ptr() and torch.equal(src, w)  # out of place leaves the input intact
            assert np.array_equal(bits(o), g[f"fn_{key}"]), key


def test_inplace_contract(golden):
    g = golden("golden_int")
    w = tensor_of(g["w_f16"], "f16")
    t = w.clone()
    r = quant_funcs.pseudo_quantize_tensor(t, n_bits=4, zero_point=True, q_group_size=128, inplace=True)
    assert r.data_ptr() == t.data_ptr()
    assert np.array_equal(bits(t), g["fn_inplace_mutated_input"])
    assert np.array_equal(bits(r), g["fn_inplace_f16_b4_asym_g128"])


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_fp_chain_matches_reference_golden(golden, dt, fmt):
    wf, E, M = FORMATS[fmt]
    g = golden("golden_fp_chain")
    w = tensor_of(g[f"w_{dt}"], dt)
    quant_linear.configure_fp_formats(**{f"{wf}_exp_bits": E, f"{wf}_mantissa_bits": M})
    try:
        if dt == "f16" and E >= 5:
            lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, device=DEV, dtype=TDT[dt])
            with pytest.raises(RuntimeError):
                quant_linear.QuantLinear.from_linear(lin, w_bit=4, w_group_size=128, symmetric=True, weight_format=wf)
            return
        for sym in (False, True):
            for grp in (32, 128, -1, -2):
                key = f"{dt}_{fmt}_{'sym' if sym else 'asym'}_g{grp}"
                lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, device=DEV, dtype=TDT[dt])
                lin.weight.data.copy_(w)
                ql = quant_linear.QuantLinear.from_linear(lin, w_bit=4, w_group_size=grp, symmetric=sym, weight_format=wf)
                assert_bits(bits(ql.weight.data), g[f"ql_{key}_out"], dt, key)
                assert ql.scales.dtype == torch.float16
                assert np.array_equal(bits(ql.scales.reshape(-1)), g[f"ql_{key}_scales_half"]), key + " scales"
                if sym:
                    assert ql.zeros is None
                else:
                    assert np.array_equal(O.from_bits(bits(ql.zeros.reshape(-1)), "f16"), O.from_bits(g[f"ql_{key}_zeros_half"], "f16")), key
    finally:
        quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("fmt", list(FORMATS))
def test_fp_codec_exhaustive_16bit(golden, dt, fmt):
    """decode(encode(x)) for every 16-bit input pattern vs the reference's _float_to_fp/_fp_to_float."""
    _, E, M = FORMATS[fmt]
    if dt == "f16" and E >= 5:
        return
    g = golden("golden_fp_tables")
    allb = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    x = tensor_of(allb, dt)
    y, codes = quant_linear._codec_roundtrip(x, E, M)
    xv = O.from_bits(allb, dt)
    fmax = O.fp_format(E, M)[3]
    ok = np.isfinite(xv) & (np.abs(xv) <= fmax)
    if dt == "bf16":
        ok &= (np.abs(xv) >= 2.0 ** -126) | (xv == 0)  # quirk Q5: the reference's int8 cast wraps below 2^-128
    assert np.array_equal(bits(y)[ok], g[f"codec_{dt}_{fmt}_val"][ok])
    assert np.array_equal(codes.cpu().numpy()[ok], g[f"codec_{dt}_{fmt}_code"][ok])


@pytest.mark.parametrize("fmt", ["e2m1", "e1m2", "e3m2", "e2m3", "e4m3"])
def test_fp_rne_exhaustive_16bit(golden, fmt):
    _, E, M = FORMATS[fmt]
    g = golden("golden_fp_tables")
    allb = np.arange(65536, dtype=np.uint32).astype(np.uint16)
    x = tensor_of(allb, "f16")
    bias = 2 ** (E - 1) - 1
    mf = (2 - 2 ** (-M)) * 2 ** (2 ** E - 1 - bias)
    q = fp4_quantize_cpu._fp_scale(x.reshape(1, -1), torch.ones(1, 1, dtype=torch.float16, device=DEV), M, bias, mf, -mf)
    ok = (allb & 0x7FFF) < 0x7C00
    assert np.array_equal(bits(q).reshape(-1)[ok], g[f"fpscale_f16_{fmt}"][ok])


def test_fp_rne_matches_reference_golden(golden):
    g = golden("golden_fp_rne")
    w = tensor_of(g["w_f16"], "f16")
    for grp in (32, 128, 512):
        o = fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(w.clone(), group_size=grp)
        assert tuple(o.shape) == tuple(g[f"e2m1_g{grp}_shape"])
        assert_bits(bits(o), g[f"e2m1_g{grp}"], "f16", f"g{grp}")
    o = fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(w.clone(), group_size=-1, per_tensor=True)
    assert tuple(o.shape) == tuple(g["e2m1_pt_shape"])
    assert_bits(bits(o), g["e2m1_pt"], "f16", "per_tensor")
    o = fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(tensor_of(g["main_w"], "f16"), group_size=128)
    assert np.array_equal(bits(o), g["main_out"])  # the reference's own __main__ input (fp4_quantize_cpu.py:74-80)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_division_exhaustive_on_device(dt):
    """rD(reciprocal*w with FMA correction) == rD(IEEE w/s) for every (w, s) pair of the dtype."""
    cnt = torch.zeros(1, dtype=torch.int64, device=DEV)
    rc = _lib.lib().iwq_selftest_division(_lib.dtype_code(TDT[dt]), cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc)
    assert int(cnt.item()) == 0


# ------------------------------------------------------------------------------------------------
# (b) against the oracle on seeded inputs, all kernel paths
# ------------------------------------------------------------------------------------------------
def seeded(rows, cols, dt, seed, scale=0.02):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) * scale).to(TDT[dt])


def run_int(w, nb, sym, grp, variant=0, want_codes=True):
    rows, cols = w.shape
    G = _lib.num_segments(rows, cols, grp)
    out = torch.empty_like(w)
    scales = torch.empty(G, dtype=w.dtype, device=w.device)
    zeros = torch.empty(G, dtype=w.dtype, device=w.device)
    codes = torch.empty(rows, cols, dtype=torch.uint8, device=w.device) if want_codes else None
    cfg = _lib.make_config(w.dtype, _lib.SCHEME_INT, grp, bits=nb, symmetric=sym, variant=variant)
    _lib.fake_quant_(w, out, cfg, rows, cols, scales=scales, zeros=zeros, codes=codes)
    return out, scales, zeros, codes


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_int_vs_oracle_all_outputs(dt, variant):
    """output, scales, zero points AND integer codes, all kernel variants (0: TMA in + TMA out, 4 vectors per
    lane; 1: direct global loads/stores; 2: TMA in + register stores)."""
    w = seeded(512, 2048, dt, 11)
    wb = bits(w)
    wd = w.to(DEV)
    for nb, sym, grp in [(4, False, 128), (4, True, 128), (3, False, 32), (8, False, 64), (8, True, 256), (2, True, 8), (4, False, 512), (5, True, 1024), (6, False, 16), (4, False, -1), (8, True, -1)]:
        out, scales, zeros, codes = run_int(wd, nb, sym, grp, variant)
        ref = O.int_fake_quant(wb, dt, nb, not sym, grp)
        tag = f"{dt} b{nb} sym={sym} g{grp} v{variant}"
        assert_bits(bits(out), ref["out"], dt, tag)
        assert_bits(bits(scales), ref["scales"], dt, tag + " scales")
        assert_bits(bits(zeros), ref["zeros"], dt, tag + " zeros")
        c = codes.cpu().numpy().astype(np.int32)
        if sym:
            c = np.where(c >= 128, c - 256, c)  # two's complement int8
        assert np.array_equal(c, ref["codes"]), tag + " codes"


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_packed_w4_codes_vs_oracle(dt):
    """code_bits = 4: two codes per byte, element 2i in the low nibble; same codes as the
    unpacked emission / the oracle, for INT (asym, sym two's complement) and FP4, all kernels that emit codes."""
    def pack(c):
        c = (c.astype(np.int64) & 0xF).astype(np.uint8).reshape(c.shape[0], -1, 2)
        return c[:, :, 0] | (c[:, :, 1] << 4)

    for rows, cols, grp, variant in [(256, 2048, 128, 0), (64, 1024, 32, 0), (48, 768, -2, 0), (30, 11008, -2, 0), (16, 512, 16, 2), (128, 1024, 128, 1), (8, 4096, -1, 0), (3, 200, 8, 0)]:
        w = seeded(rows, cols, dt, rows * 7 + cols)
        wd = w.to(DEV)
        for nb, sym in [(4, False), (4, True), (3, False), (2, True)]:
            ref = O.int_fake_quant(bits(w), dt, nb, not sym, grp)
            codes = torch.zeros(rows, cols // 2, dtype=torch.uint8, device=DEV)
            cfg = _lib.make_config(TDT[dt], _lib.SCHEME_INT, grp, bits=nb, symmetric=sym, variant=variant, code_bits=4)
            out = torch.empty_like(wd)
            _lib.fake_quant_(wd, out, cfg, rows, cols, codes=codes)
            tag = f"{dt} [{rows},{cols}] g{grp} b{nb} sym={sym} v{variant}"
            assert_bits(bits(out), ref["out"], dt, tag)
            assert np.array_equal(codes.cpu().numpy(), pack(ref["codes"])), tag + " packed codes"
    # FP4 E2M1 / E1M2: packed minifloat codes == the unpacked emission
    w = seeded(64, 1024, dt, 5).to(DEV)
    for E, M in [(2, 1), (1, 2)]:
        c8 = torch.zeros(64, 1024, dtype=torch.uint8, device=DEV)
        c4 = torch.zeros(64, 512, dtype=torch.uint8, device=DEV)
        for cb, buf in [(8, c8), (4, c4)]:
            cfg = _lib.make_config(TDT[dt], _lib.SCHEME_FP_CODEC, 128, exp_bits=E, mant_bits=M, symmetric=True, code_bits=cb)
            _lib.fake_quant_(w, torch.empty_like(w), cfg, 64, 1024, codes=buf)
        assert np.array_equal(c4.cpu().numpy(), pack(c8.cpu().numpy())), f"{dt} fp4 e{E}m{M} packed codes"
    with pytest.raises(ValueError):  # 8-bit codes do not fit a nibble
        _lib.fake_quant_(w, torch.empty_like(w), _lib.make_config(TDT[dt], _lib.SCHEME_INT, 128, bits=8, code_bits=4), 64, 1024, codes=c4)
    # the Python-level export: packed codes + scales + zero points reproduce the fake-quant weight exactly
    pk = quant_funcs.quantize_to_packed_w4(w, n_bits=4, zero_point=True, q_group_size=128)
    fq = quant_funcs.pseudo_quantize_tensor(w, n_bits=4, zero_point=True, q_group_size=128)
    q = torch.stack([pk["codes"] & 0xF, pk["codes"] >> 4], dim=-1).reshape(64, 1024).to(torch.float32)
    deq = ((q.reshape(-1, 128) - pk["zeros"].float()) * pk["scales"].float()).reshape(64, 1024)
    assert torch.equal(deq.to(TDT[dt]), fq)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_per_tensor_fused_single_launch_vs_oracle(dt):
    """iwq_fake_quant_per_tensor_fused on one GPU (world 1): several matrices of mixed sizes (ragged tiles, more
    matrices than batches and batches than matrices), asym and sym, bit-exact vs the unfused oracle."""
    shapes_ = [(64, 1024), (3, 200), (700, 512), (1, 8), (2048, 4096), (129, 264), (4096, 4096), (5, 40)]
    ws = [seeded(r, c, dt, 300 + i) for i, (r, c) in enumerate(shapes_)]
    for nb, sym in [(4, False), (8, True), (3, False)]:
        ds = [w.to(DEV) for w in ws]
        outs = sharded.pseudo_quantize_tensors_per_tensor_fused(ds, n_bits=nb, zero_point=not sym)
        assert _lib.read_and_clear_status(torch.device(DEV)) == 0
        for w, o in zip(ws, outs):
            ref = O.int_fake_quant(bits(w), dt, nb, not sym, -1)
            assert_bits(bits(o), ref["out"], dt, f"{dt} fused per-tensor b{nb} sym={sym} {tuple(w.shape)}")
    # scales / zero points of the module form
    ds = [w.to(DEV) for w in ws]
    sc = [torch.zeros(1, dtype=TDT[dt], device=DEV) for _ in ws]
    zs = [torch.zeros(1, dtype=TDT[dt], device=DEV) for _ in ws]
    cfg = _lib.make_config(TDT[dt], _lib.SCHEME_INT, -1, bits=4, symmetric=False)
    mats = [_lib.matrix_desc(d, d, d.shape[0], d.shape[1], s_, z_) for d, s_, z_ in zip(ds, sc, zs)]
    _lib.per_tensor_fused_(mats, cfg, torch.device(DEV))
    for w, s_, z_ in zip(ws, sc, zs):
        ref = O.int_fake_quant(bits(w), dt, 4, True, -1)
        assert_bits(bits(s_), ref["scales"], dt, "fused scales")
        assert_bits(bits(z_), ref["zeros"], dt, "fused zeros")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_per_channel_row_kernel_vs_oracle(dt):
    """per-channel goes through row_tma_kernel for every row length that is not a power of two <= 1024."""
    # 2048 / 4096-element rows take seg4_tma_kernel with 2 / 4 warps per segment (ragged last tiles included)
    for rows, cols in [(64, 768), (96, 3072), (40, 4096), (33, 11008), (7, 28672), (130, 1024), (50, 1032), (7, 2048), (9, 4096), (5, 8192), (1, 4096)]:
        w = seeded(rows, cols, dt, rows + cols)
        wb = bits(w)
        for nb, sym in [(8, False), (4, True), (4, False)]:
            out, scales, zeros, codes = run_int(w.to(DEV), nb, sym, -2)
            ref = O.int_fake_quant(wb, dt, nb, not sym, -2)
            tag = f"{dt} [{rows},{cols}] b{nb} sym={sym}"
            assert_bits(bits(out), ref["out"], dt, tag)
            assert_bits(bits(scales), ref["scales"], dt, tag + " scales")
            c = codes.cpu().numpy().astype(np.int32)
            if sym:
                c = np.where(c >= 128, c - 256, c)
            assert np.array_equal(c, ref["codes"]), tag + " codes"


def test_odd_group_sizes_and_ragged_tiles():
    """group sizes outside the shuffle path (96, 192, 2048) and matrices that do not fill a 16 KB tile."""
    for rows, cols, grp in [(8, 192, 96), (16, 384, 192), (4, 4096, 2048), (3, 8192, 4096), (3, 16384, 8192), (1, 8, 8), (3, 40, 8), (5, 8200, 8), (1, 128, 128), (1025, 136, 136)]:
        w = seeded(rows, cols, "f16", cols)
        out, scales, zeros, _ = run_int(w.to(DEV), 4, False, grp, want_codes=False)
        ref = O.int_fake_quant(bits(w), "f16", 4, True, grp)
        assert_bits(bits(out), ref["out"], "f16", f"[{rows},{cols}] g{grp}")
        assert_bits(bits(scales), ref["scales"], "f16", f"[{rows},{cols}] g{grp} scales")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_generic_bit_widths_vs_oracle(dt):
    """bit widths that leave the packed fast path (INT_GENERIC kernel; bf16 7/8-bit INT_WIDE)."""
    w = seeded(64, 1024, dt, 5, scale=1.0)
    wb = bits(w)
    for nb, sym in [(7, False), (8, False), (9, True), (10, False), (12, True), (1, False)]:
        ref = O.int_fake_quant(wb, dt, nb, not sym, 128)
        out, scales, _, _ = run_int(w.to(DEV), nb, sym, 128, want_codes=False)
        assert_bits(bits(out), ref["out"], dt, f"{dt} b{nb} sym={sym}")
        assert_bits(bits(scales), ref["scales"], dt, f"{dt} b{nb} scales")


def test_nan_status_and_assert():
    w = seeded(16, 256, "f16", 3).to(DEV)
    w[3, 17] = float("nan")
    with pytest.raises(AssertionError):
        quant_funcs.pseudo_quantize_tensor(w, n_bits=4, q_group_size=128)
    ok = quant_funcs.pseudo_quantize_tensor(seeded(16, 256, "f16", 3).to(DEV), n_bits=4, q_group_size=128)
    assert not torch.isnan(ok).any()


def test_error_contract():
    w = seeded(8, 100, "f16", 1).to(DEV)
    with pytest.raises(AssertionError):
        quant_funcs.pseudo_quantize_tensor(w, n_bits=4, q_group_size=32)  # quant_funcs.py:11
    with pytest.raises(AssertionError):
        quant_funcs.pseudo_quantize_tensor(w.reshape(-1), n_bits=4)  # quant_funcs.py:15
    with pytest.raises(ValueError):
        fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(w, group_size=32)  # fp4_quantize_cpu.py:55-56
    with pytest.raises(ValueError):
        fp4_quantize_cpu.quantize_fp16_to_fp4_e1m2(w.reshape(-1))  # fp4_quantize_cpu.py:50-51
    lin = torch.nn.Linear(128, 8, bias=False, device=DEV, dtype=torch.float16)
    with pytest.raises(ValueError):
        quant_linear.QuantLinear.from_linear(lin, w_group_size=-3)  # quant_linear.py:906
    with pytest.raises(ValueError):
        quant_linear.QuantLinear(8, 8, weight_format="int3")  # quant_linear.py:441
    with pytest.raises(RuntimeError):
        quant_funcs.pseudo_quantize_tensor(torch.zeros(4, 128, dtype=torch.float16), n_bits=4)  # CPU tensor: no fallback
    with pytest.raises(TypeError):
        quant_funcs.pseudo_quantize_tensor(torch.zeros(4, 128, device=DEV), n_bits=4)  # fp32: not implemented
    ql = quant_linear.QuantLinear.from_linear(lin, w_bit=16)  # quant_linear.py:887-892: untouched
    assert not bool(ql.quantized)


def test_wrappers_match_oracle():
    w = seeded(32, 512, "f16", 9)
    wb = bits(w)
    wd = w.to(DEV)
    assert np.array_equal(bits(quant_funcs.quantize_weight_per_channel_absmax(wd, 8)), O.int_fake_quant(wb, "f16", 8, False, -2)["out"])
    assert np.array_equal(bits(quant_funcs.quantize_weight_per_tensor_absmax(wd, 8)), O.int_fake_quant(wb, "f16", 8, False, -1)["out"])
    a = wd.reshape(4, 8, 512)
    r = quant_funcs.quantize_activation_per_token_absmax(a, 8)
    assert r.shape == a.shape
    assert np.array_equal(bits(r).reshape(32, 512), O.int_fake_quant(wb, "f16", 8, True, -2)["out"])
    r = quant_funcs.quantize_activation_per_tensor_absmax(a, 8)
    assert np.array_equal(bits(r).reshape(32, 512), O.int_fake_quant(wb, "f16", 8, True, -1)["out"])


# ------------------------------------------------------------------------------------------------
# model walker / batched launch / sharded per-tensor
# ------------------------------------------------------------------------------------------------
class _Args:
    def __init__(self, **kw):
        self.w_bit, self.a_bit, self.w_group_size, self.w_symmetric, self.w_format = 4, 16, 128, False, "int"
        self.__dict__.update(kw)


class _Block(torch.nn.Module):
    def __init__(self, h, f):
        super().__init__()
        self.q_proj = torch.nn.Linear(h, h, bias=False)
        self.k_proj = torch.nn.Linear(h, h // 2, bias=False)
        self.up_proj = torch.nn.Linear(h, f, bias=True)
        self.down_proj = torch.nn.Linear(f, h, bias=False)


class _Toy(torch.nn.Module):
    def __init__(self, h=256, f=640, n=3):
        super().__init__()
        self.layers = torch.nn.ModuleList([_Block(h, f) for _ in range(n)])
        self.lm_head = torch.nn.Linear(h, 1000, bias=False)


@pytest.mark.parametrize("cfg", [dict(), dict(w_symmetric=True, w_group_size=64), dict(w_group_size=-2, w_bit=8), dict(w_group_size=-1), dict(w_format="fp4", w_symmetric=True), dict(w_format="fp8", w_group_size=-2)])
def test_quantize_model_batched_matches_oracle(cfg):
    torch.manual_seed(0)
    model = _Toy().half().to(DEV)
    before = {n: bits(m.weight.data) for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)}
    ptrs = {n: m.weight.data.data_ptr() for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)}
    args = _Args(**cfg)
    out_model = quant_wrapper.quantize_model(model, args)
    assert out_model is model
    n_q = 0
    for n, m in model.named_modules():
        if isinstance(m, quant_linear.QuantLinear):
            n_q += 1
            assert m.weight.data.data_ptr() == ptrs[n]  # storage aliased, overwritten in place
            if args.w_format == "int":
                ref = O.int_fake_quant(before[n], "f16", args.w_bit, not args.w_symmetric, args.w_group_size)
                assert_bits(bits(m.scales.reshape(-1)), ref["scales"], "f16", n)
            else:
                E, M = {"fp4": (2, 1), "fp8": (4, 3)}[args.w_format]
                ref = O.fp_codec_fake_quant(before[n], "f16", E, M, args.w_symmetric, args.w_group_size)
                assert np.array_equal(bits(m.scales.reshape(-1)), ref["scales_half"]), n
            assert_bits(bits(m.weight.data), ref["out"], "f16", n)
            assert bool(m.quantized)
    assert n_q == 12
    assert isinstance(model.lm_head, torch.nn.Linear) and np.array_equal(bits(model.lm_head.weight.data), before["lm_head"])
    x = torch.randn(2, 5, 256, device=DEV, dtype=torch.float16)
    y = model.layers[0].up_proj(x)
    assert y.shape == (2, 5, 640)
    assert torch.equal(y, torch.nn.functional.linear(x, model.layers[0].up_proj.weight, model.layers[0].up_proj.bias))


def test_wrapper_gates_and_errors():
    model = _Toy().half().to(DEV)
    assert quant_wrapper.quantize_model(model, _Args(w_bit=16)) is model  # gate quant_wrapper.py:15
    assert isinstance(model.layers[0].q_proj, torch.nn.Linear)
    with pytest.raises(ValueError):
        quant_wrapper.quantize_model(model, _Args(w_format="bfp", w_group_size=-1))  # quant_wrapper.py:19-20
    with pytest.raises(NotImplementedError):
        quant_wrapper.quantize_model(model, _Args(w_format="int", approximate=True))  # quant_linear.py:601
    with pytest.raises(NotImplementedError):
        quant_wrapper.quantize_model(model, _Args(gptq=True, w_format="fp4"))  # quant_wrapper.py:24-25


def test_per_tensor_sharded_equals_unsharded():
    """Rows of each matrix split in two "ranks" on one GPU; the MAX all-reduce of {max,-min} is emulated by
    torch.maximum over the two ranks' range buffers (the gloo test covers the real collective)."""
    for sym in (False, True):
        ws = [seeded(256, 512, "f16", s, scale=0.05 * (s + 1)) for s in range(3)]
        refs = [O.int_fake_quant(bits(w), "f16", 4, not sym, -1)["out"] for w in ws]
        cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, -1, bits=4, symmetric=sym)
        halves = [[w[:100].contiguous().to(DEV) for w in ws], [w[100:].contiguous().to(DEV) for w in ws]]
        rngs = []
        for h in halves:
            mats = [_lib.matrix_desc(s, s, s.shape[0], s.shape[1]) for s in h]
            r = torch.empty(6, dtype=torch.float32, device=DEV)
            _lib.minmax_(mats, cfg, torch.device(DEV), r)
            rngs.append(r)
        allr = torch.maximum(rngs[0], rngs[1])
        for h in halves:
            mats = [_lib.matrix_desc(s, s, s.shape[0], s.shape[1]) for s in h]
            _lib.fake_quant_with_range_(mats, cfg, torch.device(DEV), allr)
        for i in range(3):
            got = np.concatenate([bits(halves[0][i]), bits(halves[1][i])], axis=0)
            assert_bits(got, refs[i], "f16", f"matrix {i} sym={sym}")
    # single-process call of the public helper (world size 1)
    w = seeded(64, 512, "f16", 21).to(DEV)
    ref = O.int_fake_quant(bits(w), "f16", 4, True, -1)["out"]
    (o,) = sharded.pseudo_quantize_tensors_per_tensor_sharded([w], n_bits=4, zero_point=True)
    assert np.array_equal(bits(o), ref)


def test_host_buffer_pipeline():
    w = seeded(700, 1024, "f16", 4)
    src = w.clone().pin_memory()
    dst = torch.empty_like(src).pin_memory()
    cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, 128, bits=4, symmetric=False)
    st = ctypes.c_int32(-1)
    rc = _lib.lib().iwq_fake_quant_host(ctypes.byref(cfg), src.data_ptr(), dst.data_ptr(), 700, 1024, 0, 256 * 1024, ctypes.byref(st))
    _lib.check(rc)
    assert st.value == 0
    assert np.array_equal(bits(dst), O.int_fake_quant(bits(w), "f16", 4, True, 128)["out"])
    cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, -2, bits=8, symmetric=True)
    w2 = seeded(300, 3072, "f16", 6)
    buf = w2.clone().pin_memory()
    rc = _lib.lib().iwq_fake_quant_host(ctypes.byref(cfg), buf.data_ptr(), buf.data_ptr(), 300, 3072, 0, 1 << 20, ctypes.byref(st))
    _lib.check(rc)
    assert np.array_equal(bits(buf), O.int_fake_quant(bits(w2), "f16", 8, False, -2)["out"])
    # several host matrices through ONE pipeline (iwq_fake_quant_host_batched): ragged chunking, shapes mixed
    cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, 128, bits=4, symmetric=False)
    shapes_ = [(700, 1024), (3, 256), (129, 2048), (64, 128), (1000, 512)]
    ws = [seeded(r, c, "f16", 100 + i) for i, (r, c) in enumerate(shapes_)]
    hs = [x.clone().pin_memory() for x in ws]
    assert _lib.fake_quant_host_(hs, cfg, device=0, chunk_bytes=192 * 1024) == 0
    for x, h in zip(ws, hs):
        assert np.array_equal(bits(h), O.int_fake_quant(bits(x), "f16", 4, True, 128)["out"])
    nan = hs[2].clone().pin_memory()
    nan[5, 7] = float("nan")
    assert _lib.fake_quant_host_([hs[0], nan], cfg) & _lib.STATUS_NAN
    # scales and zero points come back through the same pipeline (what the module form stores: quant_linear.py:924-932),
    # bit-exact against the oracle, across ragged chunk boundaries; per-group and per-channel
    for grp, sym, nb in [(128, False, 4), (-2, False, 8), (64, True, 4)]:
        cfg = _lib.make_config(torch.float16, _lib.SCHEME_INT, grp, bits=nb, symmetric=sym)
        hs = [x.clone().pin_memory() for x in ws]
        sc = [torch.empty(_lib.num_segments(r, c, grp), dtype=torch.float16).pin_memory() for r, c in shapes_]
        zs = None if sym else [torch.empty(_lib.num_segments(r, c, grp), dtype=torch.float16).pin_memory() for r, c in shapes_]
        assert _lib.fake_quant_host_(hs, cfg, device=0, chunk_bytes=192 * 1024, scales=sc, zeros=zs) == 0
        for i, (x, h) in enumerate(zip(ws, hs)):
            ref = O.int_fake_quant(bits(x), "f16", nb, not sym, grp)
            assert np.array_equal(bits(h), ref["out"])
            assert_bits(bits(sc[i]), ref["scales"], "f16", f"host scales g{grp}")
            if not sym:
                assert_bits(bits(zs[i]), ref["zeros"], "f16", f"host zeros g{grp}")
    # the call leaves the caller's current device alone
    assert torch.cuda.current_device() == 0


# ------------------------------------------------------------------------------------------------
# full-size shapes: size-independent properties + sampled oracle rows
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(8192, 8192), (28672, 8192), (8192, 28672), (11008, 4096)])
def test_full_size_properties(shape):
    rows, cols = shape
    g = torch.Generator(device=DEV).manual_seed(1234)
    w = (torch.randn(rows, cols, generator=g, device=DEV) * 0.02).half()
    cfg = dict(n_bits=4, zero_point=True, q_group_size=128)
    out = quant_funcs.pseudo_quantize_tensor(w, **cfg)
    # (1) the three kernel variants agree bit for bit; in-place equals out-of-place
    for v in (1, 2):
        o2, _, _, _ = run_int(w, 4, False, 128, variant=v, want_codes=False)
        assert torch.equal(out.view(torch.int16), o2.view(torch.int16)), v
    t = w.clone()
    quant_funcs.pseudo_quantize_tensor(t, inplace=True, **cfg)
    assert torch.equal(out.view(torch.int16), t.view(torch.int16))
    # (2) at most 2^4 distinct values per group, error bounded by one step of the group's scale
    grp = out.reshape(-1, 128)
    step = (w.reshape(-1, 128).amax(1) - w.reshape(-1, 128).amin(1)).float() / 15
    assert ((grp.float() - w.reshape(-1, 128).float()).abs().amax(1) <= step * 1.01 + 1e-3).all()
    sample = grp[:: max(1, grp.shape[0] // 512)]
    assert max(len(torch.unique(r)) for r in sample) <= 16
    # (3) idempotent on its own grid when re-quantised with the same (scale, zero): checked through
    #     the oracle on sampled rows, which also pins the large-size launch geometry
    idx = torch.linspace(0, rows - 1, 16).long()
    ref = O.int_fake_quant(bits(w[idx.to(DEV)]), "f16", 4, True, 128)["out"]
    assert np.array_equal(bits(out[idx.to(DEV)]), ref)
    # (4) per-channel on the same matrix, sampled rows
    oc = quant_funcs.pseudo_quantize_tensor(w, n_bits=8, zero_point=True)
    ref = O.int_fake_quant(bits(w[idx.to(DEV)]), "f16", 8, True, -2)["out"]
    assert np.array_equal(bits(oc[idx.to(DEV)]), ref)


# ------------------------------------------------------------------------------------------------
# the reference's real callers: HuggingFace LLaMA / OPT modules walked by quantize_model
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("family", ["llama", "opt"])
def test_quantize_model_on_hf_architectures(family):
    """Random-init tiny LLaMA / OPT (the two families main.py targets) on the GPU in fp16, walked exactly as
    main.py:326 does; every Linear except lm_head is replaced, weights equal the oracle, generation-style
    forward still runs."""
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    if family == "llama":
        cfg = transformers.LlamaConfig(hidden_size=256, intermediate_size=768, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2, vocab_size=512, max_position_embeddings=128)
        model = transformers.LlamaForCausalLM(cfg)
        n_lin = 2 * 7
    else:
        cfg = transformers.OPTConfig(hidden_size=256, ffn_dim=1024, num_hidden_layers=2, num_attention_heads=4, vocab_size=512, max_position_embeddings=128, word_embed_proj_dim=256)
        model = transformers.OPTForCausalLM(cfg)
        n_lin = 2 * 6
    model = model.half().to(DEV).eval()
    before = {n: bits(m.weight.data) for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)}
    import copy

    twin = copy.deepcopy(model)  # same architecture, weights to be overwritten with the ORACLE's fake-quant output
    args = _Args(w_bit=4, w_group_size=128, w_symmetric=False)
    quant_wrapper.quantize_model(model, args)
    q = {n: m for n, m in model.named_modules() if isinstance(m, quant_linear.QuantLinear)}
    assert len(q) == n_lin
    assert not any("lm_head" in n for n in q)
    for n, m in q.items():
        ref = O.int_fake_quant(before[n], "f16", 4, True, 128)
        assert_bits(bits(m.weight.data), ref["out"], "f16", n)
        assert_bits(bits(m.scales.reshape(-1)), ref["scales"], "f16", n + " scales")
    ids = torch.randint(0, 512, (2, 16), device=DEV)
    with torch.no_grad():
        out = model(ids, labels=ids)
    assert out.logits.shape == (2, 16, 512) and torch.isfinite(out.logits).all()
    # "PPL delta" (north star): the model quantised by the CUDA path and the twin carrying the oracle's weights give
    # the SAME loss and logits bit for bit -- with identical weights perplexity cannot differ, on any dataset
    for n, m in twin.named_modules():
        if isinstance(m, torch.nn.Linear) and "lm_head" not in n:
            m.weight.data.copy_(tensor_of(O.int_fake_quant(before[n], "f16", 4, True, 128)["out"], "f16"))
    with torch.no_grad():
        ref_out = twin(ids, labels=ids)
    assert torch.equal(out.logits, ref_out.logits) and float(out.loss) == float(ref_out.loss)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_bfp_matches_reference_golden(golden, dt):
    """weight_format='bfp' (quant_linear.py:648-723) through QuantLinear.from_linear and quantize_model."""
    g = golden("golden_bfp")
    w = tensor_of(g[f"w_{dt}"], dt)
    for w_bit in (3, 4, 5, 6, 8, 12):
        for grp in (32, 64, 128, 512):
            lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, device=DEV, dtype=TDT[dt])
            lin.weight.data.copy_(w)
            ql = quant_linear.QuantLinear.from_linear(lin, w_bit=w_bit, w_group_size=grp, symmetric=True, weight_format="bfp")
            assert_bits(bits(ql.weight.data), g[f"{dt}_b{w_bit}_g{grp}"], dt, f"b{w_bit} g{grp}")
            assert ql.scales is None and ql.zeros is None and bool(ql.quantized)
    lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, device=DEV, dtype=TDT[dt])
    with pytest.raises(ValueError):
        quant_linear.QuantLinear.from_linear(lin, w_bit=4, w_group_size=-1, weight_format="bfp")  # quant_linear.py:650


def test_bfp_vs_oracle_large_and_batched():
    w = seeded(512, 4096, "f16", 17)
    ref = O.bfp_fake_quant(bits(w), "f16", 4, 128)["out"]
    model = torch.nn.Sequential(torch.nn.Linear(4096, 512, bias=False), torch.nn.Linear(4096, 512, bias=False)).half().to(DEV)
    for m in model:
        m.weight.data.copy_(w.to(DEV))
    quant_wrapper.quantize_model(model, _Args(w_bit=4, w_group_size=128, w_symmetric=True, w_format="bfp4"))
    for m in model:
        assert isinstance(m, quant_linear.QuantLinear) and m.weight_format == "bfp"
        assert np.array_equal(bits(m.weight.data), ref)


# ------------------------------------------------------------------------------------------------
# quant_dim = 1 (segments along the output dimension) and the aligned "approximate" decode
# ------------------------------------------------------------------------------------------------
def _from_linear(w, dt, **kw):
    lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, device=DEV, dtype=TDT[dt])
    lin.weight.data.copy_(w)
    return quant_linear.QuantLinear.from_linear(lin, **kw)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_quant_dim1_matches_reference_golden(golden, dt):
    g = golden("golden_dim1_approx")
    w = tensor_of(g[f"w_{dt}"], dt)
    quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)
    for grp in (32, 128):
        for sym in (False, True):
            tag = "sym" if sym else "asym"
            for name, kw in (("int4", dict(w_bit=4, weight_format="int")), ("fp4", dict(w_bit=4, weight_format="fp4")), ("fp8", dict(w_bit=8, weight_format="fp8"))):
                key = f"{dt}_{name}_{tag}_g{grp}_d1"
                ql = _from_linear(w, dt, w_group_size=grp, symmetric=sym, quant_dim=1, **kw)
                assert_bits(bits(ql.weight.data), g[key + "_out"], dt, key)
                sd = dt if name == "int4" else "f16"
                assert_bits(bits(ql.scales.reshape(-1)), g[key + "_scales"], sd, key + " scales")
                if sym:
                    assert ql.zeros is None
                else:
                    assert np.array_equal(O.from_bits(bits(ql.zeros.reshape(-1)), sd), O.from_bits(g[key + "_zeros"], sd)), key + " zeros"
        ql = _from_linear(w, dt, w_bit=8, w_group_size=grp, symmetric=False, weight_format="int", quant_dim=1)
        assert_bits(bits(ql.weight.data), g[f"{dt}_int8_asym_g{grp}_d1_out"], dt, f"int8 g{grp}")
        ql = _from_linear(w, dt, w_bit=4, w_group_size=grp, symmetric=True, weight_format="bfp", quant_dim=1)
        assert_bits(bits(ql.weight.data), g[f"{dt}_bfp4_g{grp}_d1_out"], dt, f"bfp4 g{grp}")
    ql = _from_linear(w, dt, w_bit=4, w_group_size=-2, symmetric=False, weight_format="int", quant_dim=1)
    assert_bits(bits(ql.weight.data), g[f"{dt}_int4_asym_pc_d1_out"], dt, "per-channel d1")
    assert_bits(bits(ql.scales.reshape(-1)), g[f"{dt}_int4_asym_pc_d1_scales"], dt, "per-channel d1 scales")
    lin = torch.nn.Linear(64, 100, bias=False, device=DEV, dtype=TDT[dt])
    with pytest.raises(AssertionError):
        quant_linear.QuantLinear.from_linear(lin, w_group_size=32, quant_dim=1)  # 100 % 32 != 0 on the transposed shape


SINGLE_APPROX = [("e1m2_s", "fp4", 1, 2, 1, 1, 0), ("e2m1_s", "fp4", 2, 1, 1, 1, 0), ("e2m3_s", "fp6", 2, 3, 4, 7, 2), ("e3m2_s", "fp6", 3, 2, 4, 7, 2), ("e4m3_s", "fp8", 4, 3, 12, 15, 1)]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("case", SINGLE_APPROX, ids=[c[0] for c in SINGLE_APPROX])
def test_single_approximate_matches_reference_golden(golden, dt, case):
    name, wf, E, M, hs, hf, tp = case
    g = golden("golden_dim1_approx")
    w = tensor_of(g[f"w_{dt}"], dt)
    quant_linear.configure_fp_formats(**{f"{wf}_exp_bits": E, f"{wf}_mantissa_bits": M})
    try:
        kw = {f"{wf}_hi_align_start": hs, f"{wf}_hi_align_exp_field": hf, f"{wf}_tail_pad_bits": tp}
        for qd, grp in ((0, 32), (1, 128)):
            ql = _from_linear(w, dt, w_bit=4, w_group_size=grp, symmetric=True, weight_format=wf, approximate=True, quant_dim=qd, **kw)
            key = f"{dt}_approx_{name}_g{grp}_d{qd}"
            assert_bits(bits(ql.weight.data), g[key + "_out"], dt, key)
            assert np.array_equal(bits(ql.scales.reshape(-1)), g[key + "_scales"]), key + " scales"
            assert ql.zeros is None and ql.approximate is True
        with pytest.raises(ValueError):
            _from_linear(w, dt, w_group_size=-1, weight_format=wf, approximate=True)  # quant_linear.py:475-476
        with pytest.raises(NotImplementedError):
            _from_linear(w, dt, w_group_size=32, weight_format="int", approximate=True)  # quant_linear.py:601
    finally:
        quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_quant_dim1_vs_oracle_larger(dt):
    """column-group kernel on shapes with several row blocks, a ragged last column tile and whole-column segments"""
    for rows, cols, grp, nb, sym in [(1024, 520, 128, 4, False), (512, 1032, 64, 4, True), (384, 264, -2, 8, False), (2048, 256, 256, 3, False),
                                     (256, 72, 32, 4, True), (1024, 136, 512, 4, False), (2048, 64, 1024, 4, False), (96, 40, 8, 5, False)]:
        w = seeded(rows, cols, dt, rows + cols + 1)
        ref = O.along_output_dim(O.int_fake_quant, bits(w), dt, nb, not sym, grp)
        G = _lib.num_segments(rows, cols, grp, 1)
        wd = w.to(DEV)
        scales = torch.empty(G, dtype=TDT[dt], device=DEV)
        zeros = torch.empty(G, dtype=TDT[dt], device=DEV)
        cfg = _lib.make_config(TDT[dt], _lib.SCHEME_INT, grp, bits=nb, symmetric=sym, quant_dim=1)
        _lib.fake_quant_(wd, wd, cfg, rows, cols, scales=scales, zeros=zeros)
        tag = f"{dt} [{rows},{cols}] g{grp} b{nb} sym={sym}"
        assert_bits(bits(wd), ref["out"], dt, tag)
        assert_bits(bits(scales), ref["scales"], dt, tag + " scales")
        assert_bits(bits(zeros), ref["zeros"], dt, tag + " zeros")
    w = seeded(512, 768, dt, 77)
    ref = O.along_output_dim(O.fp_approx_fake_quant, bits(w), dt, 4, 3, 128, 12, 15, 1, False)
    wd = w.to(DEV)
    cfg = _lib.make_config(TDT[dt], _lib.SCHEME_FP_CODEC, 128, exp_bits=4, mant_bits=3, symmetric=True, quant_dim=1, approx=1, hi_align_start=12, hi_align_exp_field=15, tail_pad_bits=1)
    _lib.fake_quant_(wd, wd, cfg, 512, 768)
    assert_bits(bits(wd), ref["out"], dt, "approx e4m3 d1")


DOUBLE_APPROX = [("e2m1_d", "fp4", 2, 1, 0, 2, 1), ("e2m3_d", "fp6", 2, 3, 0, 2, -1), ("e3m2_d", "fp6", 3, 2, 4, 6, 0), ("e4m3_d", "fp8", 4, 3, 13, 15, -1), ("e4m3_d0", "fp8", 4, 3, 12, 15, 0)]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
@pytest.mark.parametrize("case", DOUBLE_APPROX, ids=[c[0] for c in DOUBLE_APPROX])
def test_double_approximate_matches_reference_golden(golden, dt, case):
    """--approximate --double_approximate --quant_dim 1 (scripts/opt-6.7b/scripts_4.sh): the reference's 4-code
    neighbourhood decode, int8 overflow included."""
    name, wf, E, M, hs, hf, tp = case
    g = golden("golden_dim1_approx")
    w = tensor_of(g[f"w_{dt}"], dt)
    quant_linear.configure_fp_formats(**{f"{wf}_exp_bits": E, f"{wf}_mantissa_bits": M})
    try:
        kw = {f"{wf}_hi_align_start": hs, f"{wf}_hi_align_exp_field": hf, f"{wf}_tail_pad_bits": tp}
        ql = _from_linear(w, dt, w_bit=4, w_group_size=128, symmetric=True, weight_format=wf, approximate=True, double_approximate=True, quant_dim=1, **kw)
        key = f"{dt}_approx_{name}_g128_d1"
        assert_bits(bits(ql.weight.data), g[key + "_out"], dt, key)
        assert np.array_equal(bits(ql.scales.reshape(-1)), g[key + "_scales"]), key + " scales"
        # quant_dim = 0 (quant_linear.py:484-488): groups of four = the same offset of four consecutive segments
        ql = _from_linear(w, dt, w_bit=4, w_group_size=32, symmetric=True, weight_format=wf, approximate=True, double_approximate=True, quant_dim=0, **kw)
        key0 = f"{dt}_approx_{name}_g32_d0"
        assert_bits(bits(ql.weight.data), g[key0 + "_out"], dt, key0)
        assert np.array_equal(bits(ql.scales.reshape(-1)), g[key0 + "_scales"]), key0 + " scales"
        for rows0, cols0, grp0 in [(96, 1024, 128), (33, 512, 64), (20, 768, 256), (64, 96, 32)]:  # ragged last unit, quads spanning rows
            w0 = seeded(rows0, cols0, dt, 11 + E + grp0)
            ref0 = O.fp_approx_fake_quant(bits(w0), dt, E, M, grp0, hs, hf, tp, True)
            ql = _from_linear(w0.to(DEV), dt, w_bit=4, w_group_size=grp0, symmetric=True, weight_format=wf, approximate=True, double_approximate=True, quant_dim=0, **kw)
            assert_bits(bits(ql.weight.data), ref0["out"], dt, f"{key0} [{rows0},{cols0}] g{grp0}")
        with pytest.raises(NotImplementedError):  # 3 segments: the reference's groups of four would wrap around
            _from_linear(seeded(3, 32, dt, 1).to(DEV), dt, w_bit=4, w_group_size=32, symmetric=True, weight_format=wf, approximate=True, double_approximate=True, quant_dim=0, **kw)
        # larger, against the oracle: 16 row blocks, ragged column tile
        wl = seeded(2048, 264, dt, 5 + E)
        ref = O.along_output_dim(O.fp_approx_fake_quant, bits(wl), dt, E, M, 128, hs, hf, tp, True)
        ql = _from_linear(wl.to(DEV), dt, w_bit=4, w_group_size=128, symmetric=True, weight_format=wf, approximate=True, double_approximate=True, quant_dim=1, **kw)
        assert_bits(bits(ql.weight.data), ref["out"], dt, key + " large")
    finally:
        quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)


# ------------------------------------------------------------------------------------------------
# round 2: shapes the vectorised kernels do not take directly, and full-size properties of the column path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_odd_shapes_match_oracle(dt):
    """The reference takes ANY shape (quant_funcs.py:10-15): tensors whose numel / rows / groups are not whole 16-byte
    vectors ([3,5], [7,129] per tensor and per channel, group 12), 1-D and 3-D inputs, and the in-place contract on a
    non-contiguous view (the caller's storage is mutated, quant_funcs.py:44-46)."""
    cases = [((3, 5), dict(per_tensor=True), -1), ((7, 129), dict(per_tensor=True), -1), ((7, 129), dict(), -2), ((5, 36), dict(q_group_size=12), 12),
             ((1, 8), dict(per_tensor=True), -1), ((2, 3, 40), dict(q_group_size=20), 20), ((6, 1000), dict(), -2), ((1, 3), dict(), -2)]
    for shape, kw, grp in cases:
        for zp in (True, False):
            n = int(np.prod(shape))
            w = seeded(n, 1, dt, 7 + n).reshape(shape)
            ref = O.int_fake_quant(bits(w).reshape(-1, shape[-1]), dt, 4, zp, grp)["out"].reshape(shape)
            out = quant_funcs.pseudo_quantize_tensor(w.to(DEV), n_bits=4, zero_point=zp, **kw)
            assert out.shape == w.shape
            assert_bits(bits(out), ref, dt, f"{dt} {shape} {kw} zp={zp}")
    # non-contiguous in place: every second column of a wider buffer
    base = seeded(16, 512, dt, 99).to(DEV)
    view = base[:, ::2]
    want = O.int_fake_quant(bits(view), dt, 4, True, 128)["out"]
    untouched = bits(base[:, 1::2])
    ret = quant_funcs.pseudo_quantize_tensor(view, n_bits=4, zero_point=True, q_group_size=128, inplace=True)
    assert ret.data_ptr() == view.data_ptr()
    assert_bits(bits(base[:, ::2]), want, dt, "strided in place")
    assert np.array_equal(bits(base[:, 1::2]), untouched)
    # the wrappers on odd shapes (quant_funcs.py:49-77)
    t = seeded(5, 77, dt, 3).to(DEV)
    assert_bits(bits(quant_funcs.quantize_weight_per_tensor_absmax(t, n_bits=8)), O.int_fake_quant(bits(t), dt, 8, False, -1)["out"], dt, "per-tensor absmax odd")
    assert_bits(bits(quant_funcs.quantize_activation_per_token_absmax(t.reshape(5, 7, 11), n_bits=8)), O.int_fake_quant(bits(t).reshape(35, 11), dt, 8, True, -2)["out"].reshape(5, 7, 11), dt, "per-token odd")


@pytest.mark.parametrize("shape", [(8192, 8192), (11008, 4096), (4096, 11008)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_quant_dim1_equals_transposed_quant_dim0_full_size(shape):
    """Size-independent property of the column-grouped path at LLaMA sizes (quant_linear.py:478-483 transposes, quantises along
    the last dim and transposes back): quant_dim = 1 on W == (quant_dim = 0 on W^T)^T, bit for bit, including the order of
    the emitted scales and zero points -- col_wide_kernel against seg4_tma_kernel, two kernels that share no code path but
    the element stage."""
    rows, cols = shape
    g = torch.Generator(device=DEV).manual_seed(4321)
    w = (torch.randn(rows, cols, generator=g, device=DEV) * 0.02).half()
    wt = w.t().contiguous()
    for name, mk in [("int4 g128 asym", lambda qd: _lib.make_config(torch.float16, _lib.SCHEME_INT, 128, bits=4, symmetric=False, quant_dim=qd)),
                     ("int3 g64 sym", lambda qd: _lib.make_config(torch.float16, _lib.SCHEME_INT, 64, bits=3, symmetric=True, quant_dim=qd)),
                     ("fp4 e2m1 g128", lambda qd: _lib.make_config(torch.float16, _lib.SCHEME_FP_CODEC, 128, exp_bits=2, mant_bits=1, symmetric=True, quant_dim=qd)),
                     ("fp8 e4m3 g64 approx", lambda qd: _lib.make_config(torch.float16, _lib.SCHEME_FP_CODEC, 64, exp_bits=4, mant_bits=3, symmetric=True, quant_dim=qd, approx=1, hi_align_start=12, hi_align_exp_field=15, tail_pad_bits=1))]:
        grp = 64 if "g64" in name else 128
        G = rows * cols // grp
        o1, s1, z1 = torch.empty_like(w), torch.empty(G, dtype=w.dtype, device=DEV), torch.zeros(G, dtype=w.dtype, device=DEV)
        o0, s0, z0 = torch.empty_like(wt), torch.empty(G, dtype=w.dtype, device=DEV), torch.zeros(G, dtype=w.dtype, device=DEV)
        asym = "asym" in name
        _lib.fake_quant_(w, o1, mk(1), rows, cols, scales=s1, zeros=z1 if asym else None)
        _lib.fake_quant_(wt, o0, mk(0), cols, rows, scales=s0, zeros=z0 if asym else None)
        assert torch.equal(o1.view(torch.int16), o0.t().contiguous().view(torch.int16)), name
        assert torch.equal(s1.view(torch.int16), s0.view(torch.int16)), name + " scales"
        assert torch.equal(z1.view(torch.int16), z0.view(torch.int16)), name + " zeros"
    assert _lib.read_and_clear_status(torch.device(DEV)) & _lib.STATUS_NAN == 0


ODD_GOLDEN = [((3, 5), dict(per_tensor=True)), ((7, 129), dict(per_tensor=True)), ((7, 129), dict()), ((5, 36), dict(q_group_size=12)),
              ((1, 8), dict(per_tensor=True)), ((2, 3, 40), dict(q_group_size=20)), ((6, 1000), dict()), ((1, 3), dict()), ((3, 51216), dict())]


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_odd_shapes_match_reference_golden(golden, dt):
    """pseudo_quantize_tensor on the shapes of tests/golden/golden_odd.npz against the reference's own outputs: numel not a multiple of 8
    per tensor, 129- and 1000-wide rows per channel, group 12 / 20, a 3-D input, and a per-channel row of 51216 elements (longer than
    any ring stage: it goes through the two-pass kernels, one job per row)."""
    g = golden("golden_odd")
    for ci, (shape, kw) in enumerate(ODD_GOLDEN):
        w = tensor_of(g[f"w_{dt}_{ci}"], dt)
        assert tuple(w.shape) == shape
        for nb, zp in ((4, True), (4, False), (8, True), (3, False)):
            out = quant_funcs.pseudo_quantize_tensor(w, n_bits=nb, zero_point=zp, **kw)
            assert_bits(bits(out), g[f"{dt}_{ci}_b{nb}_{'asym' if zp else 'sym'}"], dt, f"{dt} {shape} {kw} b{nb} zp={zp}")


@pytest.mark.parametrize("dt", ["f16", "bf16"])
def test_quant_dim1_wide_tiles_match_reference_golden(golden, dt):
    """QuantLinear.from_linear(..., quant_dim=1) on the [256, 264] fixture minted from the reference (golden_dim1_wide.npz): two 128-row
    tiles and a ragged second column tile of col_wide_kernel -- INT (scales and zero points in the reference's order), BFP, exact FP4 /
    asymmetric FP8, and the single-approximate decode of FP4 / FP6 / FP8 at group 64 and 128 (scripts/*/scripts_*.sh)."""
    g = golden("golden_dim1_wide")
    w = tensor_of(g[f"w_{dt}"], dt)
    quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)

    def check(key, **kw):
        ql = _from_linear(w, dt, quant_dim=1, **kw)
        assert_bits(bits(ql.weight.data), g[f"{dt}_{key}_out"], dt, f"{dt} {key}")
        if f"{dt}_{key}_scales" in g.files:
            assert np.array_equal(bits(ql.scales.reshape(-1)), g[f"{dt}_{key}_scales"]), f"{dt} {key} scales"
        if f"{dt}_{key}_zeros" in g.files:
            assert np.array_equal(bits(ql.zeros.reshape(-1)), g[f"{dt}_{key}_zeros"]), f"{dt} {key} zeros"

    check("int4_asym_g128", w_bit=4, w_group_size=128, symmetric=False, weight_format="int")
    check("int3_sym_g64", w_bit=3, w_group_size=64, symmetric=True, weight_format="int")
    check("bfp4_g128", w_bit=4, w_group_size=128, symmetric=True, weight_format="bfp")
    check("fp4_sym_g128", w_bit=4, w_group_size=128, symmetric=True, weight_format="fp4")
    check("fp8_asym_g64", w_bit=8, w_group_size=64, symmetric=False, weight_format="fp8")
    try:
        for name, wf, E, M, hs, hf, tp in SINGLE_APPROX:
            if name in ("e1m2_s", "e2m3_s"):
                continue
            quant_linear.configure_fp_formats(**{f"{wf}_exp_bits": E, f"{wf}_mantissa_bits": M})
            kw = {f"{wf}_hi_align_start": hs, f"{wf}_hi_align_exp_field": hf, f"{wf}_tail_pad_bits": tp}
            for grp in (64, 128):
                check(f"approx_{name}_g{grp}", w_bit=4, w_group_size=grp, symmetric=True, weight_format=wf, approximate=True, **kw)
            quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)
    finally:
        quant_linear.configure_fp_formats(2, 1, 3, 2, 4, 3)
