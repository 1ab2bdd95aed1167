// TEST-ONLY CPU emulator: runs the *host instantiation* of the product's per-segment / per-element
// math (iron_weight_only_quant_b200/csrc/iwq_math.cuh, iwq_consts.h) over whole matrices so that the
// arithmetic (clamp fusion, magic-number rounding, reciprocal+FMA quotient, minifloat table) can be
// checked against the reference's golden vectors on the GPU-less build box.  The product never loads
// this library; it is built and used by tests/test_host_emulation.py only.
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../iron_weight_only_quant_b200/csrc/iwq_consts.h"

using namespace iwq;

template <int DT>
static void seg_range(const uint16_t* w, int64_t n, bool sym, float* mx, float* mn) {
  float a = sym ? 0.0f : -INFINITY, b = INFINITY;
  bool nan = false;
  for (int64_t i = 0; i < n; ++i) {
    float v = DType<DT>::to_f32(w[i]);
    if (v != v) nan = true;
    if (sym) {
      a = fmaxf(a, fabsf(v));
    } else {
      a = fmaxf(a, v);
      b = fminf(b, v);
    }
  }
  if (nan) a = b = NAN;
  *mx = a;
  *mn = sym ? -a : b;  // the kernels unpack {absmax, -absmax} for symmetric ranges
}

template <int DT>
static int run(const iwq_config_t* cfg, const QuantConsts& c, int kind, const uint16_t* w, uint16_t* out, uint16_t* scales, uint16_t* zeros,
               int64_t rows, int64_t cols) {
  using T2 = typename DType<DT>::T2;
  const int64_t total = rows * cols;
  int64_t seg = cfg->group > 0 ? cfg->group : (cfg->group == IWQ_GROUP_PER_TENSOR ? total : cols);
  const bool by_col = cfg->quant_dim == 1 && cfg->group != IWQ_GROUP_PER_TENSOR;
  if (!by_col && cfg->group > 0 && cols % cfg->group) return IWQ_ERR_INDIVISIBLE;
  if (!by_col && seg % 2) return IWQ_ERR_UNSUPPORTED;
  const bool sym = c.symmetric != 0;
  // the minifloat step table, built by the same host code the library uses
  std::vector<FpStep> steps;
  if (kind == KIND_FP) {
    steps.resize(c.lut_n);
    if (!build_step_table(DT, c, steps.data())) return IWQ_ERR_UNSUPPORTED;
  }
  if (cfg->quant_dim == 1 && cfg->group != IWQ_GROUP_PER_TENSOR) {
    // column segments, processed as PAIRS of adjacent columns like col_kernel does
    const int64_t g = cfg->group > 0 ? cfg->group : rows;
    if (rows % g) return IWQ_ERR_INDIVISIBLE;
    if (cols % 2) return IWQ_ERR_UNSUPPORTED;
    const int64_t n_rb = rows / g;
    std::vector<uint16_t> col(g);
    if (c.approx == 2) {
      // double-approximate decode: four consecutive row blocks of a column form the code neighbourhoods,
      // mirrored from col_double_kernel
      if (n_rb % 4) return IWQ_ERR_UNSUPPORTED;
      FpQ q1;
      q1.S = 1.f; q1.r = 1.f; q1.mid = 0.f; q1.mid_h = 0.f; q1.bad = false; q1.slow_div = false;
      FpTable<DT, true> grid;
      grid.setup2(q1, q1, c, steps.data());
      for (int64_t rb4 = 0; rb4 < n_rb / 4; ++rb4) {
        for (int64_t cc = 0; cc < cols; cc += 2) {
          FpQ qs[4][2];
          for (int b = 0; b < 4; ++b)
            for (int h = 0; h < 2; ++h) {
              for (int64_t k = 0; k < g; ++k) col[k] = w[((rb4 * 4 + b) * g + k) * cols + cc + h];
              float mx, mn;
              seg_range<DT>(col.data(), g, true, &mx, &mn);
              qs[b][h] = fp_segment_params<DT>(mx, mn, c);
              if (scales) scales[(cc + h) * n_rb + rb4 * 4 + b] = DType<DT>::from_f32(qs[b][h].S);
            }
          for (int64_t k = 0; k < g; ++k) {
            uint32_t clo[4], chi[4];
            int64_t idx[4];
            for (int b = 0; b < 4; ++b) {
              idx[b] = ((rb4 * 4 + b) * g + k) * cols + cc;
              const uint32_t w2 = (uint32_t)w[idx[b]] | ((uint32_t)w[idx[b] + 1] << 16);
              const bool slow = qs[b][0].slow_div || qs[b][1].slow_div;
              const float2 S = make_float2(qs[b][0].S, qs[b][1].S), R = make_float2(qs[b][0].r, qs[b][1].r);
              const T2 n = slow ? ieee_quot2<DT>(from_u32<T2>(w2), S) : fast_quot2<DT>(from_u32<T2>(w2), S, R);
              const uint32_t gv = as_u32(grid.grid_steps(n, w2));
              clo[b] = fp_code_of<DT>((uint16_t)(gv & 0xFFFFu), c.exp_bits, c.mant_bits);
              chi[b] = fp_code_of<DT>((uint16_t)(gv >> 16), c.exp_bits, c.mant_bits);
            }
            uint16_t dlo[4], dhi[4];
            double_approx_decode4<DT>(clo, c, dlo);
            double_approx_decode4<DT>(chi, c, dhi);
            for (int b = 0; b < 4; ++b) {
              const T2 dec = from_u32<T2>((uint32_t)dlo[b] | ((uint32_t)dhi[b] << 16));
              const uint32_t o = as_u32(__hmul2(dec, DType<DT>::from_f32x2(qs[b][0].S, qs[b][1].S)));
              out[idx[b]] = (uint16_t)(o & 0xFFFF);
              out[idx[b] + 1] = (uint16_t)(o >> 16);
            }
          }
        }
      }
      return 0;
    }
    for (int64_t rb = 0; rb < n_rb; ++rb) {
      for (int64_t cc = 0; cc < cols; cc += 2) {
        float mxv[2], mnv[2];
        for (int h = 0; h < 2; ++h) {
          for (int64_t k = 0; k < g; ++k) col[k] = w[(rb * g + k) * cols + cc + h];
          seg_range<DT>(col.data(), g, sym, &mxv[h], &mnv[h]);
        }
        float sa = 0, za = 0, sb = 0, zb = 0;
        IntFast<DT, false, true> fi;
        IntFast<DT, true, true> fw;
        FpTable<DT, true> ff;
        BfpApply<DT, true> fb;
        IntQ qa, qb;
        if (kind == KIND_FP) {
          FpQ a = fp_segment_params<DT>(mxv[0], mnv[0], c), b = fp_segment_params<DT>(mxv[1], mnv[1], c);
          ff.setup2(a, b, c, steps.data());
          sa = a.S; za = a.mid; sb = b.S; zb = b.mid;
        } else if (kind == KIND_BFP) {
          fb.setup2(bfp_segment_params<DT>(mxv[0], c), bfp_segment_params<DT>(mxv[1], c), c);
        } else {
          qa = int_segment_params<DT>(mxv[0], mnv[0], c, sym);
          qb = int_segment_params<DT>(mxv[1], mnv[1], c, sym);
          fi.setup2(qa, qb);
          fw.setup2(qa, qb);
          sa = qa.s; za = qa.z; sb = qb.s; zb = qb.z;
        }
        for (int64_t k = 0; k < g; ++k) {
          const int64_t idx = (rb * g + k) * cols + cc;
          const uint32_t w2 = (uint32_t)w[idx] | ((uint32_t)w[idx + 1] << 16);
          uint32_t o;
          T2 tmp;
          if (kind == KIND_FP) {
            o = ff.slow_div ? ff.template apply<true>(w2, &tmp) : ff.template apply<false>(w2, &tmp);
          } else if (kind == KIND_BFP) {
            o = fb.apply(w2);
          } else if (kind == KIND_INT_GENERIC) {
            float code;
            o = (uint32_t)int_generic_elem<DT>(w[idx], qa, c, &code) | ((uint32_t)int_generic_elem<DT>(w[idx + 1], qb, c, &code) << 16);
          } else {
            o = kind == KIND_INT_WIDE ? fw.apply(w2, &tmp) : fi.apply(w2, &tmp);
          }
          out[idx] = (uint16_t)(o & 0xFFFF);
          out[idx + 1] = (uint16_t)(o >> 16);
        }
        if (scales) { scales[cc * n_rb + rb] = DType<DT>::from_f32(sa); scales[(cc + 1) * n_rb + rb] = DType<DT>::from_f32(sb); }
        if (zeros) { zeros[cc * n_rb + rb] = DType<DT>::from_f32(za); zeros[(cc + 1) * n_rb + rb] = DType<DT>::from_f32(zb); }
      }
    }
    return 0;
  }
  for (int64_t s0 = 0, g = 0; s0 < total; s0 += seg, ++g) {
    float mx, mn;
    seg_range<DT>(w + s0, seg, sym, &mx, &mn);
    float sc = 0, zp = 0;
    if (kind == KIND_BFP) {
      BfpApply<DT> f;
      f.setup(bfp_segment_params<DT>(mx, c), c);
      for (int64_t i = 0; i < seg; i += 2) {
        uint32_t w2 = (uint32_t)w[s0 + i] | ((uint32_t)w[s0 + i + 1] << 16);
        uint32_t o = f.apply(w2);
        out[s0 + i] = (uint16_t)(o & 0xFFFF);
        out[s0 + i + 1] = (uint16_t)(o >> 16);
      }
    } else if (kind == KIND_FP) {
      FpQ q = fp_segment_params<DT>(mx, mn, c);
      FpTable<DT> f;
      f.setup(q, c, steps.data());
      for (int64_t i = 0; i < seg; i += 2) {
        uint32_t w2 = (uint32_t)w[s0 + i] | ((uint32_t)w[s0 + i + 1] << 16);
        T2 v;
        uint32_t o = q.slow_div ? f.template apply<true>(w2, &v) : f.template apply<false>(w2, &v);
        out[s0 + i] = (uint16_t)(o & 0xFFFF);
        out[s0 + i + 1] = (uint16_t)(o >> 16);
      }
      sc = q.S;
      zp = q.mid;
    } else {
      IntQ q = int_segment_params<DT>(mx, mn, c, sym);
      if (kind == KIND_INT_GENERIC) {
        for (int64_t i = 0; i < seg; ++i) {
          float code;
          out[s0 + i] = int_generic_elem<DT>(w[s0 + i], q, c, &code);
        }
      } else {
        IntFast<DT, false> fn;
        IntFast<DT, true> fw;
        fn.setup(q);
        fw.setup(q);
        for (int64_t i = 0; i < seg; i += 2) {
          uint32_t w2 = (uint32_t)w[s0 + i] | ((uint32_t)w[s0 + i + 1] << 16);
          T2 d;
          uint32_t o = kind == KIND_INT_WIDE ? fw.apply(w2, &d) : fn.apply(w2, &d);
          out[s0 + i] = (uint16_t)(o & 0xFFFF);
          out[s0 + i + 1] = (uint16_t)(o >> 16);
        }
      }
      sc = q.s;
      zp = q.z;
    }
    if (scales) scales[g] = DType<DT>::from_f32(sc);
    if (zeros) zeros[g] = DType<DT>::from_f32(zp);
  }
  return 0;
}

extern "C" int emu_fake_quant(const iwq_config_t* cfg, const uint16_t* w, uint16_t* out, uint16_t* scales, uint16_t* zeros, int64_t rows,
                              int64_t cols, int force_kind) {
  QuantConsts c;
  int kind = 0;
  std::string err;
  int rc = build_quant_consts(cfg, &c, &kind, &err);
  if (rc) return rc;
  if (force_kind >= 0) kind = force_kind;
  return cfg->dtype == IWQ_F16 ? run<IWQ_F16>(cfg, c, kind, w, out, scales, zeros, rows, cols)
                               : run<IWQ_BF16>(cfg, c, kind, w, out, scales, zeros, rows, cols);
}

template <int DT>
static void roundtrip(const QuantConsts& c, const uint16_t* x, uint16_t* y, int64_t n) {
  using T2 = typename DType<DT>::T2;
  FpQ q;
  q.S = 1.0f; q.r = 1.0f; q.mid = 0.f; q.mid_h = 0.f; q.bad = false; q.slow_div = false;
  FpTable<DT> f;
  f.setup(q, c);
  for (int64_t i = 0; i < n; i += 2) {
    uint32_t x2 = (uint32_t)x[i] | ((uint32_t)x[i + 1] << 16);
    uint32_t o = as_u32(f.grid(from_u32<T2>(x2), x2, c.tbl));
    y[i] = (uint16_t)(o & 0xFFFF);
    y[i + 1] = (uint16_t)(o >> 16);
  }
}

extern "C" int emu_fp_roundtrip(int dtype, int flavour, int E, int M, const uint16_t* x, uint16_t* y, int64_t n) {
  iwq_config_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.dtype = dtype; cfg.scheme = flavour; cfg.exp_bits = E; cfg.mant_bits = M; cfg.symmetric = 1; cfg.group = 128;
  QuantConsts c;
  int kind = 0;
  std::string err;
  int rc = build_quant_consts(&cfg, &c, &kind, &err);
  if (rc) return rc;
  if (dtype == IWQ_F16) roundtrip<IWQ_F16>(c, x, y, n); else roundtrip<IWQ_BF16>(c, x, y, n);
  return 0;
}


// Exhaustive check of the table decomposition the device uses for the double-approximate decode (build_double_tables +
// double_approx_target2) against the direct per-group decode (double_approx_decode4, itself pinned on the reference's
// golden vectors): every 4-tuple of magnitude codes and signs when that is <= 2^24 tuples, else `samples` random ones.
// Returns the number of mismatching values (or a negative error code).
template <int DT>
static long long double_table_check_t(const QuantConsts& c, long long samples) {
  const int E = c.exp_bits, M = c.mant_bits, nm = 1 << (E + M), nc = 2 * nm;
  std::vector<FpStep> steps(c.lut_n);
  std::vector<uint16_t> val(kDblValEntries);
  if (!build_double_tables_t<DT>(c, steps.data(), val.data())) return -1;
  const uint32_t hf2 = ((uint32_t)c.hi_field << 8) * 0x00010001u, maxe2 = ((uint32_t)((1 << E) - 1) << 8) * 0x00010001u;
  long long bad = 0;
  const long long total = (long long)nc * nc * nc * nc;
  const bool exhaustive = total <= (1ll << 24);
  unsigned long long rs = 88172645463325252ull;
  const long long n = exhaustive ? total : samples;
  for (long long it = 0; it < n; ++it) {
    long long k = it;
    if (!exhaustive) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; k = (long long)(rs % (unsigned long long)total); }
    uint32_t code[4], info[4];
    for (int i = 0; i < 4; ++i) {
      const uint32_t cc = (uint32_t)(k % nc);
      k /= nc;
      const uint32_t mag = cc % nm, sign = cc / nm;
      code[i] = mag | (sign << (E + M));
      const uint32_t inf = double_approx_info(mag, c);
      info[i] = inf | (inf << 16);  // the same group in both halves
    }
    uint16_t want[4];
    double_approx_decode4<DT>(code, c, want);
    const uint32_t target = double_approx_target2<DT>(info, hf2, maxe2);
    for (int i = 0; i < 4; ++i) {
      const uint32_t sign7 = ((code[i] >> (E + M)) & 1u) << 7;
      const uint32_t idx = ((info[i] & kDblCodeMask) | target | (sign7 * 0x00010001u)) & 0xFFFFu;
      const uint32_t idx_hi = ((info[i] & kDblCodeMask) | target | (sign7 * 0x00010001u)) >> 16;
      if (val[idx] != want[i] || idx_hi != idx) ++bad;
    }
  }
  return bad;
}

extern "C" long long emu_double_table_check(int dtype, int E, int M, int hi_start, int hi_field, int tail_pad, long long samples) {
  iwq_config_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.dtype = dtype; cfg.scheme = IWQ_SCHEME_FP_CODEC; cfg.exp_bits = E; cfg.mant_bits = M; cfg.symmetric = 1; cfg.group = 128;
  cfg.quant_dim = 1; cfg.approx = 2; cfg.hi_align_start = hi_start; cfg.hi_align_exp_field = hi_field; cfg.tail_pad_bits = tail_pad;
  QuantConsts c;
  int kind = 0;
  std::string err;
  int rc = build_quant_consts(&cfg, &c, &kind, &err);
  if (rc) return rc;
  return dtype == IWQ_F16 ? double_table_check_t<IWQ_F16>(c, samples) : double_table_check_t<IWQ_BF16>(c, samples);
}
