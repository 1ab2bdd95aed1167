// iwq_consts.h -- host-side construction of the launch constants (QuantConsts), including the
// per-binade minifloat table.  Header-only so that the test-only CPU emulator (tests/host/emulate.cu)
// builds the same table as the library.
#pragma once

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "iwq_math.cuh"

namespace iwq {

enum { KIND_INT_FAST = 0, KIND_INT_WIDE = 1, KIND_FP = 2, KIND_INT_GENERIC = 3, KIND_BFP = 4 };

static inline float round_to_dtype(int dtype, float x) { return dtype == IWQ_F16 ? rD<IWQ_F16>(x) : rD<IWQ_BF16>(x); }

// number of D values a in [2^(k-1), 2^k) whose log2, rounded to D, reaches k (quirk Q3):
// log2(a) > midpoint(pred_D(k), k).  Closed form in double precision; the margin to every decision
// boundary is > 64 fp32 ulps (tests/test_oracle_golden.py::test_log2_rounding_margin).
static inline int log2_bump_count(int dtype, int k) {
  const int MB = dtype == IWQ_F16 ? 10 : 7;
  if (k == 0) return 0;
  const double ak = fabs((double)k);
  int fl = (int)floor(log2(ak));
  double ulp_below;  // spacing of D just below |k| (k > 0) / just above |k| (k < 0)
  if (k > 0) {
    const bool pow2 = ldexp(1.0, fl) == ak;
    ulp_below = ldexp(1.0, (pow2 ? fl - 1 : fl) - MB);
  } else {
    ulp_below = ldexp(1.0, fl - MB);
  }
  const double mid = (double)k - ulp_below * 0.5;
  int cnt = 0;
  for (int j = 1; j < (1 << MB); ++j) {
    const double l = (double)k + log2(1.0 - ldexp((double)j, -(MB + 1)));
    if (l > mid) ++cnt; else break;
  }
  return cnt;
}

static inline int build_quant_consts(const iwq_config_t* cfg, QuantConsts* c, int* kind, std::string* err) {
  auto fail = [&](int code, const char* msg) { *err = msg; return code; };
  memset(c, 0, sizeof(*c));
  const int dt = cfg->dtype;
  if (dt != IWQ_F16 && dt != IWQ_BF16) return fail(IWQ_ERR_INVALID, "dtype must be IWQ_F16 or IWQ_BF16");
  c->scheme = cfg->scheme;
  c->bits = cfg->bits;
  c->exp_bits = cfg->exp_bits;
  c->mant_bits = cfg->mant_bits;
  c->symmetric = cfg->symmetric ? 1 : 0;
  if (cfg->code_bits != 0 && cfg->code_bits != 8 && cfg->code_bits != 4) return fail(IWQ_ERR_INVALID, "code_bits must be 0, 8 or 4");
  c->code_nibbles = cfg->code_bits == 4 ? 1 : 0;
  if (c->code_nibbles) {
    const bool fits = cfg->scheme == IWQ_SCHEME_INT ? (cfg->bits >= 1 && cfg->bits <= 4)
                    : (cfg->scheme == IWQ_SCHEME_FP_CODEC || cfg->scheme == IWQ_SCHEME_FP_RNE) ? (1 + cfg->exp_bits + cfg->mant_bits <= 4) : false;
    if (!fits) return fail(IWQ_ERR_INVALID, "code_bits = 4 needs codes of at most 4 bits (INT bits <= 4, minifloat 1 + E + M <= 4)");
  }
  if (cfg->scheme == IWQ_SCHEME_INT) {
    const int b = cfg->bits;
    if (b < 1 || b > 15) return fail(IWQ_ERR_INVALID, "INT bits must be in [1, 15]");
    if (c->symmetric) {
      c->max_int = (float)((1 << (b - 1)) - 1);
      c->min_int = -(float)(1 << (b - 1));
    } else {
      c->max_int = (float)((1 << b) - 1);
      c->min_int = 0.0f;
    }
    c->eps = round_to_dtype(dt, 1e-5f);
    const float span = fmaxf(fabsf(c->min_int), fabsf(c->max_int));
    if (dt == IWQ_F16) {
      *kind = span <= 256.0f ? KIND_INT_FAST : KIND_INT_GENERIC;
    } else {
      *kind = span <= 64.0f ? KIND_INT_FAST : (span <= 256.0f ? KIND_INT_WIDE : KIND_INT_GENERIC);
    }
    if (getenv("IWQ_FORCE_GENERIC")) *kind = KIND_INT_GENERIC;
    // max_int is a value of the 16-bit dtype on the packed paths (<= 255): exact reciprocal-multiply division
    c->inv_div = (*kind == KIND_INT_GENERIC) ? 0.0f : 1.0f / c->max_int;
    return IWQ_OK;
  }
  if (cfg->scheme == IWQ_SCHEME_BFP) {
    const int b = cfg->bits;
    if (b < 2 || b > 15) return fail(IWQ_ERR_INVALID, "BFP w_bit must be in [2, 15]");
    if (cfg->group <= 0) return fail(IWQ_ERR_INVALID, "BFP needs a positive group size (quant_linear.py:650)");
    c->symmetric = 1;  // the range pass is the absolute maximum
    c->bfp_target = b - 1 < 11 ? b - 1 : 11;
    c->bfp_sd = 11 - c->bfp_target;
    c->bfp_round_mant = (dt == IWQ_BF16 && c->bfp_target > 8) ? 1 : 0;
    *kind = KIND_BFP;
    return IWQ_OK;
  }
  if (cfg->scheme != IWQ_SCHEME_FP_CODEC && cfg->scheme != IWQ_SCHEME_FP_RNE) return fail(IWQ_ERR_INVALID, "unknown scheme");
  const bool rne = cfg->scheme == IWQ_SCHEME_FP_RNE;
  const int E = cfg->exp_bits, M = cfg->mant_bits;
  if (E < 1 || E > 5 || M < 0 || M > 6 || E + M > 7) return fail(IWQ_ERR_INVALID, "minifloat needs 1<=E<=5, 0<=M<=6, E+M<=7");
  if (rne && M > 3) return fail(IWQ_ERR_UNSUPPORTED, "FP_RNE flavour is implemented (and verified exhaustively) for M <= 3");
  if (rne && dt != IWQ_F16) return fail(IWQ_ERR_UNSUPPORTED, "FP_RNE flavour is fp16 only (the reference casts to fp16: fp4_quantize_cpu.py:48-49)");
  if (rne) c->symmetric = 1;
  c->flavour_rne = rne ? 1 : 0;
  if (cfg->approx) {
    if (rne) return fail(IWQ_ERR_INVALID, "approximate decode applies to the FP_CODEC scheme only");
    if (cfg->approx < 0 || cfg->approx > 2) return fail(IWQ_ERR_INVALID, "approx must be 0, 1 or 2");
    if (cfg->group <= 0) return fail(IWQ_ERR_INVALID, "approximate 仅支持分组量化，w_group_size 必须 > 0");  // quant_linear.py:475-476
    c->symmetric = 1;  // quantize_weight_approximate always uses the absmax scale (:492-494)
    // FP4 with a single exponent bit never takes the double variant (:496-509)
    c->approx = (cfg->approx == 2 && E == 1) ? 1 : cfg->approx;
    c->hi_start = cfg->hi_align_start;
    c->hi_field = cfg->hi_align_exp_field;
    c->tail_pad = cfg->tail_pad_bits;
    if (c->tail_pad < -8 || c->tail_pad > 8) return fail(IWQ_ERR_INVALID, "tail_pad_bits out of range");
    if (c->approx == 2 && (c->hi_field < 0 || c->hi_field > 31)) return fail(IWQ_ERR_UNSUPPORTED, "double-approximate decode: hi_align_exp_field must be in [0, 31]");
  }
  const int MB = dt == IWQ_F16 ? 10 : 7, DB = dt == IWQ_F16 ? 15 : 127;
  const int max_field = dt == IWQ_F16 ? 30 : 254;
  const int bias = (1 << (E - 1)) - 1, emin = 1 - bias, emax = (1 << E) - 1 - bias;
  const double fmax = (2.0 - ldexp(1.0, -M)) * ldexp(1.0, emax);
  if (dt == IWQ_F16 && fmax > 65504.0) return fail(IWQ_ERR_FP16_RANGE, "value cannot be converted to type at::Half without overflow");
  c->fmax = (float)fmax;
  c->inv_div = 1.0f / c->fmax;  // fmax = (2 - 2^-M) * 2^emax has <= 7 significant bits: a value of either dtype
  c->eps = round_to_dtype(dt, rne ? 1e-8f : 1e-5f);
  int elo = emin - M - 3;
  if (elo < 1 - DB) elo = 1 - DB;
  if (elo > emin - M - 2) return fail(IWQ_ERR_UNSUPPORTED, "minifloat grid too fine for the storage dtype");
  const int n = emax - elo + 1;
  if (n > 40) return fail(IWQ_ERR_UNSUPPORTED, "minifloat exponent range too wide");
  const uint16_t inf_bits = dt == IWQ_F16 ? 0x7C00 : 0x7F80;
  for (int i = 0; i < n; ++i) {
    const int e = elo + i;
    const int ge = e > emin ? e : emin;
    const int mfield = ge - M + MB + DB;
    if (mfield < 1 || mfield > max_field) return fail(IWQ_ERR_UNSUPPORTED, "minifloat magic constant not representable in the storage dtype");
    FpEntry en;
    en.magic = (uint16_t)((mfield << MB) | (1 << (MB - 1)));
    if (rne) {
      en.bound = inf_bits;
    } else if (e >= emin) {
      en.bound = (uint16_t)(((e + DB) << MB) | (((1 << M) - 1) << (MB - M)));
    } else if (M == 0) {
      en.bound = 0;
    } else {
      en.bound = (uint16_t)(((emin - 1 + DB) << MB) | (((1 << (M - 1)) - 1) << (MB - (M - 1))));
    }
    en.bump_thr = inf_bits;
    en.bump_val = 0;
    const int k = e + 1;
    if (!rne && k >= emin && k <= emax) {
      const int cnt = log2_bump_count(dt, k);
      if (cnt > 0) {
        en.bump_thr = (uint16_t)(((e + DB) << MB) | ((1 << MB) - cnt));
        en.bump_val = (uint16_t)((k + DB) << MB);
      }
    }
    c->tbl[i] = en;
  }
  c->tbl_base = elo + DB;
  c->tbl_n = n;
  c->tiny_bits = (uint16_t)((elo + DB) << MB);
  {
    int llo = emin - M - 1;  // 2^(emin-M-1) is half the smallest step: it and everything below rounds to 0
    if (llo + DB < 1) llo = 1 - DB;
    c->lut_lo_bits = (uint16_t)((llo + DB) << MB);
    // fp16: the table starts at bit pattern 0, which saves the kernel a clamp and a subtract per pair
    // (FpTable::kFromZero); bf16 has 8 exponent bits, a table from 0 would be far too large
    if (dt == IWQ_F16) c->lut_lo_bits = 0;
  }
  c->fmax_bits = (uint16_t)(((emax + DB) << MB) | (((1 << M) - 1) << (MB - M)));
  c->lut_shift = (uint16_t)(MB - M - 1);
  c->lut_n = (((uint32_t)c->fmax_bits - (uint32_t)c->lut_lo_bits) >> c->lut_shift) + 1u;
  *kind = KIND_FP;
  return IWQ_OK;
}


// Re-shape the step table of a minifloat launch for `1 << rep_shift` interleaved copies in shared memory (seg4 kernel,
// FpTable<.., REPL = true>): the table starts at the first pattern that does not quantise to 0 (also for fp16, whose
// single-copy table starts at pattern 0) and the interval width is capped so that one entry's stride (8 << rep_shift bytes)
// is a multiple of the interval width in patterns -- the lookup address then is a single multiply-add of the masked
// pattern.  Narrower intervals still change value at most once each.  rep_shift = 0 restores nothing: call it on fresh consts.
static inline void set_lut_replication(int dt, QuantConsts* c, int rep_shift) {
  if (rep_shift <= 0) return;
  const int MB = dt == IWQ_F16 ? 10 : 7, DB = dt == IWQ_F16 ? 15 : 127;
  const int E = c->exp_bits, M = c->mant_bits;
  const int bias = (1 << (E - 1)) - 1, emin = 1 - bias;
  int llo = emin - M - 1;
  if (llo + DB < 1) llo = 1 - DB;
  c->lut_lo_bits = (uint16_t)((llo + DB) << MB);
  int sh = MB - M - 1;
  if (sh > 3 + rep_shift) sh = 3 + rep_shift;
  c->lut_shift = (uint16_t)sh;
  c->lut_n = (((uint32_t)c->fmax_bits - (uint32_t)c->lut_lo_bits) >> c->lut_shift) + 1u;
  c->lut_rep_shift = (uint32_t)rep_shift;
}

// _fp_decode_aligned (quant_linear.py:237-285; align_subnorm_exp_as_one=True, limit_align_exp_to_field=True,
// decode_dtype = weight dtype, as the driver calls it) for the code of a POSITIVE grid value: the single
// "approximate" mode only changes which value a code decodes to, so it is folded into the step table.
template <int DT>
static inline uint16_t approx_decode_single(uint16_t vbits, const QuantConsts& c) {
  const int E = c.exp_bits, M = c.mant_bits, bias = (1 << (E - 1)) - 1;
  const uint32_t code = fp_code_of<DT>(vbits, E, M);
  if (code == 0) return 0;
  const int ef = (int)((code >> M) & ((1u << E) - 1u)), mf = (int)(code & ((1u << M) - 1u));
  const int align_exp = ef == 0 ? 1 : ef;
  const int mant_full = ((ef == 0 ? 0 : 1) << M) | mf;
  int mant_padded;
  if (c.tail_pad >= 0) {
    mant_padded = mant_full << c.tail_pad;
  } else {
    const int sh = -c.tail_pad;
    mant_padded = (mant_full + (1 << (sh - 1))) >> sh;
  }
  float v;
  if (align_exp >= c.hi_start && align_exp <= c.hi_field) {
    const int shift = c.hi_field - align_exp;
    const int aligned = shift > 0 ? (mant_padded + (1 << (shift - 1))) >> shift : mant_padded;
    v = (float)aligned / ldexpf(1.0f, M + c.tail_pad) * ldexpf(1.0f, c.hi_field - bias);       // fp32  (:281)
  } else {
    float a = rD<DT>(rD<DT>((float)mant_full) / ldexpf(1.0f, M));                               // D ops (:268-269)
    v = rD<DT>(a * rD<DT>(ldexpf(1.0f, ef == 0 ? 1 - bias : ef - bias)));
  }
  return DType<DT>::from_f32(v);
}

// DIRECT table: one value per bit pattern from the first pattern that does not quantise to 0 up to fmax (FpTable<.., LUT_DIRECT>).
static inline void set_lut_direct(int dt, QuantConsts* c) {
  const int MB = dt == IWQ_F16 ? 10 : 7, DB = dt == IWQ_F16 ? 15 : 127;
  const int E = c->exp_bits, M = c->mant_bits;
  const int bias = (1 << (E - 1)) - 1, emin = 1 - bias;
  int llo = emin - M - 1;
  if (llo + DB < 1) llo = 1 - DB;
  c->lut_lo_bits = (uint16_t)((llo + DB) << MB);
  c->lut_shift = 0;
  c->lut_n = (uint32_t)c->fmax_bits - (uint32_t)c->lut_lo_bits + 1u;
  c->lut_rep_shift = 0;
  c->lut_direct = 1;
}
// values of the direct table: the table-walk quantiser grid() on every pattern (then the single approximate decode, if any)
template <int DT>
static inline void build_direct_table_t(const QuantConsts& c, uint16_t* out) {
  using T2 = typename DType<DT>::T2;
  FpQ q;
  q.S = 1.0f; q.r = 1.0f; q.mid = 0.f; q.mid_h = 0.f; q.bad = false; q.slow_div = false;
  FpTable<DT> f;
  f.setup(q, c);
  for (uint32_t i = 0; i < c.lut_n; ++i) {
    const uint32_t pat = (uint32_t)c.lut_lo_bits + i;
    uint16_t v = (uint16_t)(as_u32(f.grid(from_u32<T2>(pat | (pat << 16)), 0u, c.tbl)) & 0xFFFFu);
    if (c.approx == 1) v = approx_decode_single<DT>(v, c);
    out[i] = v;
  }
}
static inline void build_direct_table(int dtype, const QuantConsts& c, uint16_t* out) {
  if (dtype == IWQ_F16) build_direct_table_t<IWQ_F16>(c, out); else build_direct_table_t<IWQ_BF16>(c, out);
}

// Step table (see FpStep): run the table-walk quantiser grid() on every bit pattern of every interval and
// record the single value change.  Returns false if some interval changes value more than once (would mean
// the format violates the assumption; none of the supported ones does).
template <int DT>
static inline bool build_step_table_t(const QuantConsts& c, FpStep* out) {
  using T2 = typename DType<DT>::T2;
  FpQ q;
  q.S = 1.0f; q.r = 1.0f; q.mid = 0.f; q.mid_h = 0.f; q.bad = false; q.slow_div = false;
  FpTable<DT> f;
  f.setup(q, c);
  const uint32_t W = 1u << c.lut_shift;
  const uint16_t inf_bits = (uint16_t)DType<DT>::kInfBits;
  for (uint32_t i = 0; i < c.lut_n; ++i) {
    const uint32_t p0 = (uint32_t)c.lut_lo_bits + i * W;
    auto g = [&](uint32_t pat) { return (uint16_t)(as_u32(f.grid(from_u32<T2>(pat | (pat << 16)), 0u, c.tbl)) & 0xFFFFu); };
    FpStep e;
    e.below = g(p0);
    e.above = e.below;
    e.thr = inf_bits;
    e.pad = 0;
    bool changed = false;
    for (uint32_t p = p0 + 1; p < p0 + W && p <= c.fmax_bits; ++p) {
      const uint16_t v = g(p);
      if (!changed) {
        if (v != e.below) {
          changed = true;
          e.thr = (uint16_t)p;
          e.above = v;
        }
      } else if (v != e.above) {
        return false;
      }
    }
    if (c.approx == 1) {
      e.below = approx_decode_single<DT>(e.below, c);
      e.above = approx_decode_single<DT>(e.above, c);
    }
    out[i] = e;
  }
  return true;
}
static inline bool build_step_table(int dtype, const QuantConsts& c, FpStep* out) {
  return dtype == IWQ_F16 ? build_step_table_t<IWQ_F16>(c, out) : build_step_table_t<IWQ_BF16>(c, out);
}

// Tables of the double-approximate decode (see double_approx_info in iwq_math.cuh): the step table with the
// INFO word of each grid value in place of the value, and val[target << 8 | sign << 7 | code magnitude].
template <int DT>
static inline bool build_double_tables_t(const QuantConsts& c, FpStep* steps, uint16_t* val) {
  if (!build_step_table_t<DT>(c, steps)) return false;  // approx == 2 leaves the values themselves untouched
  const int E = c.exp_bits, M = c.mant_bits;
  for (uint32_t i = 0; i < c.lut_n; ++i) {
    steps[i].below = double_approx_info(fp_code_of<DT>(steps[i].below, E, M), c);
    steps[i].above = double_approx_info(fp_code_of<DT>(steps[i].above, E, M), c);
  }
  for (int t = 0; t < 32; ++t)
    for (uint32_t k = 0; k < 256; ++k) {
      const uint32_t mag = k & 0x7Fu, sign = k >> 7;
      uint16_t v = 0;
      if (mag < (1u << (E + M))) v = double_approx_elem<DT>(mag | (sign << (E + M)), t, c);
      val[(t << 8) | k] = v;
    }
  return true;
}
static inline bool build_double_tables(int dtype, const QuantConsts& c, FpStep* steps, uint16_t* val) {
  return dtype == IWQ_F16 ? build_double_tables_t<IWQ_F16>(c, steps, val) : build_double_tables_t<IWQ_BF16>(c, steps, val);
}

}  // namespace iwq
