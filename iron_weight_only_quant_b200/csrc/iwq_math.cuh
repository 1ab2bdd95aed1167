// iwq_math.cuh -- per-segment parameter computation and per-element transforms of the fake-quant
// path, written so that every intermediate is rounded exactly where the reference's eager ATen ops
// round it ("fp32 compute, round to the storage dtype after every op").
//
// Everything here is __host__ __device__: the CUDA kernels (iwq_kernels.cuh) are the product; the
// host instantiation exists ONLY so that tests/host/emulate.cu can run the very same functions on
// the CPU build box (no GPU there) against the golden vectors before GPU time is spent.  The Python
// package never calls the host instantiation.
//
// Reference citations are file:line in LiuTielong/Iron_weight_only_quant.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

#include "../../include/iwq_b200.h"

#define IWQ_HD __host__ __device__ __forceinline__

namespace iwq {

// ------------------------------------------------------------------------------------------------
// storage dtype traits
// ------------------------------------------------------------------------------------------------
template <int DT>
struct DType;

IWQ_HD uint32_t f32_bits(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
IWQ_HD float bits_f32(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

template <>
struct DType<IWQ_F16> {
  using T = __half;
  using T2 = __half2;
  static constexpr int kMantBits = 10;
  static constexpr int kExpBias = 15;
  static constexpr uint32_t kInfBits = 0x7C00u;
  static IWQ_HD float to_f32(uint16_t b) { return __half2float(__ushort_as_half(b)); }
  static IWQ_HD uint16_t from_f32(float f) { return __half_as_ushort(__float2half_rn(f)); }
  static IWQ_HD T2 from_f32x2(float lo, float hi) { return __floats2half2_rn(lo, hi); }
  static IWQ_HD float2 to_f32x2(T2 v) { return __half22float2(v); }
  static IWQ_HD T2 bcast(uint16_t b) { return __half2half2(__ushort_as_half(b)); }
};

template <>
struct DType<IWQ_BF16> {
  using T = __nv_bfloat16;
  using T2 = __nv_bfloat162;
  static constexpr int kMantBits = 7;
  static constexpr int kExpBias = 127;
  static constexpr uint32_t kInfBits = 0x7F80u;
  static IWQ_HD float to_f32(uint16_t b) { return bits_f32((uint32_t)b << 16); }
  static IWQ_HD uint16_t from_f32(float f) { return __bfloat16_as_ushort(__float2bfloat16_rn(f)); }
  static IWQ_HD T2 from_f32x2(float lo, float hi) { return __floats2bfloat162_rn(lo, hi); }
  static IWQ_HD float2 to_f32x2(T2 v) {
    // exact and one ALU op per element: a bf16 is the top half of an fp32
    uint32_t u = *reinterpret_cast<uint32_t*>(&v);
    return make_float2(bits_f32(u << 16), bits_f32(u & 0xFFFF0000u));
  }
  static IWQ_HD T2 bcast(uint16_t b) { return __bfloat162bfloat162(__ushort_as_bfloat16(b)); }
};

template <class T2>
IWQ_HD uint32_t as_u32(T2 v) {
  return *reinterpret_cast<uint32_t*>(&v);
}
template <class T2>
IWQ_HD T2 from_u32(uint32_t u) {
  return *reinterpret_cast<T2*>(&u);
}

// round an fp32 value to the storage dtype and come back (the reference's per-op rounding)
template <int DT>
IWQ_HD float rD(float x) {
  return DType<DT>::to_f32(DType<DT>::from_f32(x));
}

// torch.clamp semantics: NaN propagates (one FMNMX.NAN each on the device)
IWQ_HD float clamp_min_nan(float x, float lo) {
#ifdef __CUDA_ARCH__
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo));
  return r;
#else
  return (x != x) ? x : fmaxf(x, lo);
#endif
}
IWQ_HD float clamp_nan(float x, float lo, float hi) {
#ifdef __CUDA_ARCH__
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(hi));
  return r;
#else
  return (x != x) ? x : fminf(fmaxf(x, lo), hi);
#endif
}

IWQ_HD float ieee_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
IWQ_HD float ieee_rcp(float a) {
#ifdef __CUDA_ARCH__
  return __frcp_rn(a);
#else
  return 1.0f / a;
#endif
}
IWQ_HD float fma_rn(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}

// half2 / bfloat162 fused multiply-add with a zero addend.  On the device this is one HFMA2; the
// host emulation multiplies in fp32 (exact: 11x11 / 8x8 bit significands) and rounds once.
template <int DT>
IWQ_HD typename DType<DT>::T2 mul_add_zero(typename DType<DT>::T2 a, typename DType<DT>::T2 b) {
  using T2 = typename DType<DT>::T2;
#ifdef __CUDA_ARCH__
  return __hfma2(a, b, from_u32<T2>(0u));
#else
  float2 x = DType<DT>::to_f32x2(a), y = DType<DT>::to_f32x2(b);
  return DType<DT>::from_f32x2(x.x * y.x + 0.0f, x.y * y.y + 0.0f);
#endif
}

// quotient w/s for 16-bit w, s, such that rounding it to D equals rD(fl32(w / s)).
// r = fl32(1/s);  t = w*r;  e = fma(-s, t, w);  q = fma(r, e, t).
// Proven for every (w, s) pair of both dtypes by tests/host/div_check.c (CPU, exhaustive) and by
// iwq_selftest_division on the device.  Three FMA-pipe ops instead of an IEEE division (8-10).
IWQ_HD float fast_quot(float w, float s, float r) {
  float t = w * r;
  float e = fma_rn(-s, t, w);
  return fma_rn(r, e, t);
}

template <int DT>
IWQ_HD typename DType<DT>::T2 fast_quot2(typename DType<DT>::T2 w2, float s, float r) {
  float2 w = DType<DT>::to_f32x2(w2);
#ifdef __CUDA_ARCH__
  // packed fp32x2 FMA (sm_100a FFMA2): one issue slot for both elements
  const float2 rr = make_float2(r, r), ns = make_float2(-s, -s);
  float2 t = __fmul2_rn(w, rr);
  float2 e = __ffma2_rn(ns, t, w);
  float2 q = __ffma2_rn(rr, e, t);
  return DType<DT>::from_f32x2(q.x, q.y);
#else
  return DType<DT>::from_f32x2(fast_quot(w.x, s, r), fast_quot(w.y, s, r));
#endif
}

template <int DT>
IWQ_HD typename DType<DT>::T2 ieee_quot2(typename DType<DT>::T2 w2, float s) {
  float2 w = DType<DT>::to_f32x2(w2);
  return DType<DT>::from_f32x2(ieee_div(w.x, s), ieee_div(w.y, s));
}

// Pair forms: the two halves of a packed register belong to DIFFERENT segments (quant_dim = 1: adjacent
// columns), so scale and reciprocal are per half.
template <int DT>
IWQ_HD typename DType<DT>::T2 fast_quot2(typename DType<DT>::T2 w2, float2 s, float2 r) {
  float2 w = DType<DT>::to_f32x2(w2);
#ifdef __CUDA_ARCH__
  const float2 ns = make_float2(-s.x, -s.y);
  float2 t = __fmul2_rn(w, r);
  float2 e = __ffma2_rn(ns, t, w);
  float2 q = __ffma2_rn(r, e, t);
  return DType<DT>::from_f32x2(q.x, q.y);
#else
  return DType<DT>::from_f32x2(fast_quot(w.x, s.x, r.x), fast_quot(w.y, s.y, r.y));
#endif
}
template <int DT>
IWQ_HD typename DType<DT>::T2 ieee_quot2(typename DType<DT>::T2 w2, float2 s) {
  float2 w = DType<DT>::to_f32x2(w2);
  return DType<DT>::from_f32x2(ieee_div(w.x, s.x), ieee_div(w.y, s.y));
}

template <bool PAIR>
struct PerHalf {  // a per-segment fp32 constant: one value, or one per half of a packed pair
  using F = float;
  static IWQ_HD float make(float a, float) { return a; }
  static IWQ_HD float lo(float v) { return v; }
  static IWQ_HD float hi(float v) { return v; }
};
template <>
struct PerHalf<true> {
  using F = float2;
  static IWQ_HD float2 make(float a, float b) { return make_float2(a, b); }
  static IWQ_HD float lo(float2 v) { return v.x; }
  static IWQ_HD float hi(float2 v) { return v.y; }
};

// ------------------------------------------------------------------------------------------------
// configuration constants shared by all segments of a launch (built on the host: iwq_host.cu)
// ------------------------------------------------------------------------------------------------
struct FpEntry {      // one per binade of the clamped magnitude |n|
  uint16_t magic;     // 1.5 * 2^(grid_exp + mantissa_bits(D)): (a + magic) - magic == RNE onto the grid
  uint16_t bound;     // largest value of the binade the codec may return (no carry, quirk Q2); +inf for RNE flavour
  uint16_t bump_thr;  // a >= bump_thr: half-precision log2 rounds up to the next integer (quirk Q3) ...
  uint16_t bump_val;  // ... and the result is 2^(e+1)
};

// Step table of a minifloat format: the quantiser is a monotone step function of the clamped magnitude, and
// inside an interval of W consecutive bit patterns it changes value at most once (a round-to-even tie sits on
// an interval's first pattern; a log2-bump threshold sits inside a binade's last interval).  Entry i covers
// patterns [lut_lo_bits + i*W, +W):  value = pattern >= thr ? above : below.  Built on the host by running
// grid() over every pattern (iwq_host.cu: build_step_table), <= 4 KB, copied to shared memory by each CTA.
struct FpStep {
  uint16_t thr, below, above, pad;
};

struct QuantConsts {
  int32_t scheme, bits, exp_bits, mant_bits, symmetric, flavour_rne;
  float min_int, max_int;  // INT: clamp bounds (quant_funcs.py:19-20, :26-27)
  float inv_div;           // fl32(1 / max_int) resp. fl32(1 / fmax), or 0: the divisor is not a 16-bit value -> IEEE division
  float eps;               // rD(1e-5) (quant_funcs.py:21,25) or rD(1e-8) (fp4_quantize_cpu.py:68)
  float fmax;              // FP: largest minifloat magnitude (quant_linear.py:741 / fp4_quantize_cpu.py:64)
  uint16_t fmax_bits;      // fmax as D bits
  uint16_t tiny_bits;      // D bits of 2^(emin - M - 3): magnitudes are clamped up to it before the table lookup
  int32_t tbl_base;        // biased D exponent of table entry 0
  int32_t tbl_n;
  int32_t approx, hi_start, hi_field, tail_pad;  // aligned ("approximate") decode, quant_linear.py:237-363
  int32_t bfp_target;      // BFP: mantissa bits kept incl. the leading one = min(w_bit-1, 11)   quant_linear.py:673
  int32_t bfp_sd;          // BFP: 11 - bfp_target (rounding right shift)                          quant_linear.py:674
  int32_t bfp_round_mant;  // BFP on bf16 with more than 8 mantissa bits: `mant_rounded.to(bf16)` rounds (:707)
  int32_t code_nibbles;    // 1: emitted codes are packed two per byte (iwq_config.code_bits == 4)
  uint16_t lut_lo_bits;    // D bits of 2^(emin-M-1): everything at or below it quantises to 0 (exact tie -> even)
  uint16_t lut_shift;      // log2 of the interval width W = 2^(mantissa_bits(D) - M - 1) bit patterns (half a grid step)
  uint32_t lut_n;          // number of FpStep intervals covering bit patterns lut_lo_bits .. fmax_bits
  uint32_t lut_rep_shift;  // log2 of the number of interleaved copies of the step table in shared memory (seg4 kernel: 16
                           // copies, entry i of copy r at [i][r], so that the lanes of a half-warp hit 16 distinct bank pairs
                           // whatever their indices -- no bank conflicts); 0: a single copy
  uint32_t lut_direct;     // 1: the table is DIRECT -- one 16-bit value per bit pattern lut_lo_bits .. fmax_bits (lut_n entries):
                           // one 2-byte load per element, no threshold compare / select (FpTable<.., LUT_DIRECT>)
  FpEntry tbl[40];
};

// ------------------------------------------------------------------------------------------------
// INT min-max: per-segment parameters        (quant_funcs.py:16-29, quant_linear.py:909-922)
// ------------------------------------------------------------------------------------------------
struct IntQ {
  float s;         // scale, a D value
  float r;         // fl32(1/s)
  float z;         // zero point (integer valued), 0 for symmetric
  float lo, hi;    // min_int - z, max_int - z: clamp(rint(x)+z, min, max) - z == clamp(rint(x), lo, hi)
  bool bad;        // !(0 < s < inf): the segment yields NaN (quant_funcs.py:40 assert condition)
  bool slow_div;   // bf16 only: |w|/s can overflow fp32, use IEEE division for this segment
};

// x / c for a launch-constant divisor c (max_int or fmax).  When c is itself a value of the 16-bit dtype
// the reciprocal+FMA quotient is exact after rounding to D (same proof as fast_quot); otherwise IEEE.
IWQ_HD float div_const(float x, float c, float inv_c) {
  // (an infinite or NaN numerator takes the IEEE path so that inf / c stays inf)
  return (inv_c != 0.0f && fabsf(x) < INFINITY) ? fast_quot(x, c, inv_c) : ieee_div(x, c);
}

// mx = max (asym) or absmax (sym); mn = min (asym), ignored for sym.  D values held in fp32.
// `sym` duplicates c.symmetric so that kernels can pass their template constant (the branch folds away).
template <int DT>
IWQ_HD IntQ int_segment_params(float mx, float mn, const QuantConsts& c, bool sym) {
  IntQ q;
  float d;
  if (sym) {
    d = rD<DT>(clamp_min_nan(mx, c.eps));                 // absmax.clamp(min=1e-5)            :25
  } else {
    d = rD<DT>(mx - mn);                                  // max_val - min_val                  :21
    d = rD<DT>(clamp_min_nan(d, c.eps));                  // .clamp(min=1e-5)                   :21
  }
  q.s = rD<DT>(div_const(d, c.max_int, c.inv_div));       // / max_int                          :21 / :28
  q.r = ieee_rcp(q.s);
  // "suspect": the segment MAY yield NaN (quant_funcs.py:40 asserts on NaN in the result).  s NaN / inf always does;
  // s == 0 (a range below ~max_int * 2^-25 at >= 9 bits in fp16) does only where an element is 0 (0 / 0) -- a suspect
  // segment takes the IEEE-division form (w / 0 = inf -> clamps like the reference's) and its outputs are inspected
  q.bad = !(q.s > 0.0f && q.s < INFINITY);
  q.slow_div = q.bad;
  if (DT == IWQ_BF16) {
    float big = fmaxf(fabsf(mx), sym ? 0.0f : fabsf(mn));
    q.slow_div = q.bad || !(big * q.r < 1e37f);
  }
  if (sym) {
    q.z = 0.0f;                                           // zeros = 0                          :29
  } else {
    float zq = rD<DT>((q.slow_div || c.inv_div == 0.0f) ? ieee_div(mn, q.s) : fast_quot(mn, q.s, q.r));  // min_val / scales  :22
    // (-round(.)).clamp_(min_int,max_int) :22.  A zero keeps its sign through torch.clamp (min/s in [0, 0.5] gives -0.0,
    // the reference stores that bit pattern), a negative value comes back as the bound +0.0
    const float rz = -rintf(zq);
    q.z = (rz == 0.0f) ? rz : clamp_nan(rz, c.min_int, c.max_int);
  }
  q.lo = c.min_int - q.z;
  q.hi = c.max_int - q.z;
  return q;
}

// Packed element stage for the fast INT path.
//   x  = rD(w / s)                                  one F2FP pack after the fp32 quotient
//   r  = rint(x)                                    (x + K) - K in D arithmetic, K = 1.5 * 2^mantissa_bits
//   d  = clamp(r, lo, hi)          == clamp(r + z, min_int, max_int) - z   (all values are small integers)
//   out= fma(d, s, +0)             == rD(d * s), and turns -0 into +0 exactly like "(-0.0) + zeros" does
// Valid when |lo|,|hi| <= 512 (fp16) / 64 (bf16); kWide extends bf16 to 256 via the |x|+128 trick.
template <int DT, bool kWide, bool PAIR = false>
struct IntFast {
  using D = DType<DT>;
  using T2 = typename D::T2;
  using F = typename PerHalf<PAIR>::F;
  T2 lo2, hi2, s2;
  F s, r;
  bool slow_div;

  IWQ_HD void setup(const IntQ& q) { setup2(q, q); }
  // low half <- qa, high half <- qb (equal for the row-grouped kernels)
  IWQ_HD void setup2(const IntQ& qa, const IntQ& qb) {
    lo2 = D::from_f32x2(qa.lo, qb.lo);
    hi2 = D::from_f32x2(qa.hi, qb.hi);
    s2 = D::from_f32x2(qa.s, qb.s);
    s = PerHalf<PAIR>::make(qa.s, qb.s);
    r = PerHalf<PAIR>::make(qa.r, qb.r);
    slow_div = qa.slow_div || qb.slow_div;
  }

  IWQ_HD T2 rint2(T2 x) const {
    if (DT == IWQ_F16) {
      const T2 K = D::bcast(0x6600);  // 1536.0
      return __hsub2(__hadd2(x, K), K);
    } else if (!kWide) {
      const T2 K = D::bcast(0x4340);  // 192.0
      return __hsub2(__hadd2(x, K), K);
    } else {
      // bf16, |range| up to 256: |x| < 128 -> (|x| + 128) - 128 rounds to integer (spacing 1 in
      // [128,256)); |x| >= 128 is already an integer.  Sign restored by OR (rint is odd).
      const T2 K = D::bcast(0x4300);  // 128.0
      uint32_t xb = as_u32(x);
      uint32_t ab = xb & 0x7FFF7FFFu;
      T2 a = from_u32<T2>(ab);
      T2 ra = __hsub2(__hadd2(a, K), K);
      uint32_t big = __hge2_mask(a, K);  // also false for NaN -> ra (NaN) is kept
      uint32_t rb = (as_u32(ra) & ~big) | (ab & big);
      return from_u32<T2>(rb | (xb & 0x80008000u));
    }
  }

  // returns the dequantised pair; *d2 receives clamp(rint(x), lo, hi) (for code emission).
  // SLOW: 1 / 0 fixes the division form at compile time (callers that hoisted the segment-uniform branch), -1 decides here
  template <int SLOW = -1>
  IWQ_HD uint32_t apply(uint32_t w2, T2* d2) const {
    T2 w = from_u32<T2>(w2);
    const bool sl = SLOW < 0 ? slow_div : (SLOW != 0);
    T2 x = sl ? ieee_quot2<DT>(w, s) : fast_quot2<DT>(w, s, r);
    T2 q = rint2(x);
    q = __hmin2_nan(__hmax2_nan(q, lo2), hi2);
    *d2 = q;
    return as_u32(mul_add_zero<DT>(q, s2));
  }
};

// Reference-order scalar element stage: any bit width, any magnitude (quant_funcs.py:36-38).
template <int DT>
IWQ_HD uint16_t int_generic_elem(uint16_t wb, const IntQ& q, const QuantConsts& c, float* code) {
  float w = DType<DT>::to_f32(wb);
  float x = rD<DT>(ieee_div(w, q.s));        // tensor / scales
  x = rintf(x);                              // torch.round
  x = rD<DT>(x + q.z);                       // + zeros
  x = clamp_nan(x, c.min_int, c.max_int);    // torch.clamp
  *code = x;
  float o = rD<DT>(x - q.z);                 // - zeros
  o = rD<DT>(o * q.s);                       // * scales
  return DType<DT>::from_f32(o);
}

// ------------------------------------------------------------------------------------------------
// minifloat: per-segment parameters
//   codec flavour: quant_linear.py:744-755 ; RNE flavour: fp4_quantize_cpu.py:67-69
// ------------------------------------------------------------------------------------------------
struct FpQ {
  float S;        // scale (D value)
  float r;        // fl32(1/S)
  float mid;      // asym mid point (D value), 0 for sym
  float mid_h;    // rD(fp16(mid)): what is added back (quirk Q1: `self.zeros` is a .half() copy)
  bool bad;       // !(0 < S < inf)
  bool slow_div;
};

template <int DT>
IWQ_HD FpQ fp_segment_params(float mx, float mn, const QuantConsts& c) {
  FpQ q;
  q.mid = 0.0f;
  q.mid_h = 0.0f;
  if (c.flavour_rne) {
    float d = rD<DT>(clamp_min_nan(mx, c.eps));                    // absmax.clamp(min=1e-8)   fp4_quantize_cpu.py:68
    q.S = rD<DT>(div_const(d, c.fmax, c.inv_div));                // / max_float              :69
  } else if (c.symmetric) {
    float d = rD<DT>(clamp_min_nan(mx, c.eps));                    // absmax.clamp(min=1e-5)   quant_linear.py:744
    q.S = rD<DT>(clamp_min_nan(rD<DT>(div_const(d, c.fmax, c.inv_div)), c.eps));  // (/ fmax).clamp(1e-5)  :745
  } else {
    q.mid = rD<DT>(rD<DT>(mx + mn) * 0.5f);                        // (max + min) * 0.5        :751
    float span = rD<DT>(rD<DT>(mx - mn) * 0.5f);                   // (max - min) * 0.5        :752
    span = rD<DT>(clamp_min_nan(span, c.eps));                     // .clamp(min=1e-5)         :752
    q.S = rD<DT>(clamp_min_nan(rD<DT>(div_const(span, c.fmax, c.inv_div)), c.eps));  //          :753
    q.mid_h = rD<DT>(DType<IWQ_F16>::to_f32(DType<IWQ_F16>::from_f32(q.mid)));  // zeros.half().to(D) :765,:776
  }
  q.r = ieee_rcp(q.S);
  q.bad = !(q.S > 0.0f && q.S < INFINITY);
  float big = fmaxf(fabsf(mx), fabsf(mn)) * 2.0f;
  q.slow_div = q.bad || !(big * q.r < 1e37f);
  return q;
}

// Packed element stage for both minifloat flavours.
//   n   = clamp(rD([w - mid] / S), +-fmax)                                         quant_linear.py:747,755
//   a   = max(|n|, tiny)            tiny = 2^(emin-M-3) quantises to 0 like everything below it
//   per-binade constants from the table (indexed by the D exponent field of a)
//   v   = min((a + magic) - magic, bound)     RNE onto the binade's grid, mantissa overflow does not
//                                             carry (quirk Q2, :148-149); subnormal grid below emin (:152-154)
//   v   = a >= bump_thr ? bump_val : v        half-precision log2 rounding (quirk Q3, :139)
//   sign: codec -> (n < 0), so a negative input that flushes to zero gives -0.0 (quirk Q4, :130,:234)
//         RNE   -> sign bit of n (rint(-0.0/sc)*sc = -0.0, fp4_quantize_cpu.py:42-43)
//   out = rD(v * S) [+ mid_h]                                                      :774-777
enum { LUT_STEP = 0, LUT_REPL = 1, LUT_DIRECT = 2 };

template <int DT, bool PAIR = false, int LUT = LUT_STEP>
struct FpTable {
  static constexpr bool REPL = LUT == LUT_REPL;
  using D = DType<DT>;
  using T2 = typename D::T2;
  using F = typename PerHalf<PAIR>::F;
  T2 S2, mid2, midh2, fmax2, tiny2, lutlo2;
  F S, r;
  bool asym, rne, slow_div;
  int tbl_off;  // entry index = exponent_field - tbl_base
  uint32_t lo2, sh;          // step table: packed low pattern, log2 interval width
  const FpStep* steps;       // host: pointer; device: also as a 32-bit shared-memory address
  uint32_t steps_saddr;
  // REPL (replicated table): byte address of entry i for this lane = rep_base + i * (8 << rep_shift); the interval index is
  // never materialised -- the masked pattern is scaled straight into the address by one multiply-add per half
  uint32_t rep_base, rep_mask_hi, rep_mul_lo, rep_mul_hi;

  IWQ_HD void setup(const FpQ& q, const QuantConsts& c, const FpStep* step_tbl = nullptr) { setup2(q, q, c, step_tbl); }
  IWQ_HD void setup2(const FpQ& q, const FpQ& qb, const QuantConsts& c, const FpStep* step_tbl = nullptr) {
    steps = step_tbl;
    steps_saddr = 0;
#ifdef __CUDA_ARCH__
    if (step_tbl) steps_saddr = (uint32_t)__cvta_generic_to_shared(step_tbl);
#endif
    lo2 = (uint32_t)c.lut_lo_bits | ((uint32_t)c.lut_lo_bits << 16);
    sh = c.lut_shift;
    S2 = D::from_f32x2(q.S, qb.S);
    mid2 = D::from_f32x2(q.mid, qb.mid);
    midh2 = D::from_f32x2(q.mid_h, qb.mid_h);
    fmax2 = D::bcast(c.fmax_bits);
    tiny2 = D::bcast(c.tiny_bits);
    lutlo2 = D::bcast(c.lut_lo_bits);
    S = PerHalf<PAIR>::make(q.S, qb.S);
    r = PerHalf<PAIR>::make(q.r, qb.r);
    asym = !c.symmetric && !c.flavour_rne;
    rne = c.flavour_rne != 0;
    slow_div = q.slow_div || qb.slow_div;
    tbl_off = c.tbl_base;
    rep_base = rep_mask_hi = rep_mul_lo = rep_mul_hi = 0;
#ifdef __CUDA_ARCH__
    if (LUT == LUT_DIRECT) rep_base = steps_saddr - 2u * (uint32_t)c.lut_lo_bits;  // byte address of pattern p: rep_base + 2 p
    if (REPL) {
      // stride = 8 << rep_shift bytes per entry, k = log2(stride) - sh >= 0 (the host caps the interval width accordingly)
      const uint32_t k = 3u + c.lut_rep_shift - c.lut_shift;
      rep_mask_hi = (0xFFFFu & ~((1u << c.lut_shift) - 1u)) << 16;
      rep_mul_lo = 1u << k;           // (pattern_lo & mask) * 2^k          = index * stride
      rep_mul_hi = 16u - k;           // ((pattern_hi & mask) << 16) >> (16 - k) = index * stride
      rep_base = steps_saddr + (threadIdx.x & ((1u << c.lut_rep_shift) - 1u)) * 8u - (((uint32_t)c.lut_lo_bits >> c.lut_shift) << (3u + c.lut_rep_shift));
    }
#endif
  }

  // quantise an already-normalised pair n2 (D values) onto the minifloat grid; sign handled here.
  // `sign_src`: bits whose sign the RNE flavour keeps -- the un-normalised input, because the
  // reciprocal+FMA quotient turns -0.0 into +0.0 (the scale is never negative, so sign(w) == sign(w/S))
  IWQ_HD T2 grid(T2 n2, uint32_t sign_src, const FpEntry* tbl) const {
    constexpr int MB = D::kMantBits;
    const uint32_t nb = as_u32(n2);
    T2 a = from_u32<T2>(nb & 0x7FFF7FFFu);
    a = __hmin2(a, fmax2);   // clamp(+-fmax); a NaN becomes fmax here and is restored by the * S below
    a = __hmax2(a, tiny2);
    const uint32_t ab = as_u32(a);
    const uint32_t i0 = ((ab >> MB) & ((1u << (15 - MB)) - 1u)) - tbl_off;
    const uint32_t i1 = (ab >> (16 + MB)) - tbl_off;
    const uint2 e0 = *reinterpret_cast<const uint2*>(tbl + i0);
    const uint2 e1 = *reinterpret_cast<const uint2*>(tbl + i1);
    const uint32_t magic = (e0.x & 0xFFFFu) | (e1.x << 16);
    const uint32_t bound = (e0.x >> 16) | (e1.x & 0xFFFF0000u);
    const uint32_t thr = (e0.y & 0xFFFFu) | (e1.y << 16);
    const uint32_t bval = (e0.y >> 16) | (e1.y & 0xFFFF0000u);
    T2 v = __hsub2(__hadd2(a, from_u32<T2>(magic)), from_u32<T2>(magic));
    v = __hmin2(v, from_u32<T2>(bound));
    const uint32_t bump = __hge2_mask(a, from_u32<T2>(thr));
    uint32_t vb = (as_u32(v) & ~bump) | (bval & bump);
    const uint32_t sgn = rne ? (sign_src & 0x80008000u) : (__hlt2_mask(n2, from_u32<T2>(0u)) & 0x80008000u);
    return from_u32<T2>(vb | sgn);
  }

  // Same function as grid(), evaluated through the step table: one 8-byte shared-memory load per element, a
  // packed compare and a select.  No per-binade constants, no table walk.  The kernel is bound by the ALU pipe
  // (LOP3 / SHF / PRMT / HMNMX2 / HSET2 issue at half rate), so the sequence is written to minimise those:
  //   * fp16: the table starts at bit pattern 0 (kFromZero), so there is no lower clamp and no offset subtract;
  //   * |n| is a source modifier of the min, the sign comes from n + (+0.0) (an FMA-pipe HADD2 that turns the
  //     -0.0 of an underflowed quotient into +0.0: sign = (n < 0), quant_linear.py:130) and one LOP3;
  //   * the three packed constants are gathered with one PRMT each.
  static constexpr bool kFromZero = (DT == IWQ_F16);
  // magnitude part: the selected `below` / `above` entries of both halves, packed
  IWQ_HD uint32_t steps_select(T2 n2) const {
    T2 a = __hmin2(__habs2(n2), fmax2);
#ifdef __CUDA_ARCH__
    if (LUT == LUT_DIRECT) {
      a = __hmax2(a, lutlo2);
      const uint32_t d = as_u32(a);
      uint16_t v0, v1;
      // low half: pattern * 2; high half: (d >> 16) * 2 == d >> 15 because bit 15 (the low half's sign) is clear
      uint32_t a0;
      asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(a0) : "r"(d & 0xFFFFu), "r"(rep_base));
      asm("ld.shared.u16 %0, [%1];" : "=h"(v0) : "r"(a0));
      asm("ld.shared.u16 %0, [%1];" : "=h"(v1) : "r"(rep_base + (d >> 15)));
      return (uint32_t)v0 | ((uint32_t)v1 << 16);
    }
    if (REPL) {
      a = __hmax2(a, lutlo2);
      const uint32_t d = as_u32(a);
      uint32_t a0, a1;
      asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a0) : "r"(d & (rep_mask_hi >> 16)), "r"(rep_mul_lo), "r"(rep_base));
      a1 = rep_base + ((d & rep_mask_hi) >> rep_mul_hi);  // (a mad.hi.u32 drags a 64-bit addend along: two extra moves)
      uint2 e0, e1;
      asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(e0.x), "=r"(e0.y) : "r"(a0));
      asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(e1.x), "=r"(e1.y) : "r"(a1));
      const uint32_t thr = __byte_perm(e0.x, e1.x, 0x5410);
      const uint32_t below = __byte_perm(e0.x, e1.x, 0x7632);
      const uint32_t above = __byte_perm(e0.y, e1.y, 0x5410);
      const uint32_t m = __hge2_mask(a, from_u32<T2>(thr));
      return (below & ~m) | (above & m);
    }
#endif
    uint32_t d = as_u32(a);
    if (!kFromZero) {
      a = __hmax2(a, lutlo2);
      d = as_u32(a) - lo2;  // per-half pattern offset; no borrow (a >= lo in both halves)
    }
    const uint32_t i0 = (d << 16) >> (16u + sh);
    const uint32_t i1 = d >> (16u + sh);
    uint2 e0, e1;
#ifdef __CUDA_ARCH__
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(e0.x), "=r"(e0.y) : "r"(steps_saddr + i0 * 8u));
    asm("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(e1.x), "=r"(e1.y) : "r"(steps_saddr + i1 * 8u));
    const uint32_t thr = __byte_perm(e0.x, e1.x, 0x5410);
    const uint32_t below = __byte_perm(e0.x, e1.x, 0x7632);
    const uint32_t above = __byte_perm(e0.y, e1.y, 0x5410);
#else
    e0 = *reinterpret_cast<const uint2*>(steps + i0);
    e1 = *reinterpret_cast<const uint2*>(steps + i1);
    const uint32_t thr = (e0.x & 0xFFFFu) | (e1.x << 16);
    const uint32_t below = (e0.x >> 16) | (e1.x & 0xFFFF0000u);
    const uint32_t above = (e0.y & 0xFFFFu) | (e1.y << 16);
#endif
    const uint32_t m = __hge2_mask(a, from_u32<T2>(thr));
    return (below & ~m) | (above & m);
  }
  // sign bits of the codec: (n < 0), i.e. the sign of n + (+0.0)
  IWQ_HD uint32_t codec_sign(T2 n2) const { return as_u32(__hadd2(n2, from_u32<T2>(0u))) & 0x80008000u; }

  IWQ_HD T2 grid_steps(T2 n2, uint32_t sign_src) const {
    const uint32_t vb = steps_select(n2);
    const uint32_t ss = rne ? sign_src : as_u32(__hadd2(n2, from_u32<T2>(0u)));
    return from_u32<T2>(vb | (ss & 0x80008000u));
  }

  // ASYM: 1 / 0 fixes the asymmetric (mid/span) form at compile time (the kernels pass their SYM template
  // constant); -1 reads the runtime flag (test emulator).
  template <bool SLOW, int ASYM = -1>
  IWQ_HD uint32_t apply(uint32_t w2, T2* v2) const {
    const bool as = ASYM < 0 ? asym : (ASYM != 0);
    T2 w = from_u32<T2>(w2);
    if (as) w = __hsub2(w, mid2);                                     // weight - zeros            :755
    T2 n = SLOW ? ieee_quot2<DT>(w, S) : fast_quot2<DT>(w, S, r);     // / scales
    T2 v = grid_steps(n, w2);
    // a NaN quotient (only possible when S is 0/inf/NaN, i.e. the slow path) must stay NaN: the clamp removed it
    if (SLOW) v = from_u32<T2>(as_u32(v) | __hneu2_mask(n, n));
    *v2 = v;
    // _rn forms: the product is rounded to D before the add (two ATen ops); plain __hmul2 + __hadd2 may be
    // contracted into one HFMA2 by ptxas once `as` is a compile-time constant
    T2 o = __hmul2_rn(v, S2);                                         // decoded.to(D) * scales    :774
    if (as) o = __hadd2_rn(o, midh2);                                 // + zeros.half().to(D)      :776
    return as_u32(o);
  }
};

// ------------------------------------------------------------------------------------------------
// block floating point          quant_linear.py:648-723 (+ _rounding_rshift :112-123)
//
// The reference works on the fp16 bit fields: m = (implicit one << 10 | mantissa) >> (E_max - E), then
// (m + 2^(sd-1)) >> sd, clamp to 2^target - 1, times 2^(E_max - 15 - (target-1)).  With
//     a_ref = (implicit one << 10 | mantissa) * 2^(E - 25)   (= |w|, except that exponent field 0 counts HALF:
//                                                              the reference treats field 0 like exponent -15)
// and step = 2^(E_max - 25 + sd) this is   q = floor(a_ref / step + 1/2)   (truncate-then-round-half-up collapses
// to one floor), clamped, times step.  All of it is exact in fp32: a_ref/step has <= 11 significant bits,
// adding 2^-12 turns round-half-even (the magic-number add) into round-half-up without touching non-ties.
// ------------------------------------------------------------------------------------------------
struct BfpQ {
  float step, inv_step, top, tie;
};

template <int DT>
IWQ_HD BfpQ bfp_segment_params(float absmax, const QuantConsts& c) {
  // E_max = exponent field of the group's largest magnitude after the cast to fp16 (:656, :668)
  const uint32_t hb = DType<IWQ_F16>::from_f32(absmax);
  const int eb = (int)((hb >> 10) & 31u);
  BfpQ q;
  q.step = bits_f32((uint32_t)(eb - 25 + c.bfp_sd + 127) << 23);
  q.inv_step = bits_f32((uint32_t)(25 - eb - c.bfp_sd + 127) << 23);
  q.top = (float)((1 << c.bfp_target) - 1) + 0.25f;  // floor(min(x, M + 1/4) + 1/2) == min(floor(x + 1/2), M)
  // sd > 0: round half up = RNE(x + 2^-12).  sd == 0 (w_bit >= 12): no rounding shift at all, plain truncation
  // (:674-699) = floor(x) = round-half-up(x - 1/2).
  q.tie = c.bfp_sd > 0 ? 0.000244140625f : (0.000244140625f - 0.5f);
  return q;
}

// exact fp32 product of two fp16 values (one FHFMA on sm_100a: fma.rn.f32.f16)
IWQ_HD float mul_f16_to_f32(uint16_t a, uint16_t b) {
#ifdef __CUDA_ARCH__
  float r;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(a), "h"(b), "f"(0.0f));
  return r;
#else
  return DType<IWQ_F16>::to_f32(a) * DType<IWQ_F16>::to_f32(b);
#endif
}

template <int DT, bool PAIR = false>
struct BfpApply {
  using D = DType<DT>;
  using T2 = typename D::T2;
  using F = typename PerHalf<PAIR>::F;
  F step, inv_step;
  float top, tie;
  bool round_mant;

  IWQ_HD void setup(const BfpQ& q, const QuantConsts& c) { setup2(q, q, c); }
  IWQ_HD void setup2(const BfpQ& q, const BfpQ& qb, const QuantConsts& c) {
    step = PerHalf<PAIR>::make(q.step, qb.step);
    inv_step = PerHalf<PAIR>::make(q.inv_step, qb.inv_step);
    top = q.top;
    tie = q.tie;
    round_mant = c.bfp_round_mant != 0;
  }

  IWQ_HD uint32_t apply(uint32_t w2) const {
    // .to(torch.float16)                                                                     :656
    uint32_t h2 = w2;
    if (DT != IWQ_F16) {
      float2 f = D::to_f32x2(from_u32<T2>(w2));
      __half2 h = __floats2half2_rn(f.x, f.y);
      h2 = *reinterpret_cast<uint32_t*>(&h);
    }
    const uint32_t ab = h2 & 0x7FFF7FFFu;
    // exponent field 0 (|w| < 2^-14): no implicit one AND the field is used as-is -> a_ref = |w| / 2   :664-670
    const uint32_t sub = __hlt2_mask(from_u32<__half2>(ab), from_u32<__half2>(0x04000400u));
    const uint32_t mul2 = 0x3C003C00u ^ (sub & 0x04000400u);  // 1.0, or 0.5 where the field is 0
#ifdef __CUDA_ARCH__
    {
      // packed fp32x2 form of the scalar code below (same operations in the same order, two lanes per FADD2 / FMUL2)
      float2 x = __fmul2_rn(make_float2(mul_f16_to_f32((uint16_t)(ab & 0xFFFFu), (uint16_t)(mul2 & 0xFFFFu)),
                                        mul_f16_to_f32((uint16_t)(ab >> 16), (uint16_t)(mul2 >> 16))),
                            make_float2(PerHalf<PAIR>::lo(inv_step), PerHalf<PAIR>::hi(inv_step)));
      x.x = fminf(x.x, top);                                                               // clamp  :700-701
      x.y = fminf(x.y, top);
      const float2 m2 = make_float2(12582912.0f, 12582912.0f), nm2 = make_float2(-12582912.0f, -12582912.0f);
      x = __fadd2_rn(__fadd2_rn(__fadd2_rn(x, make_float2(tie, tie)), m2), nm2);           // floor(x + 1/2)
      if (DT == IWQ_BF16 && round_mant) {                                                  // mant_rounded.to(bf16)   :707
        x.x = rD<DT>(x.x);
        x.y = rD<DT>(x.y);
      }
      const float2 s2 = make_float2(bits_f32(f32_bits(PerHalf<PAIR>::lo(step)) | ((h2 << 16) & 0x80000000u)),   // scale * sign_factor :705-707
                                    bits_f32(f32_bits(PerHalf<PAIR>::hi(step)) | (h2 & 0x80000000u)));
      const float2 o = __fmul2_rn(x, s2);
      return as_u32(D::from_f32x2(o.x, o.y));                                              // .to(weight dtype)       :708
    }
#endif
    float x0 = mul_f16_to_f32((uint16_t)(ab & 0xFFFFu), (uint16_t)(mul2 & 0xFFFFu)) * PerHalf<PAIR>::lo(inv_step);
    float x1 = mul_f16_to_f32((uint16_t)(ab >> 16), (uint16_t)(mul2 >> 16)) * PerHalf<PAIR>::hi(inv_step);
    x0 = fminf(x0, top);                                                                   // clamp  :700-701
    x1 = fminf(x1, top);
    const float kMagic = 12582912.0f;                         // 1.5 * 2^23
    float q0 = ((x0 + tie) + kMagic) - kMagic;                // floor(x + 1/2): align shift + _rounding_rshift
    float q1 = ((x1 + tie) + kMagic) - kMagic;
    if (DT == IWQ_BF16 && round_mant) {                        // mant_rounded.to(bf16)               :707 (exact no-op for fp16: <= 11 bits)
      q0 = rD<DT>(q0);
      q1 = rD<DT>(q1);
    }
    const float s0 = bits_f32(f32_bits(PerHalf<PAIR>::lo(step)) | ((h2 << 16) & 0x80000000u));  // scale * sign_factor :705-707
    const float s1 = bits_f32(f32_bits(PerHalf<PAIR>::hi(step)) | (h2 & 0x80000000u));
    return as_u32(D::from_f32x2(q0 * s0, q1 * s1));            // .to(weight dtype)                    :708
  }
};

// ------------------------------------------------------------------------------------------------
// double-approximate decode of FOUR codes       fp_decode_aligned_double_approx, quant_linear.py:288-363
// (align_subnorm_exp_as_one=True, handle_max_outlier=True: the driver's arguments, :511-523 etc.)
// The reference keeps sign / exponent / mantissa in torch.int8 tensors, so `1 << (shift-1)`, `val + offset`
// and the left shift wrap modulo 256; wrap8() reproduces that, garbage included.
// Returns the four decoded values as D bit patterns.
// ------------------------------------------------------------------------------------------------
IWQ_HD int wrap8(int x) { return (int)(int8_t)(uint8_t)(x & 0xFF); }

IWQ_HD int rounding_rshift8(int val, int shift) {  // _rounding_rshift (:112-123) on int8 tensors
  const int off = shift > 0 ? wrap8(shift - 1 < 31 ? (1 << (shift - 1)) : 0) : 0;
  const int v = wrap8(val + off);
  return shift >= 31 ? (v < 0 ? -1 : 0) : (v >> shift);
}

// per-code part of the decode: aligned exponent (:315) and padded mantissa (:316-321)
IWQ_HD void double_approx_fields(uint32_t code, const QuantConsts& c, int* ae, int* mp) {
  const int E = c.exp_bits, M = c.mant_bits, tp = c.tail_pad;
  const int ef = (int)((code >> M) & ((1u << E) - 1u)), mf = (int)(code & ((1u << M) - 1u));
  *ae = ef == 0 ? 1 : ef;                                                                // :315
  const int full = ((ef == 0 ? 0 : 1) << M) | mf;                                        // :316-317
  *mp = tp >= 0 ? wrap8(full << tp) : rounding_rshift8(full, -tp);                       // :318-321
}

// value of ONE code once the group's target exponent is known (:345-362); `code` includes the sign bit
template <int DT>
IWQ_HD uint16_t double_approx_elem(uint32_t code, int target, const QuantConsts& c) {
  const int E = c.exp_bits, M = c.mant_bits, bias = (1 << (E - 1)) - 1, tp = c.tail_pad;
  int ae, mp;
  double_approx_fields(code, c, &ae, &mp);
  const int cap = tp >= 0 ? (((1 << (M + 1)) - 1) << tp) : (((1 << (M + 1)) - 1) >> (-tp));  // :351-355
  const float hi_unb = rD<DT>((float)(target - bias));                                   // (target - bias).to(dtype)
  const float p2 = rD<DT>(exp2f(hi_unb));                                                // torch.pow(2.0 (dtype), .)
  const float den = ldexpf(1.0f, M + tp);
  const int shift = target - ae;                                                         // :345
  int al;
  if (shift >= 0) {
    al = rounding_rshift8(mp, shift);                                                    // :348
  } else {
    const int ls = -shift;
    al = wrap8(ls < 31 ? (mp << ls) : 0);                                                // :349 (int8 wrap)
    al = al < cap ? al : cap;                                                            // :356
  }
  float v = rD<DT>(rD<DT>((float)al) / den);                                             // mant.to(dtype) / 2^(M+tail)  :358
  v = rD<DT>(v * p2);
  const uint32_t sign = (code >> (E + M)) & 1u;
  uint16_t b = DType<DT>::from_f32(v);
  if (sign) b ^= 0x8000u;                                                                // torch.where(sign == 1, -value, value)
  if (code == 0) b = 0;                                                                  // zero_mask
  return b;
}

// the group's target exponent from the four codes (:334-343)
IWQ_HD int double_approx_target(const uint32_t (&code)[4], const QuantConsts& c) {
  const int max_exp = (1 << c.exp_bits) - 1;
  int cnt = 0, gmax = 0;
  bool has_max = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int ae, mp;
    double_approx_fields(code[i], c, &ae, &mp);
    const bool outl = ae < c.hi_start || ae > c.hi_field;                                // :334
    cnt += outl ? 1 : 0;
    gmax = ae > gmax ? ae : gmax;
    has_max = has_max || (ae == max_exp && outl);                                        // :342
  }
  int target = cnt <= 1 ? c.hi_field : gmax;                                             // :338
  if (has_max) target = max_exp;                                                         // :343
  return target;
}

template <int DT>
IWQ_HD void double_approx_decode4(const uint32_t (&code)[4], const QuantConsts& c, uint16_t (&out)[4]) {
  const int target = double_approx_target(code, c);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = double_approx_elem<DT>(code[i], target, c);
}

// Device form of the same decode: everything per-code is tabulated (iwq_consts.h: build_double_tables).
//   * the step table's `below` / `above` hold a 16-bit INFO word of the grid value instead of the value:
//       code magnitude [0:6] | outlier [7] | aligned exponent [8:12] | outlier-at-max-exponent [13]
//     so the packed quantiser (steps_select) yields the info of both halves of a pair in one go;
//   * a value table val[target << 8 | sign << 7 | code magnitude] holds double_approx_elem for every
//     (target exponent 0..31, 8-bit code): 16 KB in shared memory.
// The group logic runs packed on the two halves (= two neighbouring columns = two independent groups): the
// fields are positioned so that sums cannot carry between halves and compare as positive fp16 patterns.
constexpr uint32_t kDblCodeMask = 0x007F007Fu, kDblOutl = 0x00800080u, kDblAe = 0x1F001F00u, kDblMaxOutl = 0x20002000u;
constexpr int kDblValEntries = 32 * 256;

// packed group logic on the INFO words of the four codes of a group -- two groups at once, one per 16-bit half:
// returns the groups' target exponents in the aligned-exponent position (bits 8..12 of each half)      :334-343
template <int DT>
IWQ_HD uint32_t double_approx_target2(const uint32_t (&info)[4], uint32_t hi_field2, uint32_t max_exp2) {
  using T2 = typename DType<DT>::T2;
  const uint32_t cnt = (info[0] & kDblOutl) + (info[1] & kDblOutl) + (info[2] & kDblOutl) + (info[3] & kDblOutl);  // k * 0x80 per half
  const uint32_t few = __hle2_mask(from_u32<T2>(cnt), from_u32<T2>(kDblOutl));             // outlier_count <= 1      :338
  const T2 g01 = __hmax2(from_u32<T2>(info[0] & kDblAe), from_u32<T2>(info[1] & kDblAe));  // positive patterns: float max == integer max
  const T2 g23 = __hmax2(from_u32<T2>(info[2] & kDblAe), from_u32<T2>(info[3] & kDblAe));
  const uint32_t gmax = as_u32(__hmax2(g01, g23));
  uint32_t target = (gmax & ~few) | (hi_field2 & few);
  const uint32_t anymax = (info[0] | info[1] | info[2] | info[3]) & kDblMaxOutl;
  const uint32_t hm = __hne2_mask(from_u32<T2>(anymax), from_u32<T2>(0u));                 // has_max_outlier          :342-343
  return (target & ~hm) | (max_exp2 & hm);
}

IWQ_HD uint16_t double_approx_info(uint32_t mag_code, const QuantConsts& c) {
  int ae, mp;
  double_approx_fields(mag_code, c, &ae, &mp);
  const int max_exp = (1 << c.exp_bits) - 1;
  const bool outl = ae < c.hi_start || ae > c.hi_field;
  return (uint16_t)((mag_code & 0x7Fu) | (outl ? 0x80u : 0u) | ((uint32_t)ae << 8) | ((outl && ae == max_exp) ? 0x2000u : 0u));
}

// minifloat code of a grid value (for optional code emission): sign | exponent field | mantissa,
// zero -> 0 unless the sign survived (quirk Q4).                     quant_linear.py:156-161
template <int DT>
IWQ_HD uint8_t fp_code_of(uint16_t vbits, int E, int M) {
  float v = DType<DT>::to_f32(vbits);
  uint32_t sign = (vbits >> 15) & 1u;
  float a = fabsf(v);
  int bias = (1 << (E - 1)) - 1;
  int emin = 1 - bias;
  uint32_t ef = 0, mf = 0;
  if (a > 0.0f) {
    int e;
    float m = frexpf(a, &e);  // a = m * 2^e, m in [0.5, 1)
    e -= 1;                   // a = (2m) * 2^e, 2m in [1, 2)
    if (e < emin) {
      mf = (uint32_t)(a * exp2f((float)(M - emin)));
    } else {
      ef = (uint32_t)(e + bias);
      mf = (uint32_t)((2.0f * m - 1.0f) * (float)(1 << M));
    }
  }
  return (uint8_t)((sign << (E + M)) | (ef << M) | mf);
}

}  // namespace iwq
