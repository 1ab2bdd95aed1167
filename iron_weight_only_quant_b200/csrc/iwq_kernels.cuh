// iwq_kernels.cuh -- sm_100a kernels of the weight fake-quant path.
//
// Data layout in HBM: every weight matrix is a row-major contiguous array of 16-bit values, so a
// matrix at per-group granularity is nothing but a flat run of independent `group`-element segments,
// and at per-channel granularity a flat run of `cols`-element segments.  All kernels are HBM-bound
// streaming passes (4 algorithmic bytes per element: one 16-bit read, one 16-bit write); none uses
// tensor cores.
//
//   seg_tma_kernel    small segments (group = 8..1024): persistent CTAs, one producer lane streams
//                     16 KB tiles global->shared with 1-D TMA bulk copies (cp.async.bulk + mbarrier
//                     complete_tx) through an n-stage ring; 8 consumer warps pull 128-bit vectors from
//                     shared memory, reduce min/max with warp shuffles inside the lanes of a segment,
//                     derive scale / zero-point in registers, quantise-dequantise and store 128-bit.
//   seg_direct_kernel same consumer code fed by plain 128-bit global loads (A/B variant).
//   row_tma_kernel    long segments (per-channel rows): one row per ring stage (one bulk copy per row),
//                     two sweeps over the row in shared memory (range, then quantise).
//   minmax_kernel     per-tensor granularity, pass 1: per-matrix {max|absmax, -min} via ordered atomics.
//                     Pass 2 is seg_*_kernel<GIVEN=true> reading that range (after an NCCL MAX
//                     all-reduce when the matrix is sharded across GPUs).
#pragma once

#include <cuda_runtime.h>

#include <type_traits>

#include "iwq_consts.h"
#include "iwq_math.cuh"

namespace iwq {

// ------------------------------------------------------------------------------------------------
// launch-wide parameter blocks
// ------------------------------------------------------------------------------------------------
struct Job {               // one matrix (64 bytes)
  const uint8_t* src;
  uint8_t* dst;
  uint8_t* scales;         // optional, D[G]
  uint8_t* zeros;          // optional, D[G]
  uint8_t* codes;          // optional, uint8[rows*cols]
  uint64_t bytes;          // rows*cols*2
  uint32_t first_unit;     // first tile (seg kernels) / first row (row kernel) of this job in the launch
  uint32_t row_bytes;      // cols*2
  uint64_t pad;
};

// Consumer warps per CTA of the segment kernels; a tile is kConsumerWarps x 2 KB.  Measured on the headline workload (one
// B200, same box, interleaved rounds; tools/r2_exp9..11.sh, tools/tmabench.cu): 8 warps / 16 KB tiles x 3 stages x 3 CTAs per SM
// 5.90-6.17 TB/s; 16 / 32 KB x 4 x 1: 6.09-6.38; 20 / 40 KB x 4 x 1: 6.15-6.51; 24 / 48 KB x 4 x 1: 6.05-6.52 -- the memory system
// prefers ONE ring of large bulk copies per SM to three rings of small ones (the copy-only probe shows the same: 16 KB x 3 x 3
// 6.19, 32 KB x 3 x 1 6.62 TB/s), and 24 warps are what the element stage needs to keep up.
#ifndef IWQ_CW
#define IWQ_CW 24
#endif
#ifndef IWQ_SEG4_MINB
#define IWQ_SEG4_MINB 1   // __launch_bounds__ minimum CTAs per SM of seg4_tma_kernel
#endif
constexpr int kConsumerWarps = IWQ_CW;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kVPT = 4;                                  // 128-bit vectors per thread per tile
constexpr int kTileBytes = kConsumers * kVPT * 16;       // 16 KB
constexpr int kMaxStages = 8;
constexpr int kSmemHeader = 1024;                        // barriers + tile contexts + range slots
constexpr int kSeg4MaxStages = (kSmemHeader - 640) / (4 * kConsumerWarps) < kMaxStages ? (kSmemHeader - 640) / (4 * kConsumerWarps) : kMaxStages;  // range slots [stage][warp]

struct LaunchParams {
  const Job* jobs;         // device array (n_jobs > 1) or nullptr -> `single`
  Job single;
  int32_t n_jobs;
  uint32_t total_units;    // tiles or rows over all jobs
  int32_t n_stages;
  uint32_t stage_bytes;    // row kernel: bytes reserved per ring stage
  int32_t lps;             // lanes per segment (1..32)      } group = 8 * lps * vpl elements
  int32_t vpl;             // vectors per lane per segment   }
  const float* ranges;     // GIVEN mode: {max|absmax, -min} per job
  float* ranges_out;       // minmax kernel output
  int32_t* status;
  const FpStep* lut;       // KIND_FP: device step table (qc.lut_n entries), copied into shared memory by every CTA
  uint32_t lut_smem_bytes; // bytes reserved for it in front of the ring (multiple of 1024; 0 for INT kinds)
  uint32_t lut_extra_bytes;// double-approximate decode: value table behind the step table (device: lut + lut_smem_bytes)
  uint32_t row_tma_out;    // row kernel: results go back into the stage and leave through one bulk store per warp
  uint32_t col_group;      // quant_dim = 1: rows per segment (0: the whole column, i.e. per-channel)
  uint32_t col_tile;       // col_reg_kernel: columns per tile (64 / 128 / 256), 0 = col_kernel
  uint32_t col_rpl;        // col_reg_kernel: vectors (rows) per lane = col_group * (col_tile / 64) / 32
  uint32_t unit_begin;     // first unit of this launch (a launch may cover a SLAB [unit_begin, total_units) of the work list)
  uint32_t reverse;        // seg4 / GIVEN: walk the slab's tiles from its end (per-tensor apply pass: the tiles the range pass
                           // read last are still in L2)
  uint32_t l2_hint;        // bit 0: bulk loads evict_first, bit 1: bulk loads evict_last, bit 2: bulk stores evict_first,
                           // bit 3: minmax kernel walks tiles strided (blockIdx.x + k * gridDim.x) instead of in contiguous runs,
                           // bit 4: bulk stores evict_last
  uint32_t seg_vpl;        // seg4: 16-byte vectors per lane (4 / 8 / 16); lps = group / (8 * seg_vpl)
  uint32_t col_wide;       // quant_dim = 1: 1 = col_wide_kernel (128 x 256 tiles, CTA-wide ring) instead of col_tma_kernel
  uint32_t col_rows;       // ... rows per tile (128 or 64)
  uint32_t tile_map;       // seg4: 0 = CTA b takes tiles b, b + grid, b + 2 grid, ... (a window of grid tiles slides over the work
                           // list); 1 = CTA b takes the contiguous run [b n / grid, (b + 1) n / grid) -- measured faster for the
                           // in-place read+write stream (tools/tmabench.cu)
  QuantConsts qc;
};


// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of burning issue slots in a software spin (ncu showed ~30 spins per tile without it)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared::cta bulk copy, completion signalled on `bar` as transaction bytes (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// the same with an L2 eviction-priority hint (createpolicy): evict_last keeps a tile that will be re-read soon
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// shared::cta -> global bulk store (SASS: UBLKCP), tracked by the thread's bulk async-group
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_1d_hint(void* gmem_dst, const void* smem_src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before a bulk store reads them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void sts128(void* p, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------
// range of one 128-bit vector, packed as a T2 pair {max|absmax, -min} (both reduce with max)
// ------------------------------------------------------------------------------------------------
template <int DT, bool SYM>
__device__ __forceinline__ uint32_t vec_range(const uint4& v) {
  using T2 = typename DType<DT>::T2;
  if (SYM) {
    T2 a = from_u32<T2>(v.x & 0x7FFF7FFFu), b = from_u32<T2>(v.y & 0x7FFF7FFFu);
    T2 c = from_u32<T2>(v.z & 0x7FFF7FFFu), d = from_u32<T2>(v.w & 0x7FFF7FFFu);
    T2 m = __hmax2_nan(__hmax2_nan(a, b), __hmax2_nan(c, d));
    uint32_t mb = as_u32(m);
    T2 sw = from_u32<T2>(__byte_perm(mb, mb, 0x1032));
    return as_u32(__hmax2_nan(m, sw));
  } else {
    T2 a = from_u32<T2>(v.x), b = from_u32<T2>(v.y), c = from_u32<T2>(v.z), d = from_u32<T2>(v.w);
    T2 mx = __hmax2_nan(__hmax2_nan(a, b), __hmax2_nan(c, d));
    T2 mn = __hmin2_nan(__hmin2_nan(a, b), __hmin2_nan(c, d));
    uint32_t xb = as_u32(mx), nb = as_u32(mn);
    mx = __hmax2_nan(mx, from_u32<T2>(__byte_perm(xb, xb, 0x1032)));
    mn = __hmin2_nan(mn, from_u32<T2>(__byte_perm(nb, nb, 0x1032)));
    return __byte_perm(as_u32(mx), as_u32(mn) ^ 0x80008000u, 0x5410);  // {max, -min}
  }
}

template <int DT>
__device__ __forceinline__ uint32_t range_max(uint32_t a, uint32_t b) {
  using T2 = typename DType<DT>::T2;
  return as_u32(__hmax2_nan(from_u32<T2>(a), from_u32<T2>(b)));
}

template <int DT>
__device__ __forceinline__ void range_unpack(uint32_t pk, float* mx, float* mn) {
  *mx = DType<DT>::to_f32((uint16_t)(pk & 0xFFFFu));
  *mn = -DType<DT>::to_f32((uint16_t)(pk >> 16));
}

// ------------------------------------------------------------------------------------------------
// element stage on 128-bit vectors: Prepared<DT,KIND> holds one segment's parameters in registers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_codes(const uint32_t (&cb)[8], uint2* codes8) {
  codes8->x = cb[0] | (cb[1] << 8) | (cb[2] << 16) | (cb[3] << 24);
  codes8->y = cb[4] | (cb[5] << 8) | (cb[6] << 16) | (cb[7] << 24);
}

// store the 8 codes of one 16-byte vector: 8 bytes, or 4 bytes in the packed W4 layout (low nibble = even element)
__device__ __forceinline__ void store_codes(uint8_t* codes, uint64_t vec, const uint2& c8, bool nibbles) {
  if (!nibbles) {
    *reinterpret_cast<uint2*>(codes + vec * 8) = c8;
  } else {
    uint32_t a = c8.x & 0x0F0F0F0Fu, b = c8.y & 0x0F0F0F0Fu;
    a = (a | (a >> 4)) & 0x00FF00FFu;
    b = (b | (b >> 4)) & 0x00FF00FFu;
    a = (a | (a >> 8)) & 0xFFFFu;
    b = (b | (b >> 8)) & 0xFFFFu;
    *reinterpret_cast<uint32_t*>(codes + vec * 4) = a | (b << 16);
  }
}

// any NaN among the 8 values of a vector (suspect segments only: see IntQ::bad)
template <int DT>
__device__ __forceinline__ bool vec_has_nan(const uint4& v) {
  using T2 = typename DType<DT>::T2;
  const T2 a = from_u32<T2>(v.x), b = from_u32<T2>(v.y), c = from_u32<T2>(v.z), d = from_u32<T2>(v.w);
  return (__hneu2_mask(a, a) | __hneu2_mask(b, b) | __hneu2_mask(c, c) | __hneu2_mask(d, d)) != 0u;
}

template <int DT, int KIND>
struct Prepared;

template <int DT, int KIND>
struct PreparedInt {  // KIND_INT_FAST / KIND_INT_WIDE
  using D = DType<DT>;
  using T2 = typename D::T2;
  IntFast<DT, KIND == KIND_INT_WIDE> f;
  float scale, zero;
  bool bad;
  __device__ __forceinline__ void setup(float mx, float mn, const QuantConsts& c, bool sym, const FpStep*) {
    IntQ q = int_segment_params<DT>(mx, mn, c, sym);
    f.setup(q);
    scale = q.s;
    zero = q.z;
    bad = q.bad;
  }
  __device__ __forceinline__ bool slow() const { return f.slow_div; }
  template <bool SYM, int SLOW = -1>
  __device__ __forceinline__ uint4 apply(const uint4& v, const QuantConsts&, const FpStep*, bool want_codes, uint2* codes8) const {
    T2 d[4];
    uint4 o;
    o.x = f.template apply<SLOW>(v.x, &d[0]);
    o.y = f.template apply<SLOW>(v.y, &d[1]);
    o.z = f.template apply<SLOW>(v.z, &d[2]);
    o.w = f.template apply<SLOW>(v.w, &d[3]);
    if (want_codes) {
      uint32_t cb[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 a = D::to_f32x2(d[i]);
        cb[2 * i] = (uint32_t)(int)(a.x + zero) & 0xFFu;  // code = clamp(rint(x), lo, hi) + z
        cb[2 * i + 1] = (uint32_t)(int)(a.y + zero) & 0xFFu;
      }
      pack_codes(cb, codes8);
    }
    return o;
  }
};
template <int DT>
struct Prepared<DT, KIND_INT_FAST> : PreparedInt<DT, KIND_INT_FAST> {};
template <int DT>
struct Prepared<DT, KIND_INT_WIDE> : PreparedInt<DT, KIND_INT_WIDE> {};

template <int DT>
struct Prepared<DT, KIND_INT_GENERIC> {
  IntQ q;
  float scale, zero;
  bool bad;
  __device__ __forceinline__ void setup(float mx, float mn, const QuantConsts& c, bool sym, const FpStep*) {
    q = int_segment_params<DT>(mx, mn, c, sym);
    scale = q.s;
    zero = q.z;
    bad = q.bad;
  }
  __device__ __forceinline__ bool slow() const { return false; }
  template <bool SYM, int SLOW = -1>
  __device__ __forceinline__ uint4 apply(const uint4& v, const QuantConsts& c, const FpStep*, bool want_codes, uint2* codes8) const {
    const uint32_t in[4] = {v.x, v.y, v.z, v.w};
    uint32_t out[4], cb[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float k0, k1;
      uint32_t lo = int_generic_elem<DT>((uint16_t)(in[i] & 0xFFFFu), q, c, &k0);
      uint32_t hi = int_generic_elem<DT>((uint16_t)(in[i] >> 16), q, c, &k1);
      out[i] = lo | (hi << 16);
      cb[2 * i] = (uint32_t)(int)k0 & 0xFFu;
      cb[2 * i + 1] = (uint32_t)(int)k1 & 0xFFu;
    }
    if (want_codes) pack_codes(cb, codes8);
    return make_uint4(out[0], out[1], out[2], out[3]);
  }
};

template <int DT, int LUT>
struct PreparedFp {
  using T2 = typename DType<DT>::T2;
  FpTable<DT, false, LUT> f;
  float scale, zero;
  bool bad;
  __device__ __forceinline__ void setup(float mx, float mn, const QuantConsts& c, bool /*sym*/, const FpStep* steps) {
    FpQ q = fp_segment_params<DT>(mx, mn, c);
    f.setup(q, c, steps);
    scale = q.S;
    zero = q.mid;
    bad = q.bad;
  }
  template <bool SLOW, bool SYM>
  __device__ __forceinline__ uint4 apply_t(const uint4& v, T2 (&g)[4]) const {
    uint4 o;
    o.x = f.template apply<SLOW, SYM ? 0 : 1>(v.x, &g[0]);
    o.y = f.template apply<SLOW, SYM ? 0 : 1>(v.y, &g[1]);
    o.z = f.template apply<SLOW, SYM ? 0 : 1>(v.z, &g[2]);
    o.w = f.template apply<SLOW, SYM ? 0 : 1>(v.w, &g[3]);
    return o;
  }
  __device__ __forceinline__ bool slow() const { return f.slow_div; }
  template <bool SYM, int SLOW = -1>
  __device__ __forceinline__ uint4 apply(const uint4& v, const QuantConsts& c, const FpStep*, bool want_codes, uint2* codes8) const {
    T2 g[4];
    // IEEE division only for degenerate scales: a segment-uniform branch, per vector (SLOW < 0) or hoisted by the caller
    const uint4 o = SLOW < 0 ? (f.slow_div ? apply_t<true, SYM>(v, g) : apply_t<false, SYM>(v, g)) : (SLOW ? apply_t<true, SYM>(v, g) : apply_t<false, SYM>(v, g));
    if (want_codes) {
      uint32_t cb[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t gb = as_u32(g[i]);
        cb[2 * i] = fp_code_of<DT>((uint16_t)(gb & 0xFFFFu), c.exp_bits, c.mant_bits);
        cb[2 * i + 1] = fp_code_of<DT>((uint16_t)(gb >> 16), c.exp_bits, c.mant_bits);
      }
      pack_codes(cb, codes8);
    }
    return o;
  }
};

template <int DT>
struct Prepared<DT, KIND_FP> : PreparedFp<DT, LUT_STEP> {};

template <int DT>
struct Prepared<DT, KIND_BFP> {
  BfpApply<DT> f;
  float scale, zero;
  bool bad;
  __device__ __forceinline__ void setup(float mx, float /*mn*/, const QuantConsts& c, bool /*sym*/, const FpStep*) {
    f.setup(bfp_segment_params<DT>(mx, c), c);
    scale = 0.f;
    zero = 0.f;
    bad = false;
  }
  __device__ __forceinline__ bool slow() const { return false; }
  template <bool SYM, int SLOW = -1>
  __device__ __forceinline__ uint4 apply(const uint4& v, const QuantConsts&, const FpStep*, bool want_codes, uint2* codes8) const {
    if (want_codes) *codes8 = make_uint2(0u, 0u);  // the reference keeps no BFP codes (quant_linear.py:712-715)
    return make_uint4(f.apply(v.x), f.apply(v.y), f.apply(v.z), f.apply(v.w));
  }
};

// per-tile context handed from the producer to the consumers through shared memory
struct TileCtx {
  uint8_t* dst;      // output address of the tile's first byte
  uint8_t* scales;
  uint8_t* zeros;
  uint8_t* codes;
  uint64_t vec0;     // index (in 16-byte vectors, within the matrix) of the tile's first vector
  uint32_t n_vec;    // valid vectors in the tile
  float given_mx;    // GIVEN mode
  float given_mn;
  uint32_t pad;
};

// The sequence of tile numbers one CTA visits (see LaunchParams::tile_map).
struct TileWalk {
  uint32_t cur, end, step;
};
__device__ __forceinline__ TileWalk make_walk(const LaunchParams& p) {
  TileWalk w;
  if (p.tile_map == 1u) {
    const uint64_t n = p.total_units - p.unit_begin;
    w.cur = p.unit_begin + (uint32_t)(n * blockIdx.x / gridDim.x);
    w.end = p.unit_begin + (uint32_t)(n * (blockIdx.x + 1u) / gridDim.x);
    w.step = 1u;
  } else {
    w.cur = p.unit_begin + blockIdx.x;
    w.end = p.total_units;
    w.step = gridDim.x;
  }
  return w;
}

__device__ __forceinline__ const Job* job_table(const LaunchParams& p) { return p.n_jobs == 1 ? &p.single : p.jobs; }

// Track the job of `unit`, scanning forward from *cur (units are visited in increasing order).
// `*jb` is reloaded only when the job changes; *cur == -1 forces the first load.
__device__ __forceinline__ void advance_job(const LaunchParams& p, uint32_t unit, int* cur, Job* jb) {
  const Job* jobs = job_table(p);
  int j = *cur < 0 ? 0 : *cur;
  if (*cur < 0 && p.n_jobs > 8) {
    // first lookup: binary search (a CTA that starts deep inside the work list would otherwise chase up to n_jobs
    // dependent global loads: ~30 us for the 192 matrices of an OPT-6.7B shard)
    int lo = 0, hi = p.n_jobs - 1;  // invariant: jobs[lo].first_unit <= unit
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].first_unit <= unit) lo = mid; else hi = mid - 1;
    }
    j = lo;
  }
  while (j + 1 < p.n_jobs && unit >= jobs[j + 1].first_unit) ++j;
  while (j > 0 && unit < jobs[j].first_unit) --j;  // reverse walks (per-tensor apply pass) step backwards
  if (j != *cur) {
    *cur = j;
    *jb = jobs[j];
  }
}

__device__ __forceinline__ TileCtx make_tile_ctx(const LaunchParams& p, const Job& jb, int job_idx, uint32_t tile, bool given) {
  TileCtx t;
  const uint64_t off = (uint64_t)(tile - jb.first_unit) * kTileBytes;
  const uint64_t left = jb.bytes - off;
  t.dst = jb.dst + off;
  t.scales = jb.scales;
  t.zeros = jb.zeros;
  t.codes = jb.codes;
  t.vec0 = off >> 4;
  t.n_vec = (uint32_t)((left < (uint64_t)kTileBytes ? left : (uint64_t)kTileBytes) >> 4);
  t.given_mx = 0.f;
  t.given_mn = 0.f;
  if (given) {
    t.given_mx = p.ranges[2 * job_idx];
    t.given_mn = -p.ranges[2 * job_idx + 1];
  }
  t.pad = 0;
  return t;
}

// Consumer side of one tile: v[it] is the vector at tile-local index warp*128 + it*32 + lane.
template <int DT, int KIND, bool SYM, bool GIVEN>
__device__ __forceinline__ void consume_tile(uint4 (&v)[kVPT], const TileCtx& t, const LaunchParams& p, const FpStep* tbl, int warp,
                                             int lane) {
  const int lps = p.lps, vpl = p.vpl;
  uint32_t pk[kVPT];
  if (!GIVEN) {
#pragma unroll
    for (int it = 0; it < kVPT; ++it) pk[it] = vec_range<DT, SYM>(v[it]);
    if (vpl >= 2) {
      pk[0] = pk[1] = range_max<DT>(pk[0], pk[1]);
      pk[2] = pk[3] = range_max<DT>(pk[2], pk[3]);
    }
    if (vpl >= 4) pk[0] = pk[1] = pk[2] = pk[3] = range_max<DT>(pk[0], pk[2]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      if (o < lps) {
#pragma unroll
        for (int it = 0; it < kVPT; ++it) pk[it] = range_max<DT>(pk[it], __shfl_xor_sync(0xFFFFFFFFu, pk[it], o));
      }
    }
  }
  const bool want_codes = t.codes != nullptr;
  bool any_bad = false;
  Prepared<DT, KIND> prep;
  if (GIVEN) prep.setup(t.given_mx, t.given_mn, p.qc, SYM, tbl);
#pragma unroll
  for (int it = 0; it < kVPT; ++it) {
    const uint32_t idx = (uint32_t)warp * (32 * kVPT) + it * 32 + lane;
    if (!GIVEN && (it & (vpl - 1)) == 0) {  // its of one segment share the parameters
      float mx, mn;
      range_unpack<DT>(pk[it], &mx, &mn);
      prep.setup(mx, mn, p.qc, SYM, tbl);
    }
    uint2 codes8;
    const uint4 o = prep.template apply<SYM>(v[it], p.qc, tbl, want_codes, &codes8);
    if (idx < t.n_vec) {
      stg128(t.dst + (size_t)idx * 16, o);
      if (want_codes) store_codes(t.codes, t.vec0 + idx, codes8, p.qc.code_nibbles != 0);
      if (prep.bad) any_bad |= vec_has_nan<DT>(o);
      // segment parameters: one lane per segment; per-tensor (GIVEN) has a single segment per matrix
      const bool owner = GIVEN ? (t.vec0 + idx == 0) : ((lane & (lps - 1)) == 0 && (it & (vpl - 1)) == 0);
      if (owner) {
        const uint64_t g = GIVEN ? 0 : (t.vec0 + idx) / (uint32_t)(lps * vpl);
        if (t.scales) reinterpret_cast<uint16_t*>(t.scales)[g] = DType<DT>::from_f32(prep.scale);
        if (t.zeros) reinterpret_cast<uint16_t*>(t.zeros)[g] = DType<DT>::from_f32(prep.zero);
      }
    }
  }
  if (any_bad && p.status) atomicOr(p.status, IWQ_STATUS_NAN);
}

// Copy the minifloat step table (and the value table of the double-approximate decode behind it) into shared
// memory (all threads; caller syncs).
__device__ __forceinline__ const FpStep* load_fp_lut(const LaunchParams& p, uint8_t* s_lut) {
  const uint2* g = reinterpret_cast<const uint2*>(p.lut);
  uint2* s = reinterpret_cast<uint2*>(s_lut);
  const uint32_t rs = p.qc.lut_rep_shift;  // interleaved copies: entry i of copy r at [i][r]
  if (p.qc.lut_direct) {                   // direct table: lut_n 16-bit values (device buffer padded to 8 bytes)
    for (uint32_t i = threadIdx.x; i < (p.qc.lut_n * 2u + 7u) / 8u; i += blockDim.x) s[i] = g[i];
    return reinterpret_cast<const FpStep*>(s_lut);
  }
  for (uint32_t i = threadIdx.x; i < (p.qc.lut_n << rs); i += blockDim.x) s[i] = g[i >> rs];
  if (p.lut_extra_bytes) {
    const uint32_t o = p.lut_smem_bytes / 8u, n = p.lut_extra_bytes / 8u;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s[o + i] = g[o + i];
  }
  return reinterpret_cast<const FpStep*>(s_lut);
}

// ------------------------------------------------------------------------------------------------
// seg_tma_kernel: TMA bulk-copy ring -> consumer warps
// shared memory: [0,64) full barriers | [64,128) empty barriers | [128,640) TileCtx[8] | [640,1024) seg4 range slots [stage][warp] | [1024,...) LUT, stages
// ------------------------------------------------------------------------------------------------
template <int DT, int KIND, bool SYM, bool GIVEN>
__global__ void __launch_bounds__(kConsumers + 32) seg_tma_kernel(const __grid_constant__ LaunchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  TileCtx* ctx = reinterpret_cast<TileCtx*>(smem + 128);
  const FpStep* s_tbl = nullptr;
  uint8_t* data = smem + kSmemHeader + p.lut_smem_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kConsumerWarps);
    }
    fence_mbar_init();
  }
  if (KIND == KIND_FP) s_tbl = load_fp_lut(p, smem + kSmemHeader);
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ---------------- producer: one lane issues the bulk copies ----------------
    if (lane == 0) {
      int stage = 0, cur = -1;
      uint32_t phase = 0;
      Job jb;
      for (uint32_t tile = p.unit_begin + blockIdx.x; tile < p.total_units; tile += gridDim.x) {
        mbar_wait(&empty[stage], phase ^ 1u);
        advance_job(p, tile, &cur, &jb);
        const TileCtx t = make_tile_ctx(p, jb, cur, tile, GIVEN);
        ctx[stage] = t;
        const uint32_t bytes = t.n_vec * 16u;
        mbar_arrive_expect_tx(&full[stage], bytes);
        tma_load_1d(data + (size_t)stage * kTileBytes, jb.src + (t.vec0 << 4), bytes, &full[stage]);
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // ---------------- consumers ----------------
    int stage = 0;
    uint32_t phase = 0;
    for (uint32_t tile = p.unit_begin + blockIdx.x; tile < p.total_units; tile += gridDim.x) {
      mbar_wait(&full[stage], phase);
      const TileCtx t = ctx[stage];
      const uint8_t* sbase = data + (size_t)stage * kTileBytes + (size_t)warp * (32 * kVPT * 16) + lane * 16;
      uint4 v[kVPT];
#pragma unroll
      for (int it = 0; it < kVPT; ++it) {
        const uint32_t idx = (uint32_t)warp * (32 * kVPT) + it * 32 + lane;
        v[it] = (idx < t.n_vec) ? lds128(sbase + it * 512) : make_uint4(0u, 0u, 0u, 0u);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);  // data is in registers: hand the stage back early
      consume_tile<DT, KIND, SYM, GIVEN>(v, t, p, s_tbl, warp, lane);
      if (++stage == n_stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// seg4_tma_kernel (default for group >= 32 and for the per-tensor apply pass)
//
//   global --TMA bulk load--> shared stage --LDS.128--> registers --STS.128--> same stage --TMA bulk store--> global
//
// One CTA per SM: a producer lane streams 48 KB tiles (one cp.async.bulk each) into a ring of 4 stages; 96 / VPL consumer warps.
// Each lane owns VPL consecutive 16-byte vectors (VPL = 4: 32 elements, 64 bytes; 8; 16: a whole group of 128) of its warp's
// VPL x 512-byte slice of the tile, so a segment of `group` elements is held by group / (8 VPL) lanes: the min/max reduction
// needs log2 of that many shuffle rounds and the scale / zero-point arithmetic is done once per lane.
// Lane L touches its vectors in a rotated order -- (k + L/2) mod 4 at VPL 4, (k + L) mod VPL otherwise -- which makes every
// LDS.128 / STS.128 quarter-warp hit 8 distinct 16-byte bank groups (conflict-free) although lanes are 64 / 128 / 256 bytes apart.
// Results go back into the lane's own bytes of the stage; one elected lane per warp issues one bulk store for the warp's slice,
// and hands the stage back to the producer once the PREVIOUS tile's store has finished reading shared memory
// (cp.async.bulk.wait_group.read 1), so the store latency is never exposed.
// No per-element global store instructions, no per-vector address arithmetic.
// (25 warps per CTA: one SM sub-partition holds 7 of them, so the VPL = 4 form must fit 72 registers per thread.)
// ------------------------------------------------------------------------------------------------
// One lane's share of a tile in seg4_tma_kernel: VPL vectors through registers, results back to the stage.
// FAST = full tile and no code emission: no per-vector predicates at all.
template <int DT, int KIND, bool SYM, bool GIVEN, bool FAST, int LUT = LUT_STEP, int VPL = 4>
__device__ __forceinline__ void seg4_consume(const TileCtx& t, const LaunchParams& p, const FpStep* s_tbl, uint8_t* lbase, uint32_t lane_vec,
                                             uint32_t rot, int lane, int lps, int warp, uint32_t* xchg) {
  // VPL = vectors per lane (4, 8 or 16 consecutive 16-byte vectors): a lane that owns more of its segment needs fewer shuffle
  // rounds and -- what counts -- fewer lanes repeat the same scale / zero-point arithmetic (group 128: 4 lanes per segment at
  // VPL 4, one at VPL 16), and the per-tile bookkeeping is spread over more elements
  uint4 v[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const uint32_t j = (k + rot) & (uint32_t)(VPL - 1);
    v[k] = (FAST || lane_vec + j < t.n_vec) ? lds128(lbase + j * 16) : make_uint4(0u, 0u, 0u, 0u);
  }
  typename std::conditional<LUT != LUT_STEP && KIND == KIND_FP, PreparedFp<DT, LUT>, Prepared<DT, KIND>>::type prep;
  if (GIVEN) {
    prep.setup(t.given_mx, t.given_mn, p.qc, SYM, s_tbl);
  } else {
    uint32_t pk;
    if (VPL == 4) {
      pk = range_max<DT>(range_max<DT>(vec_range<DT, SYM>(v[0]), vec_range<DT, SYM>(v[1])), range_max<DT>(vec_range<DT, SYM>(v[2]), vec_range<DT, SYM>(v[3])));
    } else {
      uint32_t pr[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) pr[k] = vec_range<DT, SYM>(v[k]);
#pragma unroll
      for (int w = VPL / 2; w > 0; w >>= 1)
#pragma unroll
        for (int k = 0; k < w; ++k) pr[k] = range_max<DT>(pr[k], pr[k + w]);
      pk = pr[0];
    }
    // log2(lps) butterfly rounds; lps is launch-uniform, so this is a short uniform loop (no divergence handling)
#pragma unroll 1
    for (int o = (lps > 32 ? 32 : lps) >> 1; o > 0; o >>= 1) pk = range_max<DT>(pk, __shfl_xor_sync(0xFFFFFFFFu, pk, o));
    if (lps > 32) {
      // segment = wps consecutive warps (per-channel rows / groups of 2048, 4096 or 8192 elements): exchange the warp
      // ranges through this stage's slots; a named barrier couples just the warps of the segment
      const int wps = lps >> 5, w0 = warp & ~(wps - 1);
      if (lane == 0) xchg[warp] = pk;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + (w0 >> (__ffs(wps) - 1))), "r"(wps * 32) : "memory");
      pk = xchg[w0];
#pragma unroll 1
      for (int w = 1; w < wps; ++w) pk = range_max<DT>(pk, xchg[w0 + w]);
    }
    float mx, mn;
    range_unpack<DT>(pk, &mx, &mn);
    prep.setup(mx, mn, p.qc, SYM, s_tbl);
  }
  const bool want_codes = !FAST && t.codes != nullptr;
  bool nan_seen = false;
  if (FAST && !prep.slow()) {
    // the hot path: no code emission, reciprocal+FMA quotient; one segment-uniform branch per lane-tile keeps the
    // IEEE-division variant (degenerate scales only) out of the instruction stream
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const uint32_t j = (k + rot) & (uint32_t)(VPL - 1);
      uint2 codes8;
      sts128(lbase + j * 16, prep.template apply<SYM, 0>(v[k], p.qc, s_tbl, false, &codes8));
    }
  } else {
    // rolled loop (code size): the inputs are re-read from the stage, which still holds them (each pass overwrites
    // only its own slot), so v[] is never indexed dynamically and stays in registers on the hot path
#pragma unroll 1
    for (int k = 0; k < VPL; ++k) {
      const uint32_t j = (k + rot) & (uint32_t)(VPL - 1);
      uint2 codes8;
      const uint4 vv = (FAST || lane_vec + j < t.n_vec) ? lds128(lbase + j * 16) : make_uint4(0u, 0u, 0u, 0u);
      const uint4 o = prep.template apply<SYM>(vv, p.qc, s_tbl, want_codes, &codes8);
      sts128(lbase + j * 16, o);
      if (want_codes && lane_vec + j < t.n_vec) store_codes(t.codes, t.vec0 + lane_vec + j, codes8, p.qc.code_nibbles != 0);
      if (prep.bad && (FAST || lane_vec + j < t.n_vec)) nan_seen |= vec_has_nan<DT>(o);
    }
  }
  if (FAST || lane_vec < t.n_vec) {
    const bool owner = GIVEN ? (t.vec0 + lane_vec == 0) : (lps > 32 ? ((lane_vec & (uint32_t)(VPL * lps - 1)) == 0) : ((lane & (lps - 1)) == 0));
    if (owner) {
      // (a shift by log2(4 * lps) instead of this division was measured 0.7 % SLOWER on the same box -- A/B with two
      //  libraries, 3 runs each: 6176 vs 6130 GB/s -- so the division stays)
      const uint64_t g = GIVEN ? 0 : (t.vec0 + lane_vec) / (uint32_t)(VPL * lps);
      if (t.scales) reinterpret_cast<uint16_t*>(t.scales)[g] = DType<DT>::from_f32(prep.scale);
      if (t.zeros) reinterpret_cast<uint16_t*>(t.zeros)[g] = DType<DT>::from_f32(prep.zero);
    }
    if (nan_seen && p.status) atomicOr(p.status, IWQ_STATUS_NAN);
  }
}

template <int DT, int KIND, bool SYM, bool GIVEN, int LUT = LUT_STEP, int VPL = 4>
__global__ void __launch_bounds__(kConsumers * 4 / VPL + 32, IWQ_SEG4_MINB) seg4_tma_kernel(const __grid_constant__ LaunchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  TileCtx* ctx = reinterpret_cast<TileCtx*>(smem + 128);
  const FpStep* s_tbl = nullptr;
  uint8_t* data = smem + kSmemHeader + p.lut_smem_bytes;

  constexpr int CWv = kConsumerWarps * 4 / VPL;  // consumer warps: a warp owns VPL x 512 bytes of the tile
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], (uint32_t)CWv);
    }
    fence_mbar_init();
  }
  if (KIND == KIND_FP) s_tbl = load_fp_lut(p, smem + kSmemHeader);
  __syncthreads();

  if (warp == CWv) {
    if (lane == 0) {
      int stage = 0, cur = -1;
      uint32_t phase = 0;
      Job jb;
      const uint32_t hint = p.l2_hint & 3u;
      const uint64_t pol = hint == 1u ? l2_policy_evict_first() : (hint == 2u ? l2_policy_evict_last() : 0ull);
      const TileWalk walk = make_walk(p);
      for (uint32_t seq = walk.cur; seq < walk.end; seq += walk.step) {
        const uint32_t tile = p.reverse ? p.total_units - 1u - (seq - p.unit_begin) : seq;
        mbar_wait(&empty[stage], phase ^ 1u);
        advance_job(p, tile, &cur, &jb);
        const TileCtx t = make_tile_ctx(p, jb, cur, tile, GIVEN);
        ctx[stage] = t;
        const uint32_t bytes = t.n_vec * 16u;
        mbar_arrive_expect_tx(&full[stage], bytes);
        if (hint) tma_load_1d_hint(data + (size_t)stage * kTileBytes, jb.src + (t.vec0 << 4), bytes, &full[stage], pol);
        else tma_load_1d(data + (size_t)stage * kTileBytes, jb.src + (t.vec0 << 4), bytes, &full[stage]);
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int lps = p.lps;                      // lanes per segment = group / 32 (1..32); unused when GIVEN
  const uint32_t lane_vec = (uint32_t)warp * (32u * VPL) + (uint32_t)lane * VPL;  // first vector of this lane in the tile
  const uint32_t rot = VPL == 4 ? (uint32_t)lane >> 1 : (uint32_t)lane;       // rotated vector order: conflict-free LDS.128 / STS.128
  int stage = 0, prev_stage = -1;
  uint32_t phase = 0;
  const TileWalk walk = make_walk(p);
  for (uint32_t seq = walk.cur; seq < walk.end; seq += walk.step) {
    mbar_wait(&full[stage], phase);
    const TileCtx t = ctx[stage];
    uint8_t* lbase = data + (size_t)stage * kTileBytes + (size_t)lane_vec * 16;
    uint32_t* xchg = reinterpret_cast<uint32_t*>(smem + 640) + stage * kConsumerWarps;  // [stage][warp] range slots
    if (t.n_vec == (uint32_t)(kTileBytes / 16) && t.codes == nullptr)
      seg4_consume<DT, KIND, SYM, GIVEN, true, LUT, VPL>(t, p, s_tbl, lbase, lane_vec, rot, lane, lps, warp, xchg);
    else
      seg4_consume<DT, KIND, SYM, GIVEN, false, LUT, VPL>(t, p, s_tbl, lbase, lane_vec, rot, lane, lps, warp, xchg);
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const uint32_t w0 = (uint32_t)warp * (32u * VPL);
      if (w0 < t.n_vec) {
        const uint32_t nv = t.n_vec - w0 < 32u * VPL ? t.n_vec - w0 : 32u * VPL;
        if (p.l2_hint & (4u | 16u))  // (experiments only; the policy is built here so that it costs the hot path no registers)
          tma_store_1d_hint(t.dst + (size_t)w0 * 16, data + (size_t)stage * kTileBytes + (size_t)w0 * 16, nv * 16u, (p.l2_hint & 16u) ? l2_policy_evict_last() : l2_policy_evict_first());
        else tma_store_1d(t.dst + (size_t)w0 * 16, data + (size_t)stage * kTileBytes + (size_t)w0 * 16, nv * 16u);
      }
      tma_store_commit();
      tma_store_wait_read<1>();                       // the previous tile's store no longer reads its stage
      if (prev_stage >= 0) mbar_arrive(&empty[prev_stage]);
    }
    prev_stage = stage;
    if (++stage == n_stages) {
      stage = 0;
      phase ^= 1u;
    }
  }
  if (lane == 0) tma_store_wait_all();  // all bulk stores complete (and visible) before the CTA retires
}

// ------------------------------------------------------------------------------------------------
// seg_direct_kernel: same consumer code, plain 128-bit global loads (no shared-memory staging)
// ------------------------------------------------------------------------------------------------
template <int DT, int KIND, bool SYM, bool GIVEN>
__global__ void __launch_bounds__(kConsumers) seg_direct_kernel(const __grid_constant__ LaunchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const FpStep* s_tbl = nullptr;
  if (KIND == KIND_FP) {
    s_tbl = load_fp_lut(p, smem);
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int cur = -1;
  Job jb;
  for (uint32_t tile = p.unit_begin + blockIdx.x; tile < p.total_units; tile += gridDim.x) {
    advance_job(p, tile, &cur, &jb);
    const TileCtx t = make_tile_ctx(p, jb, cur, tile, GIVEN);
    const uint8_t* gbase = jb.src + (t.vec0 << 4);
    uint4 v[kVPT];
#pragma unroll
    for (int it = 0; it < kVPT; ++it) {
      const uint32_t idx = (uint32_t)warp * (32 * kVPT) + it * 32 + lane;
      v[it] = (idx < t.n_vec) ? ldg128(gbase + (size_t)idx * 16) : make_uint4(0u, 0u, 0u, 0u);
    }
    consume_tile<DT, KIND, SYM, GIVEN>(v, t, p, s_tbl, warp, lane);
  }
}

// ------------------------------------------------------------------------------------------------
// row_tma_kernel: one row (segment of `row_bytes`) per ring stage; blockDim.x = consumers + 32
// shared memory: header as above (partial ranges at [640, 768): 2 parities x 16 warps) | LUT | ring
// ------------------------------------------------------------------------------------------------
struct RowCtx {
  uint8_t* dst;
  uint8_t* scales;
  uint8_t* zeros;
  uint8_t* codes;
  uint64_t row;       // row index within the matrix
  uint32_t n_vec;
  uint32_t pad;
};

template <int DT, int KIND, bool SYM>
__global__ void __launch_bounds__(512 + 32) row_tma_kernel(const __grid_constant__ LaunchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  RowCtx* ctx = reinterpret_cast<RowCtx*>(smem + 128);        // 8 x 48 B = 384 B
  const FpStep* s_tbl = nullptr;
  uint32_t* partial = reinterpret_cast<uint32_t*>(smem + 640); // [2][16] per-warp partial ranges
  uint8_t* data = smem + kSmemHeader + p.lut_smem_bytes;

  const int n_cons = (int)blockDim.x - 32;
  const int n_cwarps = n_cons >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], n_cwarps);
    }
    fence_mbar_init();
  }
  if (KIND == KIND_FP) s_tbl = load_fp_lut(p, smem + kSmemHeader);
  __syncthreads();

  if (warp == n_cwarps) {
    if (lane == 0) {
      int stage = 0, cur = -1;
      uint32_t phase = 0;
      Job jb;
      for (uint32_t row = blockIdx.x; row < p.total_units; row += gridDim.x) {
        mbar_wait(&empty[stage], phase ^ 1u);
        advance_job(p, row, &cur, &jb);
        const uint64_t r = row - jb.first_unit;
        RowCtx t;
        t.dst = jb.dst + r * jb.row_bytes;
        t.scales = jb.scales;
        t.zeros = jb.zeros;
        t.codes = jb.codes ? jb.codes + ((r * (jb.row_bytes >> 1)) >> (p.qc.code_nibbles ? 1 : 0)) : nullptr;
        t.row = r;
        t.n_vec = jb.row_bytes >> 4;
        t.pad = 0;
        ctx[stage] = t;
        mbar_arrive_expect_tx(&full[stage], jb.row_bytes);
        tma_load_1d(data + (size_t)stage * p.stage_bytes, jb.src + r * jb.row_bytes, jb.row_bytes, &full[stage]);
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    int stage = 0, prev_stage = -1;  // prev_stage: stage whose bulk store may still be reading shared memory (TMA-out form)
    uint32_t phase = 0, par = 0;
    const int tid = threadIdx.x;
    const bool tma_out = p.row_tma_out != 0;
    for (uint32_t row = blockIdx.x; row < p.total_units; row += gridDim.x) {
      mbar_wait(&full[stage], phase);
      const RowCtx t = ctx[stage];
      uint8_t* sdata = data + (size_t)stage * p.stage_bytes;
      // sweep 1: range of the row
      uint32_t pk = SYM ? 0u : 0xFC00FC00u;  // {0,0} for absmax; {-inf,-inf} otherwise (bf16: see below)
      if (!SYM && DT == IWQ_BF16) pk = 0xFF80FF80u;
      for (uint32_t i = tid; i < t.n_vec; i += n_cons) pk = range_max<DT>(pk, vec_range<DT, SYM>(lds128(sdata + (size_t)i * 16)));
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) pk = range_max<DT>(pk, __shfl_xor_sync(0xFFFFFFFFu, pk, o));
      if (lane == 0) partial[par * 16 + warp] = pk;
      asm volatile("bar.sync 1, %0;" ::"r"(n_cons) : "memory");
      uint32_t all = partial[par * 16];
      for (int w = 1; w < n_cwarps; ++w) all = range_max<DT>(all, partial[par * 16 + w]);
      par ^= 1u;
      float mx, mn;
      range_unpack<DT>(all, &mx, &mn);
      // sweep 2: quantise from shared memory, store to global
      const bool want_codes = t.codes != nullptr;
      Prepared<DT, KIND> prep;
      prep.setup(mx, mn, p.qc, SYM, s_tbl);
      bool stored_by_tma = false;
      if (tma_out && !want_codes && !prep.slow()) {
        // hot path, TMA-out form (ring of >= 3 stages): each warp quantises a contiguous slice of the row in place in
        // the stage and one lane sends it off with a single bulk store -- no per-vector global stores or addresses;
        // the stage is handed back once the PREVIOUS row's store has finished reading shared memory (as in seg4)
        const uint32_t slice = (t.n_vec + (uint32_t)n_cwarps * 32u - 1u) / ((uint32_t)n_cwarps * 32u) * 32u;
        const uint32_t v0 = (uint32_t)warp * slice, v1 = v0 + slice < t.n_vec ? v0 + slice : t.n_vec;
        for (uint32_t i = v0 + (uint32_t)lane; i < v1; i += 32u) {
          uint2 codes8;
          sts128(sdata + (size_t)i * 16, prep.template apply<SYM, 0>(lds128(sdata + (size_t)i * 16), p.qc, s_tbl, false, &codes8));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (v0 < v1) tma_store_1d(t.dst + (size_t)v0 * 16, sdata + (size_t)v0 * 16, (v1 - v0) * 16u);
          tma_store_commit();
          tma_store_wait_read<1>();
          if (prev_stage >= 0) mbar_arrive(&empty[prev_stage]);
        }
        stored_by_tma = true;
      } else if (!want_codes && !prep.slow()) {
        // hot path: the degenerate-scale (IEEE division) variant is selected once per row, not once per pair
        for (uint32_t i = tid; i < t.n_vec; i += n_cons) {
          uint2 codes8;
          stg128(t.dst + (size_t)i * 16, prep.template apply<SYM, 0>(lds128(sdata + (size_t)i * 16), p.qc, s_tbl, false, &codes8));
        }
      } else {
        for (uint32_t i = tid; i < t.n_vec; i += n_cons) {
          uint2 codes8;
          const uint4 o = prep.template apply<SYM>(lds128(sdata + (size_t)i * 16), p.qc, s_tbl, want_codes, &codes8);
          stg128(t.dst + (size_t)i * 16, o);
          if (want_codes) store_codes(t.codes, i, codes8, p.qc.code_nibbles != 0);
          if (prep.bad && vec_has_nan<DT>(o) && p.status) atomicOr(p.status, IWQ_STATUS_NAN);  // suspect rows only
        }
      }
      if (tid == 0) {
        if (t.scales) reinterpret_cast<uint16_t*>(t.scales)[t.row] = DType<DT>::from_f32(prep.scale);
        if (t.zeros) reinterpret_cast<uint16_t*>(t.zeros)[t.row] = DType<DT>::from_f32(prep.zero);
      }
      __syncwarp();
      if (stored_by_tma) {
        prev_stage = stage;
      } else {
        if (lane == 0) {
          if (prev_stage >= 0) {  // a rare register-store row after TMA-out rows: settle the parked stage first
            tma_store_wait_read<0>();
            mbar_arrive(&empty[prev_stage]);
          }
          mbar_arrive(&empty[stage]);
        }
        prev_stage = -1;
      }
      if (++stage == n_stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (tma_out && lane == 0) tma_store_wait_all();
  }
}

// ------------------------------------------------------------------------------------------------
// quant_dim = 1: segments run DOWN the rows (reference: transpose, group along the last dim, transpose back;
// quant_linear.py:478-483, :640-647).  No transpose here: columns are independent, so the whole path is
// vectorised ACROSS columns.
//
//   tile = `g` rows x 256 columns.  A lane owns 8 adjacent columns (one 16-byte vector per row: a warp reads
//   512 contiguous bytes of a row), the 8 warps take rows k = warp, warp+8, ...  The range of a column is a
//   plain element-wise HMNMX2 over the rows a warp visits (no shuffles at all), combined across the warps
//   through shared memory.  Every lane then derives the parameters of its 8 columns; because the two halves of
//   a packed register now belong to different segments, the element stage uses the PAIR forms (per-half scale
//   and reciprocal).  The second sweep re-reads the tile (64 KB for g = 128: served by L2) and stores in place.
// Units of the work list: tiles, column tile fastest.  blockDim.x = 256.
// ------------------------------------------------------------------------------------------------
template <int DT, int KIND>
struct PreparedPair;

template <int DT, int KIND>
struct PreparedPairInt {
  IntFast<DT, KIND == KIND_INT_WIDE, true> f;
  __device__ __forceinline__ void setup2(float mxa, float mna, float mxb, float mnb, const QuantConsts& c, bool sym, const FpStep*, float* sa, float* za, float* sb,
                                         float* zb, bool* bad) {
    IntQ qa = int_segment_params<DT>(mxa, mna, c, sym), qb = int_segment_params<DT>(mxb, mnb, c, sym);
    f.setup2(qa, qb);
    *sa = qa.s; *za = qa.z; *sb = qb.s; *zb = qb.z;
    *bad = qa.bad || qb.bad;
  }
  __device__ __forceinline__ bool slow() const { return f.slow_div; }
  // take the per-segment constants from lane `src` (which ran setup2 for this column pair)
  __device__ __forceinline__ void shfl_from(const PreparedPairInt& o, int src) {
    using T2 = typename DType<DT>::T2;
    f.lo2 = from_u32<T2>(__shfl_sync(0xFFFFFFFFu, as_u32(o.f.lo2), src));
    f.hi2 = from_u32<T2>(__shfl_sync(0xFFFFFFFFu, as_u32(o.f.hi2), src));
    f.s2 = from_u32<T2>(__shfl_sync(0xFFFFFFFFu, as_u32(o.f.s2), src));
    f.s.x = __shfl_sync(0xFFFFFFFFu, o.f.s.x, src);
    f.s.y = __shfl_sync(0xFFFFFFFFu, o.f.s.y, src);
    f.r.x = __shfl_sync(0xFFFFFFFFu, o.f.r.x, src);
    f.r.y = __shfl_sync(0xFFFFFFFFu, o.f.r.y, src);
    f.slow_div = __shfl_sync(0xFFFFFFFFu, (int)o.f.slow_div, src) != 0;
  }
  // generic field exchange (col_wide_kernel hands the constants round through shared memory): x maps every per-pair word
  static constexpr int kWords = 8;
  template <class X>
  __device__ __forceinline__ void map_from(const PreparedPairInt& o, X x) {
    using T2 = typename DType<DT>::T2;
    f.lo2 = from_u32<T2>(x(as_u32(o.f.lo2)));
    f.hi2 = from_u32<T2>(x(as_u32(o.f.hi2)));
    f.s2 = from_u32<T2>(x(as_u32(o.f.s2)));
    f.s.x = __uint_as_float(x(__float_as_uint(o.f.s.x)));
    f.s.y = __uint_as_float(x(__float_as_uint(o.f.s.y)));
    f.r.x = __uint_as_float(x(__float_as_uint(o.f.r.x)));
    f.r.y = __uint_as_float(x(__float_as_uint(o.f.r.y)));
    f.slow_div = x((uint32_t)o.f.slow_div) != 0u;
  }
  template <bool SYM, int SLOW = -1>
  __device__ __forceinline__ uint32_t apply(uint32_t w2, const QuantConsts&) const {
    typename DType<DT>::T2 d;
    return f.template apply<SLOW>(w2, &d);
  }
};
template <int DT>
struct PreparedPair<DT, KIND_INT_FAST> : PreparedPairInt<DT, KIND_INT_FAST> {};
template <int DT>
struct PreparedPair<DT, KIND_INT_WIDE> : PreparedPairInt<DT, KIND_INT_WIDE> {};

template <int DT>
struct PreparedPair<DT, KIND_INT_GENERIC> {
  IntQ qa, qb;
  __device__ __forceinline__ void setup2(float mxa, float mna, float mxb, float mnb, const QuantConsts& c, bool sym, const FpStep*, float* sa, float* za, float* sb,
                                         float* zb, bool* bad) {
    qa = int_segment_params<DT>(mxa, mna, c, sym);
    qb = int_segment_params<DT>(mxb, mnb, c, sym);
    *sa = qa.s; *za = qa.z; *sb = qb.s; *zb = qb.z;
    *bad = qa.bad || qb.bad;
  }
  __device__ __forceinline__ bool slow() const { return false; }
  __device__ __forceinline__ void shfl_from(const PreparedPair& o, int src) {
    auto one = [&](IntQ& d, const IntQ& q) {
      d.s = __shfl_sync(0xFFFFFFFFu, q.s, src);
      d.r = __shfl_sync(0xFFFFFFFFu, q.r, src);
      d.z = __shfl_sync(0xFFFFFFFFu, q.z, src);
      d.lo = __shfl_sync(0xFFFFFFFFu, q.lo, src);
      d.hi = __shfl_sync(0xFFFFFFFFu, q.hi, src);
      d.bad = __shfl_sync(0xFFFFFFFFu, (int)q.bad, src) != 0;
      d.slow_div = __shfl_sync(0xFFFFFFFFu, (int)q.slow_div, src) != 0;
    };
    one(qa, o.qa);
    one(qb, o.qb);
  }
  static constexpr int kWords = 16;
  template <class X>
  __device__ __forceinline__ void map_from(const PreparedPair& o, X x) {
    auto one = [&](IntQ& d, const IntQ& q) {
      d.s = __uint_as_float(x(__float_as_uint(q.s)));
      d.r = __uint_as_float(x(__float_as_uint(q.r)));
      d.z = __uint_as_float(x(__float_as_uint(q.z)));
      d.lo = __uint_as_float(x(__float_as_uint(q.lo)));
      d.hi = __uint_as_float(x(__float_as_uint(q.hi)));
      d.bad = x((uint32_t)q.bad) != 0u;
      d.slow_div = x((uint32_t)q.slow_div) != 0u;
    };
    one(qa, o.qa);
    one(qb, o.qb);
  }
  template <bool SYM, int SLOW = -1>
  __device__ __forceinline__ uint32_t apply(uint32_t w2, const QuantConsts& c) const {
    float k;
    const uint32_t lo = int_generic_elem<DT>((uint16_t)(w2 & 0xFFFFu), qa, c, &k);
    const uint32_t hi = int_generic_elem<DT>((uint16_t)(w2 >> 16), qb, c, &k);
    return lo | (hi << 16);
  }
};

template <int DT, int LUT>
struct PreparedPairFp {
  FpTable<DT, true, LUT> f;
  __device__ __forceinline__ void setup2(float mxa, float mna, float mxb, float mnb, const QuantConsts& c, bool, const FpStep* steps, float* sa, float* za, float* sb,
                                         float* zb, bool* bad) {
    FpQ qa = fp_segment_params<DT>(mxa, mna, c), qb = fp_segment_params<DT>(mxb, mnb, c);
    f.setup2(qa, qb, c, steps);
    *sa = qa.S; *za = qa.mid; *sb = qb.S; *zb = qb.mid;
    *bad = qa.bad || qb.bad;
  }
  __device__ __forceinline__ bool slow() const { return f.slow_div; }
  __device__ __forceinline__ void shfl_from(const PreparedPairFp& o, int src) {
    using T2 = typename DType<DT>::T2;
    f = o.f;  // launch-uniform members (clamp bounds, table address, flags)
    f.S2 = from_u32<T2>(__shfl_sync(0xFFFFFFFFu, as_u32(o.f.S2), src));
    f.mid2 = from_u32<T2>(__shfl_sync(0xFFFFFFFFu, as_u32(o.f.mid2), src));
    f.midh2 = from_u32<T2>(__shfl_sync(0xFFFFFFFFu, as_u32(o.f.midh2), src));
    f.S.x = __shfl_sync(0xFFFFFFFFu, o.f.S.x, src);
    f.S.y = __shfl_sync(0xFFFFFFFFu, o.f.S.y, src);
    f.r.x = __shfl_sync(0xFFFFFFFFu, o.f.r.x, src);
    f.r.y = __shfl_sync(0xFFFFFFFFu, o.f.r.y, src);
    f.slow_div = __shfl_sync(0xFFFFFFFFu, (int)o.f.slow_div, src) != 0;
  }
  static constexpr int kWords = 8;
  template <class X>
  __device__ __forceinline__ void map_from(const PreparedPairFp& o, X x) {
    using T2 = typename DType<DT>::T2;
    f = o.f;  // launch-uniform members
    f.S2 = from_u32<T2>(x(as_u32(o.f.S2)));
    f.mid2 = from_u32<T2>(x(as_u32(o.f.mid2)));
    f.midh2 = from_u32<T2>(x(as_u32(o.f.midh2)));
    f.S.x = __uint_as_float(x(__float_as_uint(o.f.S.x)));
    f.S.y = __uint_as_float(x(__float_as_uint(o.f.S.y)));
    f.r.x = __uint_as_float(x(__float_as_uint(o.f.r.x)));
    f.r.y = __uint_as_float(x(__float_as_uint(o.f.r.y)));
    f.slow_div = x((uint32_t)o.f.slow_div) != 0u;
  }
  template <bool SYM, int SLOW = -1>
  __device__ __forceinline__ uint32_t apply(uint32_t w2, const QuantConsts&) const {
    typename DType<DT>::T2 v;
    if (SLOW == 0) return f.template apply<false, SYM ? 0 : 1>(w2, &v);
    return f.slow_div ? f.template apply<true, SYM ? 0 : 1>(w2, &v) : f.template apply<false, SYM ? 0 : 1>(w2, &v);
  }
};
template <int DT>
struct PreparedPair<DT, KIND_FP> : PreparedPairFp<DT, LUT_STEP> {};

template <int DT>
struct PreparedPair<DT, KIND_BFP> {
  BfpApply<DT, true> f;
  __device__ __forceinline__ void setup2(float mxa, float, float mxb, float, const QuantConsts& c, bool, const FpStep*, float* sa, float* za, float* sb, float* zb,
                                         bool* bad) {
    f.setup2(bfp_segment_params<DT>(mxa, c), bfp_segment_params<DT>(mxb, c), c);
    *sa = 0.f; *za = 0.f; *sb = 0.f; *zb = 0.f;
    *bad = false;
  }
  __device__ __forceinline__ bool slow() const { return false; }
  __device__ __forceinline__ void shfl_from(const PreparedPair& o, int src) {
    f = o.f;
    f.step.x = __shfl_sync(0xFFFFFFFFu, o.f.step.x, src);
    f.step.y = __shfl_sync(0xFFFFFFFFu, o.f.step.y, src);
    f.inv_step.x = __shfl_sync(0xFFFFFFFFu, o.f.inv_step.x, src);
    f.inv_step.y = __shfl_sync(0xFFFFFFFFu, o.f.inv_step.y, src);
  }
  static constexpr int kWords = 4;
  template <class X>
  __device__ __forceinline__ void map_from(const PreparedPair& o, X x) {
    f = o.f;
    f.step.x = __uint_as_float(x(__float_as_uint(o.f.step.x)));
    f.step.y = __uint_as_float(x(__float_as_uint(o.f.step.y)));
    f.inv_step.x = __uint_as_float(x(__float_as_uint(o.f.inv_step.x)));
    f.inv_step.y = __uint_as_float(x(__float_as_uint(o.f.inv_step.y)));
  }
  template <bool SYM, int SLOW = -1>
  __device__ __forceinline__ uint32_t apply(uint32_t w2, const QuantConsts&) const { return f.apply(w2); }
};

constexpr int kColTile = 256;  // columns per tile: 32 lanes x 8

template <int DT, int KIND, bool SYM>
__global__ void __launch_bounds__(256) col_kernel(const __grid_constant__ LaunchParams p) {
  using T2 = typename DType<DT>::T2;
  extern __shared__ __align__(1024) uint8_t smem[];
  // [2 parities][8 warps][2 (max | min)][32 lanes] uint4 = 16 KB, then the minifloat step table
  uint4* part = reinterpret_cast<uint4*>(smem);
  const FpStep* s_tbl = nullptr;
  if (KIND == KIND_FP) s_tbl = load_fp_lut(p, smem + 16384);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ident_max = SYM ? 0u : (DT == IWQ_F16 ? 0xFC00FC00u : 0xFF80FF80u);  // {0,0} | {-inf,-inf}
  const uint32_t ident_min = DT == IWQ_F16 ? 0x7C007C00u : 0x7F807F80u;                // {+inf,+inf}
  int cur = -1;
  Job jb;
  uint32_t par = 0;
  for (uint32_t tile = blockIdx.x; tile < p.total_units; tile += gridDim.x) {
    advance_job(p, tile, &cur, &jb);
    const uint32_t cols = jb.row_bytes >> 1;
    const uint32_t rows = (uint32_t)(jb.bytes / jb.row_bytes);
    const uint32_t g = p.col_group ? p.col_group : rows;
    const uint32_t n_ct = (cols + kColTile - 1) / kColTile;
    const uint32_t local = tile - jb.first_unit;
    const uint32_t rb = local / n_ct, ct = local - rb * n_ct;
    const uint32_t col = ct * kColTile + (uint32_t)lane * 8u;
    const bool live = col < cols;  // cols is a multiple of 8 (checked on the host)
    const size_t off = (size_t)rb * g * jb.row_bytes + (size_t)col * 2;
    const uint8_t* src = jb.src + off;
    uint8_t* dst = jb.dst + off;

    // ---- sweep 1: per-column range over the rows this warp visits ----
    uint32_t mx[4] = {ident_max, ident_max, ident_max, ident_max};
    uint32_t mn[4] = {ident_min, ident_min, ident_min, ident_min};
    if (live) {
      uint32_t k = (uint32_t)warp;
      for (; k + 24 < g; k += 32) {  // 4 rows in flight
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg128(src + (size_t)(k + 8 * u) * jb.row_bytes);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (SYM) {
              mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(w4[j] & 0x7FFF7FFFu)));
            } else {
              mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(w4[j])));
              mn[j] = as_u32(__hmin2_nan(from_u32<T2>(mn[j]), from_u32<T2>(w4[j])));
            }
          }
        }
      }
      for (; k < g; k += 8) {
        const uint4 v = ldg128(src + (size_t)k * jb.row_bytes);
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (SYM) {
            mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(w4[j] & 0x7FFF7FFFu)));
          } else {
            mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(w4[j])));
            mn[j] = as_u32(__hmin2_nan(from_u32<T2>(mn[j]), from_u32<T2>(w4[j])));
          }
        }
      }
    }
    uint4* mine = part + ((par * 8 + warp) * 2) * 32 + lane;
    mine[0] = make_uint4(mx[0], mx[1], mx[2], mx[3]);
    if (!SYM) mine[32] = make_uint4(mn[0], mn[1], mn[2], mn[3]);
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const uint4* o = part + ((par * 8 + w) * 2) * 32 + lane;
      const uint4 a = o[0];
      mx[0] = as_u32(__hmax2_nan(from_u32<T2>(mx[0]), from_u32<T2>(a.x)));
      mx[1] = as_u32(__hmax2_nan(from_u32<T2>(mx[1]), from_u32<T2>(a.y)));
      mx[2] = as_u32(__hmax2_nan(from_u32<T2>(mx[2]), from_u32<T2>(a.z)));
      mx[3] = as_u32(__hmax2_nan(from_u32<T2>(mx[3]), from_u32<T2>(a.w)));
      if (!SYM) {
        const uint4 b = o[32];
        mn[0] = as_u32(__hmin2_nan(from_u32<T2>(mn[0]), from_u32<T2>(b.x)));
        mn[1] = as_u32(__hmin2_nan(from_u32<T2>(mn[1]), from_u32<T2>(b.y)));
        mn[2] = as_u32(__hmin2_nan(from_u32<T2>(mn[2]), from_u32<T2>(b.z)));
        mn[3] = as_u32(__hmin2_nan(from_u32<T2>(mn[3]), from_u32<T2>(b.w)));
      }
    }
    par ^= 1u;  // the next tile writes the other half: one barrier per tile is enough

    // ---- per-column parameters: 4 packed pairs per lane ----
    PreparedPair<DT, KIND> prep[4];
    bool any_bad = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mxa = DType<DT>::to_f32((uint16_t)(mx[j] & 0xFFFFu)), mxb = DType<DT>::to_f32((uint16_t)(mx[j] >> 16));
      const float mna = DType<DT>::to_f32((uint16_t)(mn[j] & 0xFFFFu)), mnb = DType<DT>::to_f32((uint16_t)(mn[j] >> 16));
      float sa, za, sb, zb;
      bool bad;
      prep[j].setup2(mxa, SYM ? -mxa : mna, mxb, SYM ? -mxb : mnb, p.qc, SYM, s_tbl, &sa, &za, &sb, &zb, &bad);
      any_bad |= bad;
      if (live && warp == 0) {
        // segment order of the reference for quant_dim = 1: column-major over (column, row block)
        const size_t n_rb = rows / g;
        const size_t ga = (size_t)(col + 2 * j) * n_rb + rb, gb = ga + n_rb;
        if (jb.scales) {
          reinterpret_cast<uint16_t*>(jb.scales)[ga] = DType<DT>::from_f32(sa);
          reinterpret_cast<uint16_t*>(jb.scales)[gb] = DType<DT>::from_f32(sb);
        }
        if (jb.zeros) {
          reinterpret_cast<uint16_t*>(jb.zeros)[ga] = DType<DT>::from_f32(za);
          reinterpret_cast<uint16_t*>(jb.zeros)[gb] = DType<DT>::from_f32(zb);
        }
      }
    }
    bool nan_seen = false;

    // ---- sweep 2: quantise (tile re-read from L2), store in place ----
    if (live) {
      uint32_t k = (uint32_t)warp;
      for (; k + 8 < g && !any_bad; k += 16) {  // 2 rows in flight (suspect tiles take the one-row loop, which inspects its outputs)
        const uint4 v0 = ldg128(src + (size_t)k * jb.row_bytes);
        const uint4 v1 = ldg128(src + (size_t)(k + 8) * jb.row_bytes);
        uint4 o0, o1;
        o0.x = prep[0].template apply<SYM>(v0.x, p.qc); o0.y = prep[1].template apply<SYM>(v0.y, p.qc); o0.z = prep[2].template apply<SYM>(v0.z, p.qc); o0.w = prep[3].template apply<SYM>(v0.w, p.qc);
        o1.x = prep[0].template apply<SYM>(v1.x, p.qc); o1.y = prep[1].template apply<SYM>(v1.y, p.qc); o1.z = prep[2].template apply<SYM>(v1.z, p.qc); o1.w = prep[3].template apply<SYM>(v1.w, p.qc);
        stg128(dst + (size_t)k * jb.row_bytes, o0);
        stg128(dst + (size_t)(k + 8) * jb.row_bytes, o1);
      }
      for (; k < g; k += 8) {
        const uint4 v0 = ldg128(src + (size_t)k * jb.row_bytes);
        uint4 o0;
        o0.x = prep[0].template apply<SYM>(v0.x, p.qc); o0.y = prep[1].template apply<SYM>(v0.y, p.qc); o0.z = prep[2].template apply<SYM>(v0.z, p.qc); o0.w = prep[3].template apply<SYM>(v0.w, p.qc);
        stg128(dst + (size_t)k * jb.row_bytes, o0);
        if (any_bad) nan_seen |= vec_has_nan<DT>(o0);
      }
    }
    if (nan_seen && p.status) atomicOr(p.status, IWQ_STATUS_NAN);
  }
}

// ------------------------------------------------------------------------------------------------
// col_reg_kernel (default for quant_dim = 1 with group 32 .. 512): the whole segment lives in REGISTERS.
//
// A tile is `g` rows x C columns (C = 256 / 128 / 64).  Warp w owns the strip of C/8 columns [w*C/8, ...) for
// all rows, so a segment never leaves the warp: no shared memory, no barrier.  Inside the warp lpr = C/64 lanes
// cover one row of the strip (one 16-byte vector = 8 columns per lane) and the 32/lpr lane groups take the rows
// r, r + 32/lpr, ...: RPL = g * lpr / 32 vectors per lane, all loaded up front (RPL 128-bit loads in flight per
// lane: 8 KB per warp at RPL = 16, which is what hides the HBM latency), reduced element-wise (HMNMX2) and with
// log2(32/lpr) butterfly rounds, quantised in registers and stored.  Every byte crosses HBM exactly once each way.
// (A 1-D TMA staged variant was measured first: one bulk copy per 256..512-byte row piece makes the kernel
// TMA-issue-bound at ~3.7 TB/s, see profiles/.)
// ------------------------------------------------------------------------------------------------
template <int DT, bool SYM>
__device__ __forceinline__ void col_range_update(uint32_t (&mx)[4], uint32_t (&mn)[4], const uint4& v) {
  using T2 = typename DType<DT>::T2;
  const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (SYM) {
      mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), __habs2(from_u32<T2>(w4[j]))));
    } else {
      mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(w4[j])));
      mn[j] = as_u32(__hmin2_nan(from_u32<T2>(mn[j]), from_u32<T2>(w4[j])));
    }
  }
}

template <int DT, int KIND, bool SYM, int RPL>
__global__ void __launch_bounds__(256, 2) col_reg_kernel(const __grid_constant__ LaunchParams p) {
  using T2 = typename DType<DT>::T2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const FpStep* s_tbl = nullptr;
  if (KIND == KIND_FP) {
    s_tbl = load_fp_lut(p, smem);
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t C = p.col_tile, g = p.col_group;
  const uint32_t lpr = C >> 6;                 // lanes per strip row: 4 / 2 / 1
  const uint32_t rps = 32u / lpr;              // rows per warp-wide access
  const uint32_t sub = (uint32_t)lane & (lpr - 1u), rph = (uint32_t)lane / lpr;
  const uint32_t tcol = (uint32_t)warp * (C >> 3) + sub * 8u;  // this lane's first column inside the tile
  const uint32_t ident_max = SYM ? 0u : (DT == IWQ_F16 ? 0xFC00FC00u : 0xFF80FF80u);  // {0,0} | {-inf,-inf}
  const uint32_t ident_min = DT == IWQ_F16 ? 0x7C007C00u : 0x7F807F80u;                // {+inf,+inf}
  int cur = -1;
  Job jb;
  for (uint32_t tile = blockIdx.x; tile < p.total_units; tile += gridDim.x) {
    advance_job(p, tile, &cur, &jb);
    const uint32_t cols = jb.row_bytes >> 1;
    const uint32_t rows = (uint32_t)(jb.bytes / jb.row_bytes);
    const uint32_t n_ct = (cols + C - 1) / C;
    const uint32_t local = tile - jb.first_unit;
    const uint32_t rb = local / n_ct, ct = local - rb * n_ct;
    const uint32_t col = ct * C + tcol;
    const bool live = col < cols;  // cols is a multiple of 8 (checked on the host)
    const size_t off = ((size_t)rb * g + rph) * jb.row_bytes + (size_t)col * 2;
    const size_t step = (size_t)rps * jb.row_bytes;
    const uint8_t* src = jb.src + off;
    uint8_t* dst = jb.dst + off;

    uint4 v[RPL];
#pragma unroll
    for (int i = 0; i < RPL; ++i) v[i] = live ? ldg128(src + (size_t)i * step) : make_uint4(0u, 0u, 0u, 0u);

    // ---- per-column range: this lane's rows, then across the lane groups ----
    uint32_t mx[4] = {ident_max, ident_max, ident_max, ident_max};
    uint32_t mn[4] = {ident_min, ident_min, ident_min, ident_min};
#pragma unroll
    for (int i = 0; i < RPL; ++i) col_range_update<DT, SYM>(mx, mn, v[i]);
#pragma unroll 1
    for (uint32_t o = lpr; o < 32u; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mx[j], o))));
        if (!SYM) mn[j] = as_u32(__hmin2_nan(from_u32<T2>(mn[j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mn[j], o))));
      }
    }

    // ---- per-column parameters: 4 packed pairs per lane ----
    PreparedPair<DT, KIND> prep[4];
    bool any_bad = false;
    const size_t n_rb = rows / g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mxa = DType<DT>::to_f32((uint16_t)(mx[j] & 0xFFFFu)), mxb = DType<DT>::to_f32((uint16_t)(mx[j] >> 16));
      const float mna = DType<DT>::to_f32((uint16_t)(mn[j] & 0xFFFFu)), mnb = DType<DT>::to_f32((uint16_t)(mn[j] >> 16));
      float sa, za, sb, zb;
      bool bad;
      prep[j].setup2(mxa, SYM ? -mxa : mna, mxb, SYM ? -mxb : mnb, p.qc, SYM, s_tbl, &sa, &za, &sb, &zb, &bad);
      any_bad |= bad;
      if (live && rph == 0) {
        // segment order of the reference for quant_dim = 1: column-major over (column, row block)
        const size_t ga = (size_t)(col + 2 * j) * n_rb + rb, gb = ga + n_rb;
        if (jb.scales) {
          reinterpret_cast<uint16_t*>(jb.scales)[ga] = DType<DT>::from_f32(sa);
          reinterpret_cast<uint16_t*>(jb.scales)[gb] = DType<DT>::from_f32(sb);
        }
        if (jb.zeros) {
          reinterpret_cast<uint16_t*>(jb.zeros)[ga] = DType<DT>::from_f32(za);
          reinterpret_cast<uint16_t*>(jb.zeros)[gb] = DType<DT>::from_f32(zb);
        }
      }
    }
    // ---- element stage on the registers, store in place ----
    bool nan_seen = false;
#pragma unroll
    for (int i = 0; i < RPL; ++i) {
      uint4 o;
      o.x = prep[0].template apply<SYM>(v[i].x, p.qc);
      o.y = prep[1].template apply<SYM>(v[i].y, p.qc);
      o.z = prep[2].template apply<SYM>(v[i].z, p.qc);
      o.w = prep[3].template apply<SYM>(v[i].w, p.qc);
      if (live) stg128(dst + (size_t)i * step, o);
      if (live && any_bad) nan_seen |= vec_has_nan<DT>(o);
    }
    if (nan_seen && p.status) atomicOr(p.status, IWQ_STATUS_NAN);
  }
}

// ------------------------------------------------------------------------------------------------
// col_async_kernel (default for quant_dim = 1 with group 32 .. 512): like col_reg_kernel a warp owns a column
// strip (16*lpr bytes per row) of the tile for all `g` rows, but the strip is parked in a WARP-PRIVATE
// shared-memory ring filled with 16-byte cp.async copies issued by the warp itself, which frees the 64 data
// registers of col_reg_kernel: twice the resident warps, and with n_stages >= 2 the copies of the warp's next
// tile are in flight while it reduces and quantises the current one.  Nothing but __syncwarp() couples the lanes:
// no barrier, no mbarrier.  The strip is stored densely ([row][16*lpr bytes]), so a warp-wide LDS.128 reads 512
// contiguous bytes.  blockDim.x = 128 (4 warps): tile = g rows x 32*lpr columns; shared memory = LUT + 4 warps x
// n_stages x g*16*lpr bytes.
// (Measured alternatives, profiles/r01_col_kernel_notes.md: 1-D TMA row pieces are TMA-issue-bound; a CTA-wide
// row-coalesced cp.async ring with two barriers per tile leaves only 8 warps per SM and is compute-latency-bound.)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_dyn(int n) {  // n = groups that may stay in flight (0..3)
  if (n <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
  else if (n == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
  else if (n == 2) asm volatile("cp.async.wait_group 2;" ::: "memory");
  else asm volatile("cp.async.wait_group 3;" ::: "memory");
}

struct ColTile {        // per-lane view of one tile
  uint8_t* dst;         // this lane's first output vector (row rph of the row block, its 8 columns)
  uint8_t* scales;
  uint8_t* zeros;
  size_t step;          // bytes between the lane's consecutive rows = rps * row_bytes
  size_t seg0;          // segment index of the lane's first column: col * n_rb + rb
  uint32_t n_rb;
  bool live;
};

constexpr int kColAsyncWarps = 4;

template <int DT, int KIND, bool SYM>
__global__ void __launch_bounds__(kColAsyncWarps * 32) col_async_kernel(const __grid_constant__ LaunchParams p) {
  using T2 = typename DType<DT>::T2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const FpStep* s_tbl = nullptr;
  if (KIND == KIND_FP) {
    s_tbl = load_fp_lut(p, smem);
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t C = p.col_tile, g = p.col_group;
  const int n_stages = p.n_stages;
  const uint32_t lpr = C >> 5;                 // lanes per strip row: 4 / 2 / 1   (C = 32 * lpr columns per 4-warp tile)
  const uint32_t rps = 32u / lpr;              // rows per warp-wide access
  const uint32_t sub = (uint32_t)lane & (lpr - 1u), rph = (uint32_t)lane / lpr;
  const uint32_t tcol = (uint32_t)warp * (8u * lpr) + sub * 8u;  // this lane's first column inside the tile
  const uint32_t n_it = g / rps;               // vectors per lane per tile
  const uint32_t stage_bytes = p.stage_bytes;  // g * 16 * lpr
  uint8_t* ring = smem + p.lut_smem_bytes + (size_t)warp * n_stages * stage_bytes + (size_t)lane * 16u;  // + it * 512
  const uint32_t ident_max = SYM ? 0u : (DT == IWQ_F16 ? 0xFC00FC00u : 0xFF80FF80u);  // {0,0} | {-inf,-inf}
  const uint32_t ident_min = DT == IWQ_F16 ? 0x7C007C00u : 0x7F807F80u;                // {+inf,+inf}
  int cur_l = -1, cur_c = -1;  // job cursors: copy issue runs n_stages - 1 tiles ahead of the arithmetic
  Job jb_l, jb_c;

  auto describe = [&](uint32_t tile, int* cur, Job* jb, const uint8_t** src) -> ColTile {
    ColTile t;
    advance_job(p, tile, cur, jb);
    const uint32_t cols = jb->row_bytes >> 1;
    const uint32_t rows = (uint32_t)(jb->bytes / jb->row_bytes);
    const uint32_t n_ct = (cols + C - 1) / C;
    const uint32_t local = tile - jb->first_unit;
    const uint32_t rb = local / n_ct, ct = local - rb * n_ct;
    const uint32_t col = ct * C + tcol;
    t.live = col < cols;
    const size_t off = ((size_t)rb * g + rph) * jb->row_bytes + (size_t)col * 2;
    t.step = (size_t)rps * jb->row_bytes;
    t.dst = jb->dst + off;
    t.scales = jb->scales;
    t.zeros = jb->zeros;
    t.n_rb = rows / g;
    t.seg0 = (size_t)col * t.n_rb + rb;
    *src = jb->src + off;
    return t;
  };
  auto issue = [&](uint32_t tile, int st) {
    if (tile < p.total_units) {
      const uint8_t* src;
      const ColTile t = describe(tile, &cur_l, &jb_l, &src);
      if (t.live) {
        uint8_t* sd = ring + (size_t)st * stage_bytes;
        for (uint32_t i = 0; i < n_it; ++i) cp_async_16(sd + (size_t)i * 512u, src + (size_t)i * t.step);
      }
    }
    cp_async_commit();  // (possibly empty) group: keeps the group count per tile uniform
  };

  for (int s = 0; s < n_stages - 1; ++s) issue(blockIdx.x + (uint32_t)s * gridDim.x, s);
  int st = 0;
  for (uint32_t tile = blockIdx.x; tile < p.total_units; tile += gridDim.x) {
    {
      int ps = st + n_stages - 1;
      if (ps >= n_stages) ps -= n_stages;
      issue(tile + (uint32_t)(n_stages - 1) * gridDim.x, ps);  // n_stages == 1: the current tile itself
    }
    cp_async_wait_dyn(n_stages - 1);
    __syncwarp();  // every lane's copies of the current tile have landed
    const uint8_t* unused;
    const ColTile t = describe(tile, &cur_c, &jb_c, &unused);
    const uint8_t* sd = ring + (size_t)st * stage_bytes;

    // ---- sweep 1: per-column range over this lane's rows, then across the lane groups ----
    uint32_t mx[4] = {ident_max, ident_max, ident_max, ident_max};
    uint32_t mn[4] = {ident_min, ident_min, ident_min, ident_min};
    {
      uint32_t i = 0;
      for (; i + 4 <= n_it; i += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = lds128(sd + (size_t)(i + u) * 512u);
#pragma unroll
        for (int u = 0; u < 4; ++u) col_range_update<DT, SYM>(mx, mn, v[u]);
      }
      for (; i < n_it; ++i) col_range_update<DT, SYM>(mx, mn, lds128(sd + (size_t)i * 512u));
    }
#pragma unroll 1
    for (uint32_t o = lpr; o < 32u; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mx[j], o))));
        if (!SYM) mn[j] = as_u32(__hmin2_nan(from_u32<T2>(mn[j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mn[j], o))));
      }
    }

    // ---- per-column parameters: 4 packed pairs per lane ----
    PreparedPair<DT, KIND> prep[4];
    bool any_bad = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float mxa = DType<DT>::to_f32((uint16_t)(mx[j] & 0xFFFFu)), mxb = DType<DT>::to_f32((uint16_t)(mx[j] >> 16));
      const float mna = DType<DT>::to_f32((uint16_t)(mn[j] & 0xFFFFu)), mnb = DType<DT>::to_f32((uint16_t)(mn[j] >> 16));
      float sa, za, sb, zb;
      bool bad;
      prep[j].setup2(mxa, SYM ? -mxa : mna, mxb, SYM ? -mxb : mnb, p.qc, SYM, s_tbl, &sa, &za, &sb, &zb, &bad);
      any_bad |= bad;
      if (t.live && rph == 0) {
        // segment order of the reference for quant_dim = 1: column-major over (column, row block)
        const size_t ga = t.seg0 + (size_t)(2 * j) * t.n_rb, gb = ga + t.n_rb;
        if (t.scales) {
          reinterpret_cast<uint16_t*>(t.scales)[ga] = DType<DT>::from_f32(sa);
          reinterpret_cast<uint16_t*>(t.scales)[gb] = DType<DT>::from_f32(sb);
        }
        if (t.zeros) {
          reinterpret_cast<uint16_t*>(t.zeros)[ga] = DType<DT>::from_f32(za);
          reinterpret_cast<uint16_t*>(t.zeros)[gb] = DType<DT>::from_f32(zb);
        }
      }
    }
    // ---- sweep 2: quantise from shared memory, store from registers ----
    const bool any_slow = prep[0].slow() || prep[1].slow() || prep[2].slow() || prep[3].slow();  // includes every suspect segment
    if (t.live && !any_slow) {
      // hot path: reciprocal+FMA quotient fixed at compile time (the IEEE-division variant is for degenerate scales)
      uint32_t i = 0;
      for (; i + 2 <= n_it; i += 2) {
        const uint4 v0 = lds128(sd + (size_t)i * 512u);
        const uint4 v1 = lds128(sd + (size_t)(i + 1) * 512u);
        uint4 o0, o1;
        o0.x = prep[0].template apply<SYM, 0>(v0.x, p.qc); o0.y = prep[1].template apply<SYM, 0>(v0.y, p.qc);
        o0.z = prep[2].template apply<SYM, 0>(v0.z, p.qc); o0.w = prep[3].template apply<SYM, 0>(v0.w, p.qc);
        o1.x = prep[0].template apply<SYM, 0>(v1.x, p.qc); o1.y = prep[1].template apply<SYM, 0>(v1.y, p.qc);
        o1.z = prep[2].template apply<SYM, 0>(v1.z, p.qc); o1.w = prep[3].template apply<SYM, 0>(v1.w, p.qc);
        stg128(t.dst + (size_t)i * t.step, o0);
        stg128(t.dst + (size_t)(i + 1) * t.step, o1);
      }
      for (; i < n_it; ++i) {
        const uint4 v0 = lds128(sd + (size_t)i * 512u);
        uint4 o0;
        o0.x = prep[0].template apply<SYM, 0>(v0.x, p.qc); o0.y = prep[1].template apply<SYM, 0>(v0.y, p.qc);
        o0.z = prep[2].template apply<SYM, 0>(v0.z, p.qc); o0.w = prep[3].template apply<SYM, 0>(v0.w, p.qc);
        stg128(t.dst + (size_t)i * t.step, o0);
      }
    } else if (t.live) {
      uint32_t i = 0;
      for (; i < n_it; ++i) {
        const uint4 v0 = lds128(sd + (size_t)i * 512u);
        uint4 o0;
        o0.x = prep[0].template apply<SYM>(v0.x, p.qc); o0.y = prep[1].template apply<SYM>(v0.y, p.qc);
        o0.z = prep[2].template apply<SYM>(v0.z, p.qc); o0.w = prep[3].template apply<SYM>(v0.w, p.qc);
        stg128(t.dst + (size_t)i * t.step, o0);
        if (any_bad && vec_has_nan<DT>(o0) && p.status) atomicOr(p.status, IWQ_STATUS_NAN);
      }
    }
    __syncwarp();  // the stage may be refilled by the next iteration's copies
    if (++st == n_stages) st = 0;
  }
  cp_async_wait_dyn(0);
}

// ------------------------------------------------------------------------------------------------
// col_tma_kernel (default for quant_dim = 1 with group 32 .. 256): the [g rows x C columns] tile of a warp moves through a
// 2-D TENSOR MAP -- one cp.async.bulk.tensor.2d instruction (SASS: UTMALDG) brings the whole box into the warp's
// shared-memory stage (dense row-major [g][2C bytes]), one UTMASTG sends the quantised box back; no per-vector copy or
// store instructions and no 64-bit address arithmetic in the lanes at all.  Like col_async_kernel the pipeline is
// WARP-PRIVATE (an mbarrier per warp and stage, nothing but __syncwarp() couples the lanes): with two stages the load of
// the warp's next tile is issued right after the range sweep of the current one -- by then the store of the previous tile
// has long finished reading that stage -- and runs under the parameter + apply phases.  The tile is 8 KB for every group
// size (C = 4096 / g columns; 16 KB at g = 256), so a lane always holds 16 vectors of its 8 columns.
// The per-column parameters are computed ONCE per column pair: the 32 / lpr lanes that hold the same 8 columns (one per
// row phase) split the 4 pairs between them -- one setup2 per lane instead of four -- and hand the constants round with
// shuffles.  Tensor maps: one per matrix for the input and one for the output, built on the host
// (cuTensorMapEncodeTiled), in the launch's workspace behind the job table (or in the kernel parameters for a single
// matrix).  Ragged last column tiles: the hardware zero-fills the load and clips the store.
// ------------------------------------------------------------------------------------------------
struct alignas(64) TensorMap2D {
  uint64_t opaque[16];  // CUtensorMap (128 bytes)
};
struct TensorMapPair {
  TensorMap2D in, out;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const TensorMap2D* map, uint32_t c0, uint32_t c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const TensorMap2D* map, uint32_t c0, uint32_t c1, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const TensorMap2D* map, uint32_t c0, uint32_t c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}
// a tensor map that was written to global memory by a copy (generic proxy) must be acquired by the tensormap proxy
__device__ __forceinline__ void tensormap_acquire(const TensorMap2D* map) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(map) : "memory");
}

constexpr int kColTmaWarps = 4;

template <int DT, int KIND, bool SYM>
__global__ void __launch_bounds__(kColTmaWarps * 32) col_tma_kernel(const __grid_constant__ LaunchParams p, const __grid_constant__ TensorMapPair single_maps,
                                                                     const TensorMapPair* __restrict__ maps) {
  using T2 = typename DType<DT>::T2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const FpStep* s_tbl = nullptr;
  if (KIND == KIND_FP) s_tbl = load_fp_lut(p, smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;  // 1 or 2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.lut_smem_bytes) + warp * 2;
  if (lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t C = p.col_tile, g = p.col_group;
  const uint32_t lpr = C >> 3;                 // lanes per tile row: 4 / 8 / 16
  const uint32_t rps = 32u / lpr;              // rows per warp-wide access = row phases: 8 / 4 / 2
  const uint32_t sub = (uint32_t)lane & (lpr - 1u), rph = (uint32_t)lane / lpr;
  const uint32_t n_it = g / rps;               // vectors per lane per tile (16; 32 at g = 256)
  const uint32_t stage_bytes = p.stage_bytes;  // g * C * 2
  uint8_t* ring = smem + p.lut_smem_bytes + 128u + (size_t)warp * n_stages * stage_bytes;
  const uint8_t* lane_off = ring + (size_t)lane * 16u;  // + it * 512: dense rows, a warp-wide LDS.128 reads 512 contiguous bytes
  const uint32_t m_share = rps < 4u ? rps : 4u;         // lanes that split the 4 column pairs of an 8-column group
  const uint32_t my_slot = rph & (m_share - 1u);
  const uint32_t ident_max = SYM ? 0u : (DT == IWQ_F16 ? 0xFC00FC00u : 0xFF80FF80u);  // {0,0} | {-inf,-inf}
  const uint32_t ident_min = DT == IWQ_F16 ? 0x7C007C00u : 0x7F807F80u;                // {+inf,+inf}

  struct Tile {
    const TensorMapPair* map;
    uint8_t* scales;
    uint8_t* zeros;
    size_t seg0;      // segment index of this lane's first column: col * n_rb + rb
    uint32_t n_rb, c0, c1;
    bool live;        // this lane's 8 columns exist
  };
  int cur_l = -1, cur_c = -1, fenced = -1;  // job cursors of the load-issue path and of the arithmetic path
  Job jb_l, jb_c;
  auto describe = [&](uint32_t unit, int* cur, Job* jb) -> Tile {
    Tile t;
    advance_job(p, unit, cur, jb);
    const uint32_t cols = jb->row_bytes >> 1;
    const uint32_t rows = (uint32_t)(jb->bytes / jb->row_bytes);
    const uint32_t n_ct = (cols + C - 1) / C;
    const uint32_t local = unit - jb->first_unit;
    const uint32_t rb = local / n_ct, ct = local - rb * n_ct;
    t.c0 = ct * C;
    t.c1 = rb * g;
    const uint32_t col = t.c0 + sub * 8u;
    t.live = col < cols;
    t.scales = jb->scales;
    t.zeros = jb->zeros;
    t.n_rb = rows / g;
    t.seg0 = (size_t)col * t.n_rb + rb;
    t.map = maps ? &maps[*cur] : &single_maps;
    return t;
  };
  const uint32_t stride_units = gridDim.x * (uint32_t)kColTmaWarps;
  auto issue = [&](uint32_t unit, int st) {  // lane 0 only
    if (unit >= p.total_units) return;
    const Tile t = describe(unit, &cur_l, &jb_l);
    if (maps && fenced != cur_l) {
      tensormap_acquire(&t.map->in);
      tensormap_acquire(&t.map->out);
      fenced = cur_l;
    }
    mbar_arrive_expect_tx(&bars[st], stage_bytes);
    if (p.l2_hint & 3u) tma_load_2d_hint(ring + (size_t)st * stage_bytes, &t.map->in, t.c0, t.c1, &bars[st], (p.l2_hint & 1u) ? l2_policy_evict_first() : l2_policy_evict_last());
    else tma_load_2d(ring + (size_t)st * stage_bytes, &t.map->in, t.c0, t.c1, &bars[st]);
  };

  const uint32_t first = blockIdx.x * (uint32_t)kColTmaWarps + (uint32_t)warp;
  if (n_stages == 2 && lane == 0) issue(first, 0);
  uint32_t phase0 = 0, phase1 = 0;
  int st = 0;
  for (uint32_t unit = first; unit < p.total_units; unit += stride_units) {
    if (n_stages == 1 && lane == 0) {
      tma_store_wait_read<0>();  // the previous tile's store has finished reading the stage
      issue(unit, 0);
    }
    mbar_wait(&bars[st], st ? phase1 : phase0);
    if (st) phase1 ^= 1u; else phase0 ^= 1u;
    const Tile t = describe(unit, &cur_c, &jb_c);
    uint8_t* sd = const_cast<uint8_t*>(lane_off) + (size_t)st * stage_bytes;

    // ---- sweep 1: per-column range over this lane's rows, then across the row phases ----
    uint32_t mx[4] = {ident_max, ident_max, ident_max, ident_max};
    uint32_t mn[4] = {ident_min, ident_min, ident_min, ident_min};
    for (uint32_t i = 0; i < n_it; i += 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = lds128(sd + (size_t)(i + u) * 512u);
#pragma unroll
      for (int u = 0; u < 4; ++u) col_range_update<DT, SYM>(mx, mn, v[u]);
    }
    if (n_stages == 2 && lane == 0) {
      // the other stage held the previous tile: its store was committed a whole range sweep ago
      tma_store_wait_read<0>();
      issue(unit + stride_units, st ^ 1);
    }
#pragma unroll 1
    for (uint32_t o = lpr; o < 32u; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mx[j], o))));
        if (!SYM) mn[j] = as_u32(__hmin2_nan(from_u32<T2>(mn[j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mn[j], o))));
      }
    }

    // ---- per-column parameters: each row-phase lane takes 4 / m_share of the 4 pairs, shuffles distribute them ----
    PreparedPair<DT, KIND> prep[4];
    bool any_bad = false;
#pragma unroll 1
    for (uint32_t pass = 0; pass < 4u; pass += m_share) {
      const uint32_t jm = pass + my_slot;  // the pair this lane computes in this pass
      const uint32_t mxs = jm == 0 ? mx[0] : jm == 1 ? mx[1] : jm == 2 ? mx[2] : mx[3];
      const uint32_t mns = jm == 0 ? mn[0] : jm == 1 ? mn[1] : jm == 2 ? mn[2] : mn[3];
      const float mxa = DType<DT>::to_f32((uint16_t)(mxs & 0xFFFFu)), mxb = DType<DT>::to_f32((uint16_t)(mxs >> 16));
      const float mna = DType<DT>::to_f32((uint16_t)(mns & 0xFFFFu)), mnb = DType<DT>::to_f32((uint16_t)(mns >> 16));
      PreparedPair<DT, KIND> mine;
      float sa, za, sb, zb;
      bool bad;
      mine.setup2(mxa, SYM ? -mxa : mna, mxb, SYM ? -mxb : mnb, p.qc, SYM, s_tbl, &sa, &za, &sb, &zb, &bad);
      if (t.live && rph < m_share) {
        // segment order of the reference for quant_dim = 1: column-major over (column, row block)
        const size_t ga = t.seg0 + (size_t)(2u * jm) * t.n_rb, gb = ga + t.n_rb;
        if (t.scales) {
          reinterpret_cast<uint16_t*>(t.scales)[ga] = DType<DT>::from_f32(sa);
          reinterpret_cast<uint16_t*>(t.scales)[gb] = DType<DT>::from_f32(sb);
        }
        if (t.zeros) {
          reinterpret_cast<uint16_t*>(t.zeros)[ga] = DType<DT>::from_f32(za);
          reinterpret_cast<uint16_t*>(t.zeros)[gb] = DType<DT>::from_f32(zb);
        }
      }
      // pair (pass + s) was computed by the lanes of row phase s: fetch it from lane s * lpr + sub
#pragma unroll
      for (uint32_t sl = 0; sl < 4u; ++sl) {
        if (sl < m_share) {
          const int src = (int)(sl * lpr + sub);
          const uint32_t j = pass + sl;
          const bool b = __shfl_sync(0xFFFFFFFFu, (int)bad, src) != 0;
          any_bad |= b;
          if (j == 0) prep[0].shfl_from(mine, src);
          else if (j == 1) prep[1].shfl_from(mine, src);
          else if (j == 2) prep[2].shfl_from(mine, src);
          else prep[3].shfl_from(mine, src);
        }
      }
    }

    // ---- sweep 2: quantise in place in the stage ----
    const bool any_slow = prep[0].slow() || prep[1].slow() || prep[2].slow() || prep[3].slow();  // includes every suspect segment
    if (!any_slow) {
      for (uint32_t i = 0; i < n_it; i += 2) {
        const uint4 v0 = lds128(sd + (size_t)i * 512u);
        const uint4 v1 = lds128(sd + (size_t)(i + 1) * 512u);
        uint4 o0, o1;
        o0.x = prep[0].template apply<SYM, 0>(v0.x, p.qc); o0.y = prep[1].template apply<SYM, 0>(v0.y, p.qc);
        o0.z = prep[2].template apply<SYM, 0>(v0.z, p.qc); o0.w = prep[3].template apply<SYM, 0>(v0.w, p.qc);
        o1.x = prep[0].template apply<SYM, 0>(v1.x, p.qc); o1.y = prep[1].template apply<SYM, 0>(v1.y, p.qc);
        o1.z = prep[2].template apply<SYM, 0>(v1.z, p.qc); o1.w = prep[3].template apply<SYM, 0>(v1.w, p.qc);
        sts128(sd + (size_t)i * 512u, o0);
        sts128(sd + (size_t)(i + 1) * 512u, o1);
      }
    } else {
      for (uint32_t i = 0; i < n_it; ++i) {
        const uint4 v0 = lds128(sd + (size_t)i * 512u);
        uint4 o0;
        o0.x = prep[0].template apply<SYM>(v0.x, p.qc); o0.y = prep[1].template apply<SYM>(v0.y, p.qc);
        o0.z = prep[2].template apply<SYM>(v0.z, p.qc); o0.w = prep[3].template apply<SYM>(v0.w, p.qc);
        sts128(sd + (size_t)i * 512u, o0);
        if (t.live && any_bad && vec_has_nan<DT>(o0) && p.status) atomicOr(p.status, IWQ_STATUS_NAN);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&t.map->out, t.c0, t.c1, ring + (size_t)st * stage_bytes);
      tma_store_commit();
    }
    if (n_stages == 2) st ^= 1;
  }
  if (lane == 0) tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------------
// col_wide_kernel (quant_dim = 1, group 32 / 64 / 128): ONE CTA per SM streams [128 rows x 256 columns] tiles (64 KB) through a
// ring fed by 2-D tensor-map loads with 512-byte rows -- the seg4 lesson applied to the column-grouped layout: one ring of large
// bulk copies per SM instead of many warp-private 8 KB boxes with 64-byte rows.
//   * producer warp: one cp.async.bulk.tensor.2d per tile (box 256 x 128);
//   * 16 consumer warps, warp w owns rows 8w .. 8w+7 of the tile, lane L the 16-byte chunk L of a row (8 columns): a warp-wide
//     LDS.128 / STS.128 covers one dense 512-byte row (conflict-free), and a lane's 8 columns never change;
//   * sweep 1: per-lane column ranges over its 8 rows -> per-warp partials in shared memory -> barrier (consumers only);
//   * parameters ONCE per column pair: thread t reduces the partials of the warps of its row block for pair t (128 pairs per
//     row block, 128 / g row blocks per tile), runs setup2, writes the scales / zero points and publishes the packed constants
//     in shared memory (transposed so that the readers' LDS.128 are conflict-free) -> barrier;
//   * sweep 2: every lane fetches the constants of its 4 pairs and quantises its 8 rows in place in the stage;
//   * each warp sends its 8 rows back with one cp.async.bulk.tensor.2d store (box 256 x 8) and releases the stage as in seg4.
// Ragged edges: the hardware zero-fills loads and clips stores; parameters of non-existent columns / row blocks are not written.
// ------------------------------------------------------------------------------------------------
#ifndef IWQ_COLW_GROUPS
#define IWQ_COLW_GROUPS 2
#endif
constexpr int kColWideWarps = 16;
constexpr int kColWideGroups = IWQ_COLW_GROUPS;                 // independent consumer groups (alternate tiles)
constexpr int kColWideWG = kColWideWarps / kColWideGroups;      // warps per group
constexpr int kColWideRows = 128, kColWideCols = 256;           // largest tile; LaunchParams::col_rows = 128 or 64 rows per tile
constexpr int kColWideTile = kColWideRows * kColWideCols * 2;  // 64 KB
constexpr int kColWideHeader = 1024;                           // barriers, tile contexts [8], flags
constexpr int kColWidePart = kColWideWarps * 2 * 32 * 16;      // per-warp partial ranges: [warp][max | min][lane] uint4 = 16 KB

struct ColWideCtx {  // 64 bytes
  const TensorMapPair* map;
  uint8_t* scales;
  uint8_t* zeros;
  uint32_t c0, r0, n_rb, cols, job;
  uint32_t pad[5];
};

template <int DT, int KIND, bool SYM, int LUT = LUT_STEP>
__global__ void __launch_bounds__(kColWideWarps * 32 + 32, 1) col_wide_kernel(const __grid_constant__ LaunchParams p, const __grid_constant__ TensorMapPair single_maps,
                                                                                 const TensorMapPair* __restrict__ maps) {
  using T2 = typename DType<DT>::T2;
  using PP = typename std::conditional<KIND == KIND_FP && LUT != LUT_STEP, PreparedPairFp<DT, LUT>, PreparedPair<DT, KIND>>::type;
  constexpr int kU = PP::kWords / 4;  // 16-byte units of one pair's constants
  constexpr int WG = kColWideWG, GT = WG * 32;  // warps / threads of a group
  const uint32_t TR = p.col_rows, RW = TR / (uint32_t)WG;  // rows per tile (128 / 64) and per warp
  const uint32_t CC = p.col_tile;                          // columns per tile (256 / 128): lpr = CC / 8 lanes cover a row, a warp-wide
  const uint32_t lpr = CC >> 3;                            // 16-byte access 32 / lpr consecutive (dense) rows
  const uint32_t tile_bytes = TR * CC * 2u;
  const uint32_t n_acc = RW * lpr / 32u;                   // warp-wide accesses per warp and sweep (a multiple of 4)
  const uint32_t half_cc = CC >> 1;                        // column pairs per row block
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  ColWideCtx* ctxs = reinterpret_cast<ColWideCtx*>(smem + 128);       // [8]
  uint32_t* flags_all = reinterpret_cast<uint32_t*>(smem + 128 + 8 * 64);  // [group][2] "a segment of this tile is suspect", by tile parity
  const FpStep* s_tbl = nullptr;
  if (KIND == KIND_FP) s_tbl = load_fp_lut(p, smem + kColWideHeader);
  uint8_t* part_all = smem + kColWideHeader + p.lut_smem_bytes;
  constexpr uint32_t kPartBytes = SYM ? kColWidePart / 2 : kColWidePart;  // symmetric forms keep no minima
  constexpr uint32_t kPW = SYM ? 128u : 256u;                              // words of partials per warp
  uint8_t* pars_all = part_all + kPartBytes;
  const uint32_t g = p.col_group;
  const uint32_t wpb = g / RW;                       // warps per row block
  const uint32_t n_pairs = (TR / g) * half_cc;
  const uint32_t pars_bytes = n_pairs * PP::kWords * 4u;
  uint8_t* ring = pars_all + (size_t)kColWideGroups * pars_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], WG);
    }
    for (int i = 0; i < 2 * 4 * kColWideGroups; ++i) flags_all[i] = 0u;
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kColWideWarps) {
    // ---------------- producer ----------------
    if (lane == 0) {
      int stage = 0, cur = -1, fenced = -1;
      uint32_t phase = 0;
      Job jb;
      const uint64_t pol = (p.l2_hint & 1u) ? l2_policy_evict_first() : l2_policy_evict_last();
      for (uint32_t unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        mbar_wait(&empty[stage], phase ^ 1u);
        advance_job(p, unit, &cur, &jb);
        const uint32_t cols = jb.row_bytes >> 1;
        const uint32_t rows = (uint32_t)(jb.bytes / jb.row_bytes);
        const uint32_t n_ct = (cols + CC - 1) / CC;
        const uint32_t local = unit - jb.first_unit;
        const uint32_t rt = local / n_ct, ct = local - rt * n_ct;
        ColWideCtx c;
        c.map = maps ? &maps[cur] : &single_maps;
        c.scales = jb.scales;
        c.zeros = jb.zeros;
        c.c0 = ct * CC;
        c.r0 = rt * TR;
        c.n_rb = rows / g;
        c.cols = cols;
        c.job = (uint32_t)cur;
        ctxs[stage] = c;
        if (maps && fenced != cur) {
          tensormap_acquire(&c.map->in);
          fenced = cur;
        }
        mbar_arrive_expect_tx(&full[stage], tile_bytes);
        uint8_t* dst = ring + (size_t)stage * tile_bytes;
        if (p.l2_hint & 3u) tma_load_2d_hint(dst, &c.map->in, c.c0, c.r0, &full[stage], pol);
        else tma_load_2d(dst, &c.map->in, c.c0, c.r0, &full[stage]);
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }

  // ---------------- consumers: group gi takes tiles gi, gi + groups, ... of this CTA's sequence ----------------
  const uint32_t ident_max = SYM ? 0u : (DT == IWQ_F16 ? 0xFC00FC00u : 0xFF80FF80u);
  const uint32_t ident_min = DT == IWQ_F16 ? 0x7C007C00u : 0x7F807F80u;
  PP uni;  // launch-uniform members of the prepared constants (minifloat / BFP kinds): set once
  {
    float a, b, c, d;
    bool bad;
    uni.setup2(1.0f, -1.0f, 1.0f, -1.0f, p.qc, SYM, s_tbl, &a, &b, &c, &d, &bad);
  }
  const uint32_t gi = (uint32_t)warp / WG, wg = (uint32_t)warp % WG;
  const uint32_t my_blk = wg / wpb;                             // row block of this warp's rows within the tile
  // the warps of a ROW BLOCK synchronise among themselves only (group 64: 4 warps, group 32: 2): the row blocks of a tile are
  // independent, so a block's parameter phase overlaps the sweeps of the others
  const uint32_t blk_threads = wpb * 32u;
  const uint32_t t_b = threadIdx.x - gi * GT - my_blk * blk_threads;  // thread within the row block's warps
  uint8_t* part = part_all + (size_t)gi * (kPartBytes / kColWideGroups);
  uint8_t* pars = pars_all + (size_t)gi * pars_bytes;
  uint32_t* flags = flags_all + 2u * (gi * 4u + my_blk);
  uint4* part_mx = reinterpret_cast<uint4*>(part) + (size_t)wg * (kPW / 4) + lane;  // [warp][0][lane]
  uint4* part_mn = part_mx + 32;                                                 // [warp][1][lane]
  const uint32_t* part_w = reinterpret_cast<const uint32_t*>(part);             // word view: [warp][max | min][128 pairs]
  const uint32_t bar_id = 1u + gi * ((uint32_t)WG / wpb) + my_blk;  // <= 8 named barriers
  int fenced = -1;
  uint32_t parity = 0;
  for (uint32_t i = gi;; i += kColWideGroups) {
    const uint32_t unit = blockIdx.x + i * gridDim.x;
    if (unit >= p.total_units) break;
    const uint32_t stage = i % (uint32_t)n_stages, phase = (i / (uint32_t)n_stages) & 1u;
    mbar_wait(&full[stage], phase);
    const ColWideCtx c = ctxs[stage];
    uint8_t* wbase = ring + (size_t)stage * tile_bytes + (size_t)wg * (RW * CC * 2u);
    uint8_t* sd = wbase + (size_t)lane * 16u;

    // ---- sweep 1 ----
    uint32_t mx[4] = {ident_max, ident_max, ident_max, ident_max};
    uint32_t mn[4] = {ident_min, ident_min, ident_min, ident_min};
#pragma unroll 1
    for (uint32_t r0 = 0; r0 < n_acc; r0 += 4) {
      uint4 v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) v[r] = lds128(sd + (r0 + r) * 512);
#pragma unroll
      for (int r = 0; r < 4; ++r) col_range_update<DT, SYM>(mx, mn, v[r]);
    }
    if (lpr == 16u) {  // 256-byte rows: lanes L and L ^ 16 hold the two rows of an access, same columns
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mx[j] = as_u32(__hmax2_nan(from_u32<T2>(mx[j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mx[j], 16))));
        if (!SYM) mn[j] = as_u32(__hmin2_nan(from_u32<T2>(mn[j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mn[j], 16))));
      }
    }
    *part_mx = make_uint4(mx[0], mx[1], mx[2], mx[3]);
    if (!SYM) *part_mn = make_uint4(mn[0], mn[1], mn[2], mn[3]);
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(blk_threads) : "memory");

    // ---- parameters of this row block: one column pair per thread (the other row blocks / the other group keep sweeping) ----
    if (t_b == 0) flags[parity ^ 1u] = 0u;  // the other parity's flag: read by everybody before this tile's first barrier, set again after the next one
    for (uint32_t q = t_b; q < half_cc; q += blk_threads) {
      const uint32_t b = my_blk;
      uint32_t mxs = ident_max, mns = ident_min;
      for (uint32_t w = b * wpb; w < (b + 1u) * wpb; ++w) {
        mxs = as_u32(__hmax2_nan(from_u32<T2>(mxs), from_u32<T2>(part_w[w * kPW + q])));
        if (!SYM) mns = as_u32(__hmin2_nan(from_u32<T2>(mns), from_u32<T2>(part_w[w * kPW + 128u + q])));
      }
      const float mxa = DType<DT>::to_f32((uint16_t)(mxs & 0xFFFFu)), mxb = DType<DT>::to_f32((uint16_t)(mxs >> 16));
      const float mna = DType<DT>::to_f32((uint16_t)(mns & 0xFFFFu)), mnb = DType<DT>::to_f32((uint16_t)(mns >> 16));
      PP mine;
      float sa, za, sb, zb;
      bool bad;
      mine.setup2(mxa, SYM ? -mxa : mna, mxb, SYM ? -mxb : mnb, p.qc, SYM, s_tbl, &sa, &za, &sb, &zb, &bad);
      const uint32_t col = c.c0 + 2u * q, rb = c.r0 / g + b;
      if (col < c.cols && rb < c.n_rb) {
        // segment order of the reference for quant_dim = 1: column-major over (column, row block)
        const size_t ga = (size_t)col * c.n_rb + rb, gb = ga + c.n_rb;
        if (c.scales) {
          reinterpret_cast<uint16_t*>(c.scales)[ga] = DType<DT>::from_f32(sa);
          reinterpret_cast<uint16_t*>(c.scales)[gb] = DType<DT>::from_f32(sb);
        }
        if (c.zeros) {
          reinterpret_cast<uint16_t*>(c.zeros)[ga] = DType<DT>::from_f32(za);
          reinterpret_cast<uint16_t*>(c.zeros)[gb] = DType<DT>::from_f32(zb);
        }
        if (bad || mine.slow()) atomicOr(&flags[parity], 1u);
      }
      // publish: pair q of block b is pair j = q & 3 of chunk q >> 2; unit u of it lives at [(b * 4 + j) * kU + u][chunk] (16 B each)
      uint32_t wds[PP::kWords];
      int n = 0;
      mine.map_from(mine, [&](uint32_t v) {
        wds[n++] = v;
        return v;
      });
      uint4* dst = reinterpret_cast<uint4*>(pars) + (size_t)((b * 4u + (q & 3u)) * kU) * 32u + (q >> 2);
#pragma unroll
      for (int u = 0; u < kU; ++u) dst[u * 32] = make_uint4(wds[4 * u], wds[4 * u + 1], wds[4 * u + 2], wds[4 * u + 3]);
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(blk_threads) : "memory");

    // ---- sweep 2: the constants of this lane's 4 pairs, then quantise in place ----
    PP prep[4];
    const bool suspect = flags[parity] != 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4* src = reinterpret_cast<const uint4*>(pars) + (size_t)((my_blk * 4u + (uint32_t)j) * kU) * 32u + (lane & (lpr - 1u));
      uint32_t wds[PP::kWords];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const uint4 x = src[u * 32];
        wds[4 * u] = x.x; wds[4 * u + 1] = x.y; wds[4 * u + 2] = x.z; wds[4 * u + 3] = x.w;
      }
      int n = 0;
      prep[j].map_from(uni, [&](uint32_t) { return wds[n++]; });
    }
    if (!suspect) {
#pragma unroll 2
      for (uint32_t r = 0; r < n_acc; ++r) {
        const uint4 v = lds128(sd + r * 512);
        uint4 o;
        o.x = prep[0].template apply<SYM, 0>(v.x, p.qc); o.y = prep[1].template apply<SYM, 0>(v.y, p.qc);
        o.z = prep[2].template apply<SYM, 0>(v.z, p.qc); o.w = prep[3].template apply<SYM, 0>(v.w, p.qc);
        sts128(sd + r * 512, o);
      }
    } else {
      const bool live = c.c0 + ((uint32_t)lane & (lpr - 1u)) * 8u < c.cols && c.r0 + wg * RW < c.n_rb * g;
      bool nan_seen = false;
#pragma unroll 1
      for (uint32_t r = 0; r < n_acc; ++r) {
        const uint4 v = lds128(sd + r * 512);
        uint4 o;
        o.x = prep[0].template apply<SYM>(v.x, p.qc); o.y = prep[1].template apply<SYM>(v.y, p.qc);
        o.z = prep[2].template apply<SYM>(v.z, p.qc); o.w = prep[3].template apply<SYM>(v.w, p.qc);
        sts128(sd + r * 512, o);
        nan_seen |= vec_has_nan<DT>(o);
      }
      if (live && nan_seen && p.status) atomicOr(p.status, IWQ_STATUS_NAN);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      if (maps && fenced != (int)c.job) {
        tensormap_acquire(&c.map->out);
        fenced = (int)c.job;
      }
      tma_store_2d(&c.map->out, c.c0, c.r0 + wg * RW, wbase);
      tma_store_commit();
      // the stage goes back to the producer as soon as this warp's rows have been read out of shared memory (with two groups and
      // three stages the other group is waiting for exactly this stage's next tile)
      tma_store_wait_read<0>();
      mbar_arrive(&empty[stage]);
    }
    parity ^= 1u;
  }
  if (lane == 0) tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------------
// col_double_kernel: quant_dim = 1 with the double-approximate decode (quant_linear.py:288-363).
// The reference forms groups of four codes along the flattened transpose of the [segments, group] code
// matrix; with (rows / group) % 4 == 0 those are the codes at the same row offset k of FOUR consecutive row
// blocks of one column.  So a tile is 4 row blocks x 256 columns, a lane again owns 8 columns, and for every k
// it holds the four vectors (one per block) whose columns form the groups: the neighbourhood never leaves the
// thread.  Element stage: normalise with the block's scale, quantise through the exact step table, rebuild
// the code, decode the four codes jointly (double_approx_decode4), multiply by the scale.  This is scalar
// integer work (~100 ops per 4 elements): correctness first, it is the research path of the reference.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256, 2) col_double_kernel(const __grid_constant__ LaunchParams p) {
  using D = DType<DT>;
  using T2 = typename D::T2;
  extern __shared__ __align__(1024) uint8_t smem[];
  // [2 parities][8 warps][4 blocks][32 lanes] uint4 = 32 KB, then the step table
  uint4* part = reinterpret_cast<uint4*>(smem);
  const FpStep* s_tbl = load_fp_lut(p, smem + 32768);
  // behind the step table and the value table: the published scales [16 (block, pair)][32 lanes] float4 + flags
  float4* exch = reinterpret_cast<float4*>(smem + 32768 + p.lut_smem_bytes + p.lut_extra_bytes);
  uint32_t* exch_flags = reinterpret_cast<uint32_t*>(exch + 16 * 32);
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile = 4 row blocks x C columns.  C = 256 (default): a lane owns 8 columns of a full-width access.  C = 64 (IWQ_COL_DBL_TILE): 8
  // lanes cover a 128-byte row piece and a warp-wide access takes 4 consecutive rows, so that the second sweep finds the tile in L2
  // (with C = 256 the 296 resident tiles are 76 MB: 1.68 x read traffic) -- measured 35-48 % SLOWER (profiles/r02_sweep_col19.jsonl):
  // the 16 scale computations per lane and tile are then spread over 4 x fewer elements; the kernel is instruction-bound, not HBM-bound
  const uint32_t C = p.col_tile, lpr = C >> 3, rps = 32u / lpr;
  const uint32_t sub = (uint32_t)lane & (lpr - 1u), rph = (uint32_t)lane / lpr;
  int cur = -1;
  Job jb;
  uint32_t par = 0;
  FpQ q1;
  q1.S = 1.f; q1.r = 1.f; q1.mid = 0.f; q1.mid_h = 0.f; q1.bad = false; q1.slow_div = false;
  FpTable<DT, true> grid;  // only its scale-independent members are used (clamp bounds, step table)
  grid.setup2(q1, q1, p.qc, s_tbl);
  for (uint32_t tile = blockIdx.x; tile < p.total_units; tile += gridDim.x) {
    advance_job(p, tile, &cur, &jb);
    const uint32_t cols = jb.row_bytes >> 1;
    const uint32_t rows = (uint32_t)(jb.bytes / jb.row_bytes);
    const uint32_t g = p.col_group;
    const uint32_t n_ct = (cols + C - 1) / C;
    const uint32_t local = tile - jb.first_unit;
    const uint32_t rb4 = local / n_ct, ct = local - rb4 * n_ct;  // rb4 counts groups of 4 row blocks
    const uint32_t col = ct * C + sub * 8u;
    const bool live = col < cols;
    const size_t off = (size_t)rb4 * 4 * g * jb.row_bytes + (size_t)col * 2;
    const uint8_t* src = jb.src + off;
    uint8_t* dst = jb.dst + off;
    const size_t blk = (size_t)g * jb.row_bytes;

    // ---- sweep 1: absmax per (block, column) ----
    uint32_t mx[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int j = 0; j < 4; ++j) mx[b][j] = 0u;
    if (live) {
#pragma unroll 2
      for (uint32_t k = (uint32_t)warp * rps + rph; k < g; k += 8u * rps) {
        uint4 v[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) v[b] = ldg128(src + b * blk + (size_t)k * jb.row_bytes);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t w4[4] = {v[b].x, v[b].y, v[b].z, v[b].w};
#pragma unroll
          for (int j = 0; j < 4; ++j) mx[b][j] = as_u32(__hmax2_nan(from_u32<T2>(mx[b][j]), from_u32<T2>(w4[j] & 0x7FFF7FFFu)));
        }
      }
    }
    for (uint32_t o = lpr; o < 32u; o <<= 1) {  // the row phases of a warp-wide access hold the same columns
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int j = 0; j < 4; ++j) mx[b][j] = as_u32(__hmax2_nan(from_u32<T2>(mx[b][j]), from_u32<T2>(__shfl_xor_sync(0xFFFFFFFFu, mx[b][j], o))));
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) part[((par * 8 + warp) * 4 + b) * 32 + lane] = make_uint4(mx[b][0], mx[b][1], mx[b][2], mx[b][3]);
    __syncthreads();
    for (int w = 0; w < 8; ++w) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint4 a = part[((par * 8 + w) * 4 + b) * 32 + lane];
        mx[b][0] = as_u32(__hmax2_nan(from_u32<T2>(mx[b][0]), from_u32<T2>(a.x)));
        mx[b][1] = as_u32(__hmax2_nan(from_u32<T2>(mx[b][1]), from_u32<T2>(a.y)));
        mx[b][2] = as_u32(__hmax2_nan(from_u32<T2>(mx[b][2]), from_u32<T2>(a.z)));
        mx[b][3] = as_u32(__hmax2_nan(from_u32<T2>(mx[b][3]), from_u32<T2>(a.w)));
      }
    }
    par ^= 1u;

    // ---- scales per (block, column pair): warp w computes pairs 2w and 2w + 1 of the 16 (block, pair) combinations of every lane's
    //      columns and publishes {S, 1/S} through shared memory (round 1: all 8 warps computed all 16 -- 2000 of the 10 000
    //      instructions a lane spends on a tile) ----
    {
      uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (warp == i) {
          m0 = mx[(2 * i) >> 2][(2 * i) & 3];
          m1 = mx[(2 * i + 1) >> 2][(2 * i + 1) & 3];
        }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t m = h ? m1 : m0;
        const int idx = 2 * warp + h, b = idx >> 2, j = idx & 3;
        const float mxa = D::to_f32((uint16_t)(m & 0xFFFFu)), mxb = D::to_f32((uint16_t)(m >> 16));
        const FpQ qa = fp_segment_params<DT>(mxa, -mxa, p.qc), qb = fp_segment_params<DT>(mxb, -mxb, p.qc);
        exch[idx * 32 + lane] = make_float4(qa.S, qb.S, qa.r, qb.r);
        exch_flags[idx * 32 + lane] = ((qa.bad || qb.bad) ? 1u : 0u) | ((qa.slow_div || qb.slow_div) ? 2u : 0u);
        if (live && rph == 0u && jb.scales) {
          const size_t n_rb = rows / g;
          const size_t ga = (size_t)(col + 2 * j) * n_rb + (size_t)rb4 * 4 + b;
          reinterpret_cast<uint16_t*>(jb.scales)[ga] = D::from_f32(qa.S);
          reinterpret_cast<uint16_t*>(jb.scales)[ga + n_rb] = D::from_f32(qb.S);
        }
      }
    }
    __syncthreads();
    float2 S[4][4], R[4][4];
    bool any_bad = false, slow = false;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 e = exch[(b * 4 + j) * 32 + lane];
        const uint32_t f = exch_flags[(b * 4 + j) * 32 + lane];
        S[b][j] = make_float2(e.x, e.y);
        R[b][j] = make_float2(e.z, e.w);
        any_bad |= (f & 1u) != 0u;
        slow |= (f & 2u) != 0u;
      }
    }
    if (live && any_bad && p.status) atomicOr(p.status, IWQ_STATUS_NAN);

    // ---- sweep 2: packed quantiser -> info words -> packed group logic -> value table ----
    if (live) {
      const uint32_t hf2 = ((uint32_t)p.qc.hi_field << 8) * 0x00010001u;                    // hi_align_exp_field, ae position
      const uint32_t maxe2 = ((uint32_t)((1 << p.qc.exp_bits) - 1) << 8) * 0x00010001u;    // all-ones exponent
      const uint32_t val_saddr = smem_u32(smem + 32768 + p.lut_smem_bytes);
      auto sweep2 = [&](auto slow_c) {
      constexpr bool kSlow = decltype(slow_c)::value;  // IEEE division only for degenerate scales: chosen once per tile
      // the four vectors of the NEXT row offset are requested before the ~550 instructions of the current one run (the loop used
      // to stall on its own loads: long scoreboard 41 % of the stall samples, profiles/r02_ncu_col_double.txt)
      const uint32_t k_step = 8u * rps;
      uint32_t k = (uint32_t)warp * rps + rph;
      uint4 nv[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) nv[b] = k < g ? ldg128(src + b * blk + (size_t)k * jb.row_bytes) : make_uint4(0u, 0u, 0u, 0u);
      for (; k < g; k += k_step) {
        uint4 v[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) v[b] = nv[b];
        if (k + k_step < g) {
#pragma unroll
          for (int b = 0; b < 4; ++b) nv[b] = ldg128(src + b * blk + (size_t)(k + k_step) * jb.row_bytes);
        }
        uint32_t o[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t info[4], sg[4];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t w2 = b == 0 ? (j == 0 ? v[0].x : j == 1 ? v[0].y : j == 2 ? v[0].z : v[0].w)
                              : b == 1 ? (j == 0 ? v[1].x : j == 1 ? v[1].y : j == 2 ? v[1].z : v[1].w)
                              : b == 2 ? (j == 0 ? v[2].x : j == 1 ? v[2].y : j == 2 ? v[2].z : v[2].w)
                                       : (j == 0 ? v[3].x : j == 1 ? v[3].y : j == 2 ? v[3].z : v[3].w);
            const T2 n = kSlow ? ieee_quot2<DT>(from_u32<T2>(w2), S[b][j]) : fast_quot2<DT>(from_u32<T2>(w2), S[b][j], R[b][j]);
            info[b] = grid.steps_select(n);
            sg[b] = grid.codec_sign(n);
          }
          // group = the four blocks; low halves and high halves are two independent groups (neighbouring columns)
          const uint32_t target = double_approx_target2<DT>(info, hf2, maxe2);
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t idx = (info[b] & kDblCodeMask) | target | (sg[b] >> 8);             // target << 8 | sign << 7 | code
            uint16_t lo, hi;
            asm("ld.shared.u16 %0, [%1];" : "=h"(lo) : "r"(val_saddr + (idx & 0xFFFFu) * 2u));
            asm("ld.shared.u16 %0, [%1];" : "=h"(hi) : "r"(val_saddr + (idx >> 16) * 2u));
            const T2 dec = from_u32<T2>((uint32_t)lo | ((uint32_t)hi << 16));
            o[b][j] = as_u32(__hmul2(dec, D::from_f32x2(S[b][j].x, S[b][j].y)));  // decoded * scales   quant_linear.py:606
          }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) stg128(dst + b * blk + (size_t)k * jb.row_bytes, make_uint4(o[b][0], o[b][1], o[b][2], o[b][3]));
      }
      };
      if (slow) sweep2(std::true_type{}); else sweep2(std::false_type{});
    }
  }
}

// ------------------------------------------------------------------------------------------------
// seg_double_kernel: the double-approximate decode with quant_dim = 0 (quant_linear.py:484-488 -> :288-363).
// The reference's groups of four run along the flattened transpose of the [segments, group] code matrix: with the number
// of segments a multiple of 4 they are the codes at the same offset k of FOUR consecutive segments.  A "quad" is those
// 4 segments = 4 * group contiguous elements of the row-major matrix (possibly spanning rows); group / 8 lanes share a quad and every lane holds the four vectors at
// its offset (one per segment), so the neighbourhood never leaves the thread -- the same element logic as
// col_double_kernel, where the two halves of a packed register are two independent groups (here: offsets k and k + 1).
// Plain 128-bit global loads / stores: every byte crosses HBM once each way; the element stage (~100 scalar-ish ops per
// 4 elements) bounds this research path, not the memory system.  blockDim.x = 256; unit = 8 warps x (256 / group) quads.
// ------------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256, 2) seg_double_kernel(const __grid_constant__ LaunchParams p) {
  using D = DType<DT>;
  using T2 = typename D::T2;
  extern __shared__ __align__(1024) uint8_t smem[];
  const FpStep* s_tbl = load_fp_lut(p, smem);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t g = p.col_group, vps = g >> 3;            // vectors (= lanes) per segment: 4, 8, 16 or 32
  const uint32_t qpw = 32u / vps;                            // quads per warp
  const uint32_t qslot = (uint32_t)lane / vps, v = (uint32_t)lane & (vps - 1u);
  const uint32_t quads_per_unit = 8u * qpw;
  int cur = -1;
  Job jb;
  FpQ q1;
  q1.S = 1.f; q1.r = 1.f; q1.mid = 0.f; q1.mid_h = 0.f; q1.bad = false; q1.slow_div = false;
  FpTable<DT, true> grid;  // only its scale-independent members are used (clamp bounds, step table)
  grid.setup2(q1, q1, p.qc, s_tbl);
  const uint32_t hf2 = ((uint32_t)p.qc.hi_field << 8) * 0x00010001u;
  const uint32_t maxe2 = ((uint32_t)((1 << p.qc.exp_bits) - 1) << 8) * 0x00010001u;
  const uint32_t val_saddr = smem_u32(smem + p.lut_smem_bytes);
  for (uint32_t unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
    advance_job(p, unit, &cur, &jb);
    const uint64_t n_quads = jb.bytes / (8ull * g);        // 4 * g elements of 2 bytes
    const uint64_t quad = (uint64_t)(unit - jb.first_unit) * quads_per_unit + (uint32_t)warp * qpw + qslot;
    const bool live = quad < n_quads;
    const size_t off = (size_t)quad * 8u * g + (size_t)v * 16u;
    const uint8_t* src = jb.src + off;
    uint8_t* dst = jb.dst + off;
    const size_t seg_bytes = (size_t)g * 2u;
    uint4 vec[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) vec[b] = live ? ldg128(src + b * seg_bytes) : make_uint4(0u, 0u, 0u, 0u);
    // ---- absmax per segment: this lane's vector, then across the vps lanes of the quad ----
    uint32_t mx[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) mx[b] = vec_range<DT, true>(vec[b]);
#pragma unroll 1
    for (uint32_t o = vps >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int b = 0; b < 4; ++b) mx[b] = range_max<DT>(mx[b], __shfl_xor_sync(0xFFFFFFFFu, mx[b], o));
    }
    float2 S[4], R[4];
    uint32_t S2[4];
    bool any_bad = false, slow = false;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      float m, unused;
      range_unpack<DT>(mx[b], &m, &unused);
      const FpQ qa = fp_segment_params<DT>(m, -m, p.qc);
      S[b] = make_float2(qa.S, qa.S);
      R[b] = make_float2(qa.r, qa.r);
      S2[b] = as_u32(D::from_f32x2(qa.S, qa.S));
      any_bad |= qa.bad;
      slow |= qa.slow_div;
      if (live && v == 0 && jb.scales) reinterpret_cast<uint16_t*>(jb.scales)[quad * 4u + b] = D::from_f32(qa.S);
    }
    if (live && any_bad && p.status) atomicOr(p.status, IWQ_STATUS_NAN);
    // ---- element stage: packed quantiser -> info words -> packed group logic -> value table ----
    uint32_t o[4][4];
    auto stage = [&](auto slow_c) {
      constexpr bool kSlow = decltype(slow_c)::value;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t info[4], sg[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t w2 = j == 0 ? vec[b].x : j == 1 ? vec[b].y : j == 2 ? vec[b].z : vec[b].w;
          const T2 n = kSlow ? ieee_quot2<DT>(from_u32<T2>(w2), S[b]) : fast_quot2<DT>(from_u32<T2>(w2), S[b], R[b]);
          info[b] = grid.steps_select(n);
          sg[b] = grid.codec_sign(n);
        }
        const uint32_t target = double_approx_target2<DT>(info, hf2, maxe2);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const uint32_t idx = (info[b] & kDblCodeMask) | target | (sg[b] >> 8);
          uint16_t lo, hi;
          asm("ld.shared.u16 %0, [%1];" : "=h"(lo) : "r"(val_saddr + (idx & 0xFFFFu) * 2u));
          asm("ld.shared.u16 %0, [%1];" : "=h"(hi) : "r"(val_saddr + (idx >> 16) * 2u));
          const T2 dec = from_u32<T2>((uint32_t)lo | ((uint32_t)hi << 16));
          o[b][j] = as_u32(__hmul2(dec, from_u32<T2>(S2[b])));  // decoded * scales   quant_linear.py:606
        }
      }
    };
    if (slow) stage(std::true_type{}); else stage(std::false_type{});
    if (live) {
#pragma unroll
      for (int b = 0; b < 4; ++b) stg128(dst + b * seg_bytes, make_uint4(o[b][0], o[b][1], o[b][2], o[b][3]));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-tensor granularity, pass 1
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v != v) {
    atomicMax(reinterpret_cast<int*>(addr), 0x7FFFFFFF);  // NaN dominates (positive NaN is the largest int pattern)
  } else if (v >= 0.0f) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  } else {
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
  }
}

// returning form: the value comes back only after the atomic has been performed at L2, so an instruction that
// depends on it is ordered after it without a fence
__device__ __forceinline__ uint32_t atomic_max_float_ret(float* addr, float v) {
  if (v != v) return (uint32_t)atomicMax(reinterpret_cast<int*>(addr), 0x7FFFFFFF);
  if (v >= 0.0f) return (uint32_t)atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  return atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

static __global__ void range_init_kernel(float* ranges, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ranges[i] = -INFINITY;
}

template <int DT, bool SYM>
__global__ void __launch_bounds__(kConsumers) minmax_kernel(const __grid_constant__ LaunchParams p) {
  __shared__ uint32_t s_part[kConsumerWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ident = SYM ? 0u : (DT == IWQ_F16 ? 0xFC00FC00u : 0xFF80FF80u);
  int cur = -1, acc_job = -1;
  uint32_t acc = ident;
  Job jb;

  auto flush = [&](int job) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc = range_max<DT>(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, o));
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t all = s_part[0];
      for (int w = 1; w < kConsumerWarps; ++w) all = range_max<DT>(all, s_part[w]);
      float mx, mn;
      range_unpack<DT>(all, &mx, &mn);
      atomic_max_float(&p.ranges_out[2 * job], mx);
      if (!SYM) atomic_max_float(&p.ranges_out[2 * job + 1], -mn);
    }
    __syncthreads();
    acc = ident;
  };

  // every CTA takes a CONTIGUOUS run of tiles: it streams one contiguous region of memory and crosses only the
  // few matrix boundaries inside it, so the block reduction + atomics (flush) run once per (CTA, matrix) instead
  // of once per tile -- with the small shards of the 8-GPU per-tensor configuration (4..16 MB per matrix) a
  // strided assignment flushed on nearly every tile
  const bool strided = (p.l2_hint & 8u) != 0;  // slab mode: one matrix per launch, walk it like the apply pass does
  const uint32_t n_units = p.total_units - p.unit_begin;
  const uint32_t per = (n_units + gridDim.x - 1) / gridDim.x;
  const uint32_t t_begin = p.unit_begin + (strided ? blockIdx.x : blockIdx.x * per);
  const uint32_t t_stop = p.unit_begin + blockIdx.x * per + per;
  const uint32_t t_end = strided ? p.total_units : (t_stop < p.total_units ? t_stop : p.total_units);
  const uint32_t t_step = strided ? gridDim.x : 1u;
  for (uint32_t tile = t_begin; tile < t_end; tile += t_step) {
    advance_job(p, tile, &cur, &jb);
    if (cur != acc_job) {
      if (acc_job >= 0) flush(acc_job);  // block-uniform: every thread sees the same tile sequence
      acc_job = cur;
    }
    const TileCtx t = make_tile_ctx(p, jb, cur, tile, false);
    const uint8_t* gbase = jb.src + (t.vec0 << 4);
    uint4 v[kVPT];
#pragma unroll
    for (int it = 0; it < kVPT; ++it) {
      const uint32_t idx = (uint32_t)warp * (32 * kVPT) + it * 32 + lane;
      // out-of-range lanes re-read vector 0 of the tile: harmless for a max/min
      v[it] = ldg128(gbase + (size_t)(idx < t.n_vec ? idx : 0u) * 16);
    }
#pragma unroll
    for (int it = 0; it < kVPT; ++it) acc = range_max<DT>(acc, vec_range<DT, SYM>(v[it]));
  }
  if (acc_job >= 0) flush(acc_job);
}

// ------------------------------------------------------------------------------------------------
// minmax_tma_kernel: the same range pass fed by the bulk-copy ring of seg4_tma_kernel (read only: a stage goes back to the
// producer as soon as its vectors are in registers).  One CTA per SM, 48 KB tiles, contiguous run of tiles per CTA (so the block
// reduction + atomics run once per (CTA, matrix)).  tools/tmabench.cu: a read-only ring of this shape streams at 7.1 TB/s, the
// 128-bit global loads of minmax_kernel at 4.9-6 TB/s on the shard sizes of the 8-GPU per-tensor configuration.
// shared memory: barriers and TileCtx as in seg4 (TileCtx::pad carries the job index) | [640, ...) per-warp partial ranges | ring
// ------------------------------------------------------------------------------------------------
template <int DT, bool SYM>
__global__ void __launch_bounds__(kConsumers + 32, 1) minmax_tma_kernel(const __grid_constant__ LaunchParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  TileCtx* ctx = reinterpret_cast<TileCtx*>(smem + 128);
  uint32_t* s_part = reinterpret_cast<uint32_t*>(smem + 640);
  uint8_t* data = smem + kSmemHeader;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kConsumerWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const TileWalk walk = make_walk(p);

  if (warp == kConsumerWarps) {
    if (lane == 0) {
      int stage = 0, cur = -1;
      uint32_t phase = 0;
      Job jb;
      const uint32_t hint = p.l2_hint & 3u;
      const uint64_t pol = hint == 1u ? l2_policy_evict_first() : (hint == 2u ? l2_policy_evict_last() : 0ull);
      for (uint32_t tile = walk.cur; tile < walk.end; tile += walk.step) {
        mbar_wait(&empty[stage], phase ^ 1u);
        advance_job(p, tile, &cur, &jb);
        TileCtx t = make_tile_ctx(p, jb, cur, tile, false);
        t.pad = (uint32_t)cur;
        ctx[stage] = t;
        const uint32_t bytes = t.n_vec * 16u;
        mbar_arrive_expect_tx(&full[stage], bytes);
        if (hint) tma_load_1d_hint(data + (size_t)stage * kTileBytes, jb.src + (t.vec0 << 4), bytes, &full[stage], pol);
        else tma_load_1d(data + (size_t)stage * kTileBytes, jb.src + (t.vec0 << 4), bytes, &full[stage]);
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }

  const uint32_t ident = SYM ? 0u : (DT == IWQ_F16 ? 0xFC00FC00u : 0xFF80FF80u);
  uint32_t acc = ident;
  int acc_job = -1;
  auto flush = [&](int job) {  // all consumer warps (the tile sequence is CTA-uniform)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc = range_max<DT>(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, o));
    if (lane == 0) s_part[warp] = acc;
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    if (threadIdx.x == 0) {
      uint32_t all = s_part[0];
      for (int w = 1; w < kConsumerWarps; ++w) all = range_max<DT>(all, s_part[w]);
      float mx, mn;
      range_unpack<DT>(all, &mx, &mn);
      atomic_max_float(&p.ranges_out[2 * job], mx);
      if (!SYM) atomic_max_float(&p.ranges_out[2 * job + 1], -mn);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    acc = ident;
  };
  const uint32_t lane_vec = (uint32_t)warp * 128u + (uint32_t)lane * 4u;
  const uint32_t rot = (uint32_t)lane >> 1;
  int stage = 0;
  uint32_t phase = 0;
  for (uint32_t tile = walk.cur; tile < walk.end; tile += walk.step) {
    mbar_wait(&full[stage], phase);
    const uint32_t n_vec = ctx[stage].n_vec;
    const int job = (int)ctx[stage].pad;
    if (job != acc_job) {
      if (acc_job >= 0) flush(acc_job);
      acc_job = job;
    }
    const uint8_t* lbase = data + (size_t)stage * kTileBytes + (size_t)lane_vec * 16;
    uint4 v[4];
    bool ok[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t j = (k + rot) & 3u;
      ok[k] = lane_vec + j < n_vec;
      v[k] = ok[k] ? lds128(lbase + j * 16) : make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);  // the vectors are in registers: hand the stage back
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) acc = range_max<DT>(acc, vec_range<DT, SYM>(v[k]));
    if (++stage == n_stages) {
      stage = 0;
      phase ^= 1u;
    }
  }
  if (acc_job >= 0) flush(acc_job);
}

// ------------------------------------------------------------------------------------------------
// pertensor_fused_kernel (EXPERIMENTAL, opt-in: correct, but slower than the three-step path -- DESIGN.md section 4):
// per-tensor granularity in ONE launch, also when every matrix is row-split across the GPUs of a box (BASELINE
// config 4) -- range pass, range exchange and apply pass fused.
//
//   * the matrices are cut into BATCHES of whole matrices (16 MB by default: a few batches fit the 126 MB L2);
//   * RANGE and APPLY tiles of a batch go through the same TMA ring.  A CTA claims the tiles of an item from a
//     per-batch counter; an end-of-item marker travels through the ring.  Range tiles are read-only: the consumer
//     warps reduce them from shared memory, add their partial {max, -min} to this GPU's per-matrix ranges with
//     returning atomics and then count themselves in on the batch's arrival counter;
//   * lanes 1..31 of the producer warp run the exchange: the CTA with blockIdx.x == b (mod gridDim.x) publishes batch
//     b once its counter is complete -- it stores the ranges into slot [rank] of EVERY rank's exchange buffer through
//     the peer-mapped (NVLink / NVSwitch) pointers, fences at system scope and raises flag [rank][b] on every rank:
//     no NCCL call, no second launch; every CTA gathers the batches in order (all ranks' flags up -> maximum over
//     the ranks' slots -> shared-memory ring);
//   * lane 0 of the producer warp issues the TMA copies; before the apply tiles of batch b it waits for the batch's
//     gathered ranges and stamps them into the tile contexts; the consumers run the seg4 element stage with the
//     given range.  The tiles were read by the range item a few batches ago, so the apply item can find them in L2;
//   * schedule per CTA: R(0) .. R(D-1), A(0), R(D), A(1), ... with look-ahead D, so that a batch's exchange latency
//     hides behind D - 1 batches of work.
// All CTAs must be co-resident (persistent grid sized by the occupancy API), as for any arrival-counter scheme.
// world == 1 runs the same code with the local buffers as the only "peer" (and gathers straight from the counter).
// ------------------------------------------------------------------------------------------------
struct FusedParams {
  float* local_range;            // [2 * n_jobs] {max, -min}, initialised to -inf by the host
  uint32_t* arrive;              // [n_batches] arrival counters, zero-initialised
  const uint32_t* batch_tile0;   // [n_batches + 1] first tile of every batch
  const uint32_t* batch_job0;    // [n_batches + 1] first job of every batch
  float* const* peer_slots;      // [world] rank r's exchange buffer: float[world][2 * n_jobs], as mapped on THIS GPU
  uint32_t* const* peer_flags;   // [world] rank r's flag array: uint32[world][n_batches]
  int32_t n_batches, world, rank;
  uint32_t epoch;                // value a raised flag carries in this call (flags persist across calls)
  uint32_t spin_limit;           // bounded wait: afterwards IWQ_STATUS_TIMEOUT is raised instead of hanging
  int32_t lookahead;             // batches the range pass runs ahead of the apply pass (>= 1)
  uint32_t poll_ns;              // sleep of the exchange lanes between polls
  uint32_t debug;                // diagnostics only (IWQ_FUSED_DEBUG): 2 = no apply tiles, 3 = no range tiles (results are then wrong)
  unsigned long long* trace;     // diagnostics only: [n_batches][4] globaltimer stamps written by CTA 0 (nullptr: off)
  uint32_t* tile_ctr;            // [2][n_batches] work-stealing counters (range items, apply items), zero-initialised
};

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }

// The CTA's work list is a fixed interleaving of RANGE and APPLY items over the batches -- R(0), R(1), A(0), R(2),
// A(1), ... -- walked identically by the producer lane and by the consumers; BOTH kinds of item are tiles streamed
// through the same TMA ring (a range item is read-only: the consumers reduce it from shared memory and hand the
// stage straight back), so the range pass runs at the same bandwidth as the apply pass.
constexpr int kFusedMaxBatchJobs = 16;  // matrices per batch (the host closes a batch earlier if need be)
constexpr int kFusedRing = 8;           // batches whose reduced ranges may be parked in shared memory (lookahead < kFusedRing)

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The exchange state machine, run by lanes 1..31 of the producer warp (lane 0 issues the TMA copies and must not wait
// on microsecond-long global / peer loads).  The 31 lanes work in parallel so that a step costs ONE memory latency:
//   publish: this CTA is the publisher of the batches b == blockIdx.x (mod gridDim.x): once all consumer warps of
//            the grid have arrived for such a batch, the lanes copy the batch's local ranges into slot [rank] of every
//            rank's exchange buffer (peer-mapped stores), fence at system scope and raise flag [rank][b] everywhere;
//   gather : batches in order: lane r + 1 polls rank r's flag; when all are up, lane (job, rank) pairs fetch the
//            slots, the maximum over the ranks goes into the shared-memory ring s_rng[gather_b % kFusedRing][job].
struct FusedExchange {
  int pub_b;      // next batch this CTA has to publish (>= n_batches: none left)
  int gather_b;   // next batch whose ranges are to be gathered
  uint32_t spins;
};
constexpr uint32_t kXchgMask = 0xFFFFFFFEu;  // lanes 1..31

template <bool SYM>
__device__ __forceinline__ void fused_exchange_step(const LaunchParams& p, const FusedParams& f, FusedExchange& x, float2* s_rng, float2* s_tmp,
                                                    const uint32_t* s_bj, int lane) {
  const int nb = f.n_batches, xl = lane - 1;  // xl: 0..30
  if (f.world == 1) x.pub_b = nb;  // nothing to publish on a single GPU
  if (x.pub_b < nb) {
    uint32_t cnt = 0;
    if (xl == 0) cnt = ld_acquire_gpu_u32(&f.arrive[x.pub_b]);
    cnt = __shfl_sync(kXchgMask, cnt, 1);
    if (cnt == gridDim.x * (uint32_t)kConsumerWarps) {
      const uint32_t j0 = s_bj[x.pub_b], j1 = s_bj[x.pub_b + 1];
      const uint32_t nv = 2u * (j1 - j0), total = nv * (uint32_t)f.world;
      const size_t slot_off = ((size_t)(f.epoch & 1u) * f.world + f.rank) * 2u * p.n_jobs + 2u * j0;  // [parity][source rank][2 * n_jobs]
      for (uint32_t i = (uint32_t)xl; i < total; i += 31u) {
        const uint32_t r = i / nv, k = i - r * nv;
        f.peer_slots[r][slot_off + k] = __ldcg(&f.local_range[2u * j0 + k]);  // peer-mapped store (NVLink), local when r == rank
      }
      __threadfence_system();
      __syncwarp(kXchgMask);
      if (xl < f.world) st_release_sys_u32(&f.peer_flags[xl][(size_t)f.rank * nb + x.pub_b], f.epoch);
      x.pub_b += (int)gridDim.x;
    }
  }
  if (x.gather_b < nb && f.world == 1) {
    // single GPU: no slots, no flags -- the batch is ready when its arrival counter is complete
    uint32_t cnt = 0;
    if (xl == 0) cnt = ld_acquire_gpu_u32(&f.arrive[x.gather_b]);
    cnt = __shfl_sync(kXchgMask, cnt, 1);
    if (cnt == gridDim.x * (uint32_t)kConsumerWarps) {
      const uint32_t j0 = s_bj[x.gather_b], j1 = s_bj[x.gather_b + 1];
      for (uint32_t j = j0 + (uint32_t)xl; j < j1; j += 31u)
        s_rng[(x.gather_b % kFusedRing) * kFusedMaxBatchJobs + (j - j0)] = make_float2(__ldcg(&f.local_range[2 * j]), SYM ? 0.f : -__ldcg(&f.local_range[2 * j + 1]));
      __syncwarp(kXchgMask);
      ++x.gather_b;
    }
    return;
  }
  if (x.gather_b < nb) {
    const uint32_t* my_flags = f.peer_flags[f.rank];
    bool ok = true;
    for (int r = xl; r < f.world; r += 31) ok = ok && (ld_acquire_sys_u32(&my_flags[(size_t)r * nb + x.gather_b]) == f.epoch);
    const bool all = __all_sync(kXchgMask, ok);
    const bool give_up = !all && (++x.spins > f.spin_limit);  // bounded: do not hang the GPU on a missing peer
    if (all || give_up) {
      if (give_up && xl == 0 && p.status) atomicOr(p.status, IWQ_STATUS_TIMEOUT);
      const float* my_slots = f.peer_slots[f.rank] + (size_t)(f.epoch & 1u) * f.world * 2u * p.n_jobs;
      const uint32_t j0 = s_bj[x.gather_b], j1 = s_bj[x.gather_b + 1];
      const uint32_t nj = j1 - j0, total = nj * (uint32_t)f.world;
      for (uint32_t i = (uint32_t)xl; i < total; i += 31u) {  // (job, rank) pairs in parallel: one latency
        const uint32_t r = i / nj, j = i - r * nj;
        const float* src = &my_slots[(size_t)r * 2u * p.n_jobs + 2u * (j0 + j)];
        s_tmp[j * 16u + r] = make_float2(ld_relaxed_sys_f32(src), SYM ? -INFINITY : ld_relaxed_sys_f32(src + 1));
      }
      __syncwarp(kXchgMask);
      for (uint32_t j = (uint32_t)xl; j < nj; j += 31u) {
        float mx = -INFINITY, nmn = -INFINITY;
        for (int r = 0; r < f.world; ++r) {
          const float2 v = s_tmp[j * 16u + r];
          mx = max_nan(mx, v.x);
          nmn = max_nan(nmn, v.y);
        }
        s_rng[(x.gather_b % kFusedRing) * kFusedMaxBatchJobs + j] = make_float2(mx, -nmn);
      }
      __syncwarp(kXchgMask);
      ++x.gather_b;
      x.spins = 0;
    }
  }
}

template <int DT, int KIND, bool SYM>
__global__ void __launch_bounds__(kConsumers + 32, 3) pertensor_fused_kernel(const __grid_constant__ LaunchParams p, const __grid_constant__ FusedParams f) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMaxStages;
  TileCtx* ctx = reinterpret_cast<TileCtx*>(smem + 128);
  const FpStep* s_tbl = nullptr;
  uint8_t* data = smem + kSmemHeader + p.lut_smem_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stages = p.n_stages;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], kConsumerWarps);
    }
    fence_mbar_init();
  }
  if (KIND == KIND_FP) s_tbl = load_fp_lut(p, smem + kSmemHeader);
  const int nb = f.n_batches;
  // batch tables in shared memory: an item is only 1..4 tiles per CTA, two dependent global loads per item and warp
  // would cost as much as the item itself
  uint32_t* s_bt = reinterpret_cast<uint32_t*>(data + (size_t)n_stages * kTileBytes + (size_t)(kFusedRing + 16) * kFusedMaxBatchJobs * sizeof(float2));
  uint32_t* s_bj = s_bt + (nb + 1);
  for (int i = threadIdx.x; i <= nb; i += blockDim.x) {
    s_bt[i] = f.batch_tile0[i];
    s_bj[i] = f.batch_job0[i];
  }
  if (threadIdx.x == 0) *reinterpret_cast<volatile uint32_t*>(smem + 900) = 0u;
  __syncthreads();
  // the schedule, walked identically by the producer lane and the consumers: RANGE items run `lookahead` batches
  // ahead of the APPLY items -- R(0) .. R(D-1), A(0), R(D), A(1), R(D+1), ... -- so that a batch's publication and
  // exchange latency (~10 us) hides behind D - 1 batches of work while D + 1 batches stay L2-resident.  BOTH kinds
  // of item are tiles streamed through the same TMA ring (a range item is read-only: the consumers reduce it from
  // shared memory and hand the stage straight back).
  const int D = f.lookahead;
  int next_r = 0, next_a = 0;
  auto next_item = [&](int* batch) -> bool {  // returns true for a RANGE item; call while next_a < nb
    if (next_r < nb && next_r < next_a + D) {
      *batch = next_r++;
      return true;
    }
    *batch = next_a++;
    return false;
  };

  if (warp == kConsumerWarps) {
    float2* s_rng = reinterpret_cast<float2*>(data + (size_t)n_stages * kTileBytes);  // [kFusedRing][kFusedMaxBatchJobs]
    volatile uint32_t* s_gathered = reinterpret_cast<volatile uint32_t*>(smem + 900);  // batches whose ranges are in s_rng
    if (lane >= 1) {
      // ---------------- exchange lanes 1..31: publish this CTA's batches, gather every batch's ranges ----------------
      float2* s_tmp = s_rng + kFusedRing * kFusedMaxBatchJobs;  // [kFusedMaxBatchJobs][16 ranks] scratch
      FusedExchange x;
      x.pub_b = (int)blockIdx.x;
      x.gather_b = 0;
      x.spins = 0;
      while (x.gather_b < nb || x.pub_b < nb) {
        const int before = x.gather_b;
        fused_exchange_step<SYM>(p, f, x, s_rng, s_tmp, s_bj, lane);
        if (x.gather_b != before) {
          __threadfence_block();
          if (lane == 1) *s_gathered = (uint32_t)x.gather_b;
          if (f.trace && blockIdx.x == 0 && lane == 1) f.trace[4 * before + 1] = globaltimer_ns();
        } else {
          __nanosleep(f.poll_ns);
        }
      }
    } else if (lane == 0) {
      // ---------------- producer lane: TMA issue ----------------
      int stage = 0;
      uint32_t phase = 0;
      // two cursors over the job table (range stream, apply stream), each with the NEXT job's first tile cached in a
      // register: a range tile is only ~1 us of HBM time per CTA, a dependent global load per tile would cap the rate
      struct Cursor {
        int cur;
        uint32_t next_first;
        Job jb;
      } cr, ca;
      const Job* jobs = job_table(p);
      auto seek = [&](Cursor& c, uint32_t tile) {
        if (c.cur >= 0 && tile < c.next_first) return;
        int j = c.cur < 0 ? 0 : c.cur;
        if (c.cur < 0 && p.n_jobs > 8) {
          int lo = 0, hi = p.n_jobs - 1;
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].first_unit <= tile) lo = mid; else hi = mid - 1;
          }
          j = lo;
        }
        while (j + 1 < p.n_jobs && tile >= jobs[j + 1].first_unit) ++j;
        c.cur = j;
        c.jb = jobs[j];
        c.next_first = j + 1 < p.n_jobs ? jobs[j + 1].first_unit : 0xFFFFFFFFu;
      };
      cr.cur = -1;
      ca.cur = -1;
      const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
      while (next_a < nb) {
        int b;
        const bool is_range = next_item(&b);
        if (!is_range) {
          while (*s_gathered <= (uint32_t)b) __nanosleep(64);  // the batch's ranges from every rank (exchange lanes)
          __threadfence_block();
          if (f.trace && blockIdx.x == 0) f.trace[4 * b + 2] = globaltimer_ns();
        }
        const uint32_t j0 = s_bj[b];
        const uint32_t t1 = ((f.debug & 7u) == (is_range ? 3u : 2u)) ? 0u : s_bt[b + 1];
        // tiles of an item are CLAIMED from a per-batch counter, not assigned statically: with only 2..5 tiles per CTA
        // and item a static split leaves the slowest CTA a whole tile behind, and every CTA waits for it per batch
        uint32_t* ctr = &f.tile_ctr[(is_range ? 0 : nb) + b];
        for (uint32_t tile = s_bt[b] + atomicAdd(ctr, 1u); tile < t1; tile = s_bt[b] + atomicAdd(ctr, 1u)) {
          if (is_range) seek(cr, tile); else seek(ca, tile);
          const int cur = is_range ? cr.cur : ca.cur;
          const Job& jb = is_range ? cr.jb : ca.jb;
          mbar_wait(&empty[stage], phase ^ 1u);
          TileCtx t = make_tile_ctx(p, jb, cur, tile, false);
          t.pad = is_range ? (1u + (uint32_t)cur) : 0u;  // range item: job index + 1
          if (!is_range) {
            const float2 g = s_rng[(b % kFusedRing) * kFusedMaxBatchJobs + ((uint32_t)cur - j0)];
            t.given_mx = g.x;
            t.given_mn = g.y;
          }
          ctx[stage] = t;
          const uint32_t bytes = t.n_vec * 16u;
          mbar_arrive_expect_tx(&full[stage], bytes);
          // range tiles come back within a few batches: keep them in L2; apply tiles are never read again
          if (f.debug & 8u) tma_load_1d_hint(data + (size_t)stage * kTileBytes, jb.src + (t.vec0 << 4), bytes, &full[stage], is_range ? pol_keep : pol_drop);
          else tma_load_1d(data + (size_t)stage * kTileBytes, jb.src + (t.vec0 << 4), bytes, &full[stage]);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        {  // end-of-item marker: the consumers cannot know how many tiles this CTA claimed
          mbar_wait(&empty[stage], phase ^ 1u);
          TileCtx t;
          t.dst = nullptr; t.scales = nullptr; t.zeros = nullptr; t.codes = nullptr;
          t.vec0 = 0; t.n_vec = 0; t.given_mx = 0.f; t.given_mn = 0.f; t.pad = 0xFFFFFFFFu;
          ctx[stage] = t;
          mbar_arrive(&full[stage]);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const uint32_t lane_vec = (uint32_t)warp * 128u + (uint32_t)lane * 4u;
  const uint32_t rot = (uint32_t)lane >> 1;
  const uint32_t ident = SYM ? 0u : (DT == IWQ_F16 ? 0xFC00FC00u : 0xFF80FF80u);
  int stage = 0, pending = -1;  // pending: stage whose bulk store may still be reading shared memory
  bool aged = false;            // ... and at least one further tile has been consumed since that store was issued
  uint32_t phase = 0;
  while (next_a < nb) {
    int b;
    const bool is_range = next_item(&b);
    uint32_t acc = ident, tok = 0;  // tok: depends on every atomic of this item having returned
    int acc_job = -1;
    auto flush = [&](int job) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) acc = range_max<DT>(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, o));
      if (lane == 0) {
        float mx, mn;
        range_unpack<DT>(acc, &mx, &mn);
        tok ^= atomic_max_float_ret(&f.local_range[2 * job], mx);
        if (!SYM) tok ^= atomic_max_float_ret(&f.local_range[2 * job + 1], -mn);
      }
      acc = ident;
    };
    for (;;) {
      mbar_wait(&full[stage], phase);
      const TileCtx t = ctx[stage];
      if (t.pad == 0xFFFFFFFFu) {  // end of the item: hand the marker's stage back and leave
        __syncwarp();
        if (lane == 0) {
          if (pending >= 0 && aged) {  // (a marker ages the parked stage like a tile does: the ring must keep turning)
            tma_store_wait_read<0>();
            mbar_arrive(&empty[pending]);
          }
          mbar_arrive(&empty[stage]);
        }
        if (pending >= 0 && aged) pending = -1;
        aged = pending >= 0;
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
        break;
      }
      uint8_t* lbase = data + (size_t)stage * kTileBytes + (size_t)lane_vec * 16;
      if (is_range) {
        const int job = (int)t.pad - 1;
        if (job != acc_job) {
          if (acc_job >= 0) flush(acc_job);
          acc_job = job;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // lanes past the end of a ragged tile re-read its first vector: harmless for a max/min
          const uint32_t jj = ((uint32_t)j + rot) & 3u;
          const uint4 v = lds128(lane_vec + jj < t.n_vec ? lbase + jj * 16 : data + (size_t)stage * kTileBytes);
          acc = range_max<DT>(acc, vec_range<DT, SYM>(v));
        }
        __syncwarp();
        if (lane == 0) {
          if (pending >= 0 && aged) {  // the apply item's store was issued a whole tile ago: it has read its stage by now
            tma_store_wait_read<0>();
            mbar_arrive(&empty[pending]);
          }
          mbar_arrive(&empty[stage]);
        }
        if (pending >= 0 && aged) pending = -1;
        aged = pending >= 0;
      } else {
        if (t.n_vec == (uint32_t)(kTileBytes / 16) && t.codes == nullptr)
          seg4_consume<DT, KIND, SYM, true, true>(t, p, s_tbl, lbase, lane_vec, rot, lane, 1, warp, nullptr);
        else
          seg4_consume<DT, KIND, SYM, true, false>(t, p, s_tbl, lbase, lane_vec, rot, lane, 1, warp, nullptr);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const uint32_t w0 = (uint32_t)warp * 128u;
          if (w0 < t.n_vec) {
            const uint32_t nv = t.n_vec - w0 < 128u ? t.n_vec - w0 : 128u;
            tma_store_1d(t.dst + (size_t)w0 * 16, data + (size_t)stage * kTileBytes + (size_t)w0 * 16, nv * 16u);
          }
          tma_store_commit();
          tma_store_wait_read<1>();
          if (pending >= 0) mbar_arrive(&empty[pending]);
        }
        pending = stage;
        aged = false;
      }
      if (++stage == n_stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    if (!is_range && f.trace && blockIdx.x == 0 && warp == 0 && lane == 0) f.trace[4 * b + 3] = globaltimer_ns();
    if (is_range) {
      if (acc_job >= 0) flush(acc_job);
      // arrival of this warp for batch b: a release-ordered fire-and-forget add (the warp does not wait for it);
      // the batch's publisher CTA watches the counter reach gridDim.x * 8
      // (no fence: the increment is computed from the atomics' return values, so it issues only after all of them have
      //  been performed at L2, the point of coherence the publisher reads from -- a ~1 us round trip instead of the
      //  ~4 us MEMBAR/ERRBAR a release costs this warp while its bulk stores are in flight)
      if (lane == 0) {
        uint32_t one;
        asm volatile("{ .reg .b32 t; and.b32 t, %1, 0; add.u32 %0, t, 1; }" : "=r"(one) : "r"(tok));
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(&f.arrive[b]), "r"(one) : "memory");
        if (f.trace && blockIdx.x == 0 && warp == 0) f.trace[4 * b + 0] = globaltimer_ns();
      }
    }
  }
  if (lane == 0) tma_store_wait_all();
}

// ------------------------------------------------------------------------------------------------
// element-wise test kernels (exhaustive parity checks)
// ------------------------------------------------------------------------------------------------
// _fp_scale (fp4_quantize_cpu.py:37-44) with an element-wise scale array: y = grid(clamp(rD(x / S), +-max_float)); the
// result is NOT multiplied back by S (the caller does that, :72).  IEEE division: S is arbitrary here.
template <int DT>
__global__ void fp_scale_kernel(const uint32_t* __restrict__ x, const uint32_t* __restrict__ scale, uint32_t* __restrict__ y, int64_t n_pairs,
                                const __grid_constant__ QuantConsts qc) {
  __shared__ FpEntry s_tbl[40];
  if (threadIdx.x < 40) s_tbl[threadIdx.x] = qc.tbl[threadIdx.x];
  __syncthreads();
  using T2 = typename DType<DT>::T2;
  FpQ q;
  q.S = 1.0f; q.r = 1.0f; q.mid = 0.f; q.mid_h = 0.f; q.bad = false; q.slow_div = false;
  FpTable<DT> f;
  f.setup(q, qc);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 w = DType<DT>::to_f32x2(from_u32<T2>(x[i])), sc = DType<DT>::to_f32x2(from_u32<T2>(scale[i]));
    const T2 n = DType<DT>::from_f32x2(ieee_div(w.x, sc.x), ieee_div(w.y, sc.y));       // tensor / S                   :38
    const uint32_t nb = as_u32(n);
    uint32_t vb = as_u32(f.grid(n, nb, s_tbl));                                          // clamp, log2 scale, round     :39-43
    vb |= __hneu2_mask(n, n);                                                            // 0 / 0 stays NaN through the clamp
    y[i] = vb;
  }
}

template <int DT>
__global__ void fp_roundtrip_kernel(const uint32_t* __restrict__ x, uint32_t* __restrict__ y, uint16_t* __restrict__ codes, int64_t n_pairs,
                                    const __grid_constant__ QuantConsts qc) {
  __shared__ FpEntry s_tbl[40];
  if (threadIdx.x < 40) s_tbl[threadIdx.x] = qc.tbl[threadIdx.x];
  __syncthreads();
  using T2 = typename DType<DT>::T2;
  FpQ q;
  q.S = 1.0f;
  q.r = 1.0f;
  q.mid = 0.f;
  q.mid_h = 0.f;
  q.bad = false;
  q.slow_div = false;
  FpTable<DT> f;
  f.setup(q, qc);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
    T2 v = f.grid(from_u32<T2>(x[i]), x[i], s_tbl);
    uint32_t vb = as_u32(v);
    y[i] = vb;
    if (codes) {
      uint32_t c0 = fp_code_of<DT>((uint16_t)(vb & 0xFFFFu), qc.exp_bits, qc.mant_bits);
      uint32_t c1 = fp_code_of<DT>((uint16_t)(vb >> 16), qc.exp_bits, qc.mant_bits);
      codes[i] = (uint16_t)(c0 | (c1 << 8));
    }
  }
}

// every (w, s): rD(fast_quot) vs rD(IEEE w/s); blockIdx.x enumerates s, threads sweep w
template <int DT>
__global__ void division_selftest_kernel(unsigned long long* mismatches) {
  const uint32_t top = DT == IWQ_F16 ? 0x7C00u : 0x7F80u;
  unsigned long long bad = 0;
  for (uint32_t sb = 1 + blockIdx.x; sb < top; sb += gridDim.x) {
    const float s = DType<DT>::to_f32((uint16_t)sb);
    const float r = ieee_rcp(s);
    if (DT == IWQ_BF16 && !(s > 1e-30f && s < 1e30f)) continue;
    for (uint32_t wb = threadIdx.x; wb < top; wb += blockDim.x) {
      const float w = DType<DT>::to_f32((uint16_t)wb);
      const float want = ieee_div(w, s);
      if (DT == IWQ_BF16 && !(fabsf(want) < 1e37f && (fabsf(want) > 1e-30f || want == 0.f))) continue;
      const uint16_t a = DType<DT>::from_f32(want);
      const uint16_t b = DType<DT>::from_f32(fast_quot(w, s, r));
      bad += (a != b);
    }
  }
  if (bad) atomicAdd(mismatches, bad);
}

}  // namespace iwq
