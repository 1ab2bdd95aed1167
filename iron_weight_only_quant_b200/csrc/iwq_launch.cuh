// iwq_launch.cuh -- typed launch wrappers.  Each (dtype, kind) pair is explicitly instantiated in its
// own translation unit (iwq_inst_*.cu) so the library builds in parallel; iwq_host.cu only sees the
// declarations below.
#pragma once

#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <tuple>

#include "iwq_kernels.cuh"

namespace iwq {

struct LaunchTuning {
  int stages;        // ring depth of the TMA kernels (2..8)
  int ctas_per_sm;   // 0 = as many as fit
};

// grid size for a persistent kernel: resident CTAs per SM (occupancy API) x SM count, cached
template <class Kernel>
inline cudaError_t persistent_grid(Kernel kernel, int threads, size_t smem, int cap_per_sm, uint32_t total_units, int* grid_out) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, size_t, int>, std::pair<int, int>> cache;  // -> (occ, sms)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  auto key = std::make_tuple(reinterpret_cast<const void*>(kernel), threads, smem, dev);
  std::pair<int, int> v;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
      // kernels with a dynamic ring opt in to the full 227 KB once (a later, smaller value would lower the
      // limit again); kernels without dynamic shared memory must not be touched (static smem + 227 KB > limit)
      if (smem > 0) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
      }
      int occ = 0, sms = 0;
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
      if (e != cudaSuccess) return e;
      e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (e != cudaSuccess) return e;
      if (occ < 1) return cudaErrorLaunchOutOfResources;
      v = std::make_pair(occ, sms);
      cache[key] = v;
    } else {
      v = it->second;
    }
  }
  int occ = v.first;
  if (cap_per_sm > 0 && occ > cap_per_sm) occ = cap_per_sm;
  long long g = (long long)occ * v.second;
  if (g > (long long)total_units) g = total_units;
  if (g < 1) g = 1;
  *grid_out = (int)g;
  return cudaSuccess;
}

template <int DT, int KIND>
cudaError_t launch_seg(const LaunchParams& p, bool sym, bool given, int variant, const LaunchTuning& tune, cudaStream_t st);

template <int DT, int KIND>
cudaError_t launch_row(const LaunchParams& p, bool sym, int consumers, const LaunchTuning& tune, cudaStream_t st);

template <int DT, int KIND>
cudaError_t launch_col(const LaunchParams& p, bool sym, const LaunchTuning& tune, cudaStream_t st);

template <int DT>
cudaError_t launch_col_double(const LaunchParams& p, cudaStream_t st);

template <int DT, int KIND>
cudaError_t launch_fused(const LaunchParams& p, const FusedParams& f, bool sym, const LaunchTuning& tune, cudaStream_t st, int* grid_out);

template <int DT>
cudaError_t launch_seg_double(const LaunchParams& p, cudaStream_t st);

template <int DT, int KIND>
cudaError_t launch_col_tma(const LaunchParams& p, const TensorMapPair& single, const TensorMapPair* maps, bool sym, const LaunchTuning& tune, cudaStream_t st);

#ifdef IWQ_INSTANTIATE

template <int DT, int KIND, bool SYM>
static cudaError_t launch_col_tma_one(const LaunchParams& p, const TensorMapPair& single, const TensorMapPair* maps, const LaunchTuning& tune, cudaStream_t st) {
  if (p.col_wide) {
    const size_t n_pairs = (size_t)(p.col_rows / p.col_group) * (p.col_tile / 2);
    const size_t smem = kColWideHeader + p.lut_smem_bytes + (SYM ? kColWidePart / 2 : kColWidePart) + kColWideGroups * n_pairs * PreparedPair<DT, KIND>::kWords * 4 + (size_t)p.n_stages * p.stage_bytes;
    int grid = 0;
    cudaError_t e;
    if (KIND == KIND_FP && p.qc.lut_direct) {
      auto kw = col_wide_kernel<DT, KIND, SYM, KIND == KIND_FP ? LUT_DIRECT : LUT_STEP>;
      e = persistent_grid(kw, kColWideWarps * 32 + 32, smem, 1, p.total_units, &grid);
      if (e != cudaSuccess) return e;
      kw<<<grid, kColWideWarps * 32 + 32, smem, st>>>(p, single, maps);
    } else {
      auto kw = col_wide_kernel<DT, KIND, SYM>;
      e = persistent_grid(kw, kColWideWarps * 32 + 32, smem, 1, p.total_units, &grid);
      if (e != cudaSuccess) return e;
      kw<<<grid, kColWideWarps * 32 + 32, smem, st>>>(p, single, maps);
    }
    return cudaGetLastError();
  }
  auto k = col_tma_kernel<DT, KIND, SYM>;
  // [step table][mbarriers, 128 B][ring: warps x stages x stage]
  const size_t smem = p.lut_smem_bytes + 128 + (size_t)kColTmaWarps * p.n_stages * p.stage_bytes;
  int grid = 0;
  const uint32_t cta_units = (p.total_units + kColTmaWarps - 1) / kColTmaWarps;
  cudaError_t e = persistent_grid(k, kColTmaWarps * 32, smem, tune.ctas_per_sm, cta_units, &grid);
  if (e != cudaSuccess) return e;
  k<<<grid, kColTmaWarps * 32, smem, st>>>(p, single, maps);
  return cudaGetLastError();
}
template <int DT, int KIND>
cudaError_t launch_col_tma(const LaunchParams& p, const TensorMapPair& single, const TensorMapPair* maps, bool sym, const LaunchTuning& tune, cudaStream_t st) {
  return sym ? launch_col_tma_one<DT, KIND, true>(p, single, maps, tune, st) : launch_col_tma_one<DT, KIND, false>(p, single, maps, tune, st);
}

template <int DT>
cudaError_t launch_seg_double(const LaunchParams& p, cudaStream_t st) {
  auto k = seg_double_kernel<DT>;
  const size_t smem = p.lut_smem_bytes + p.lut_extra_bytes;
  int grid = 0;
  cudaError_t e = persistent_grid(k, 256, smem, 0, p.total_units, &grid);
  if (e != cudaSuccess) return e;
  k<<<grid, 256, smem, st>>>(p);
  return cudaGetLastError();
}

// the fused per-tensor kernel: the grid size is also the arrival-counter target, so it is returned to the caller
// (who sized the work list for it); `*grid_out` > 0 on entry means "launch with exactly this grid"
template <int DT, int KIND, bool SYM>
static cudaError_t launch_fused_one(LaunchParams p, const FusedParams& f, const LaunchTuning& tune, cudaStream_t st, int* grid_out) {
  auto k = pertensor_fused_kernel<DT, KIND, SYM>;
  const size_t smem = kSmemHeader + p.lut_smem_bytes + (size_t)p.n_stages * kTileBytes + (size_t)(kFusedRing + 16) * kFusedMaxBatchJobs * sizeof(float2) + 2 * ((size_t)f.n_batches + 1) * 4;
  int grid = 0;
  cudaError_t e = persistent_grid(k, kConsumers + 32, smem, tune.ctas_per_sm, 0xFFFFFFFFu, &grid);
  if (e != cudaSuccess) return e;
  *grid_out = grid;
  k<<<grid, kConsumers + 32, smem, st>>>(p, f);
  return cudaGetLastError();
}
template <int DT, int KIND>
cudaError_t launch_fused(const LaunchParams& p, const FusedParams& f, bool sym, const LaunchTuning& tune, cudaStream_t st, int* grid_out) {
  return sym ? launch_fused_one<DT, KIND, true>(p, f, tune, st, grid_out) : launch_fused_one<DT, KIND, false>(p, f, tune, st, grid_out);
}

template <int DT>
cudaError_t launch_col_double(const LaunchParams& p, cudaStream_t st) {
  auto k = col_double_kernel<DT>;
  const size_t smem = 32768 + p.lut_smem_bytes + p.lut_extra_bytes + 16 * 32 * (16 + 4);  // partials | step table | value table | published scales + flags
  int grid = 0;
  cudaError_t e = persistent_grid(k, 256, smem, 0, p.total_units, &grid);
  if (e != cudaSuccess) return e;
  k<<<grid, 256, smem, st>>>(p);
  return cudaGetLastError();
}

template <int DT, int KIND, bool SYM>
static cudaError_t launch_col_one(LaunchParams p, const LaunchTuning& tune, cudaStream_t st) {
  if (p.col_tile != 0 && p.col_rpl == 0) {  // warp-private cp.async ring (group 32 .. 512)
    auto k = col_async_kernel<DT, KIND, SYM>;
    const size_t smem = p.lut_smem_bytes + (size_t)kColAsyncWarps * p.n_stages * p.stage_bytes;
    int grid = 0;
    cudaError_t e = persistent_grid(k, kColAsyncWarps * 32, smem, tune.ctas_per_sm, p.total_units, &grid);
    if (e != cudaSuccess) return e;
    k<<<grid, kColAsyncWarps * 32, smem, st>>>(p);
    return cudaGetLastError();
  }
  if (p.col_tile != 0) {  // register-resident variant (group 32 .. 512)
    const size_t smem = p.lut_smem_bytes;
    int grid = 0;
    cudaError_t e;
#define IWQ_COL_REG(RPL)                                                                  \
  do {                                                                                    \
    auto k = col_reg_kernel<DT, KIND, SYM, RPL>;                                          \
    e = persistent_grid(k, 256, smem, tune.ctas_per_sm, p.total_units, &grid);            \
    if (e == cudaSuccess) {                                                               \
      k<<<grid, 256, smem, st>>>(p);                                                      \
      e = cudaGetLastError();                                                             \
    }                                                                                     \
  } while (0)
    if (p.col_rpl == 16) IWQ_COL_REG(16);
    else if (p.col_rpl == 8) IWQ_COL_REG(8);
    else if (p.col_rpl == 4) IWQ_COL_REG(4);
    else e = cudaErrorInvalidValue;
#undef IWQ_COL_REG
    return e;
  }
  auto k = col_kernel<DT, KIND, SYM>;
  const size_t smem = 16384 + p.lut_smem_bytes;
  int grid = 0;
  cudaError_t e = persistent_grid(k, 256, smem, tune.ctas_per_sm, p.total_units, &grid);
  if (e != cudaSuccess) return e;
  k<<<grid, 256, smem, st>>>(p);
  return cudaGetLastError();
}

template <int DT, int KIND>
cudaError_t launch_col(const LaunchParams& p, bool sym, const LaunchTuning& tune, cudaStream_t st) {
  return sym ? launch_col_one<DT, KIND, true>(p, tune, st) : launch_col_one<DT, KIND, false>(p, tune, st);
}

template <int DT, int KIND, bool SYM, bool GIVEN>
static cudaError_t launch_seg_one(LaunchParams p, int variant, const LaunchTuning& tune, cudaStream_t st) {
  int grid = 0;
  if (variant == 1) {
    auto k = seg_direct_kernel<DT, KIND, SYM, GIVEN>;
    const size_t smem = p.lut_smem_bytes;
    cudaError_t e = persistent_grid(k, kConsumers, smem, tune.ctas_per_sm, p.total_units - p.unit_begin, &grid);
    if (e != cudaSuccess) return e;
    k<<<grid, kConsumers, smem, st>>>(p);
  } else if (variant == 0) {
    const size_t smem = kSmemHeader + p.lut_smem_bytes + (size_t)p.n_stages * kTileBytes;
    cudaError_t e = cudaSuccess;
#define IWQ_SEG4(LUTV, VPLV)                                                                                        \
  do {                                                                                                             \
    auto k = seg4_tma_kernel<DT, KIND, SYM, GIVEN, LUTV, VPLV>;                                                    \
    constexpr int threads = kConsumers * 4 / VPLV + 32;                                                            \
    e = persistent_grid(k, threads, smem, tune.ctas_per_sm, p.total_units - p.unit_begin, &grid);                  \
    if (e == cudaSuccess) k<<<grid, threads, smem, st>>>(p);                                                       \
  } while (0)
    constexpr bool kFp = KIND == KIND_FP;
    const int vpl = (KIND == KIND_INT_GENERIC) ? 4 : (int)p.seg_vpl;
    if (kFp && p.qc.lut_direct) {  // one value per bit pattern
      if (vpl == 16) IWQ_SEG4((kFp ? LUT_DIRECT : LUT_STEP), (KIND == KIND_INT_GENERIC ? 4 : 16));
      else if (vpl == 8) IWQ_SEG4((kFp ? LUT_DIRECT : LUT_STEP), (KIND == KIND_INT_GENERIC ? 4 : 8));
      else IWQ_SEG4((kFp ? LUT_DIRECT : LUT_STEP), 4);
    } else if (kFp && p.qc.lut_rep_shift > 0) {  // interleaved copies of the step table
      IWQ_SEG4((kFp ? LUT_REPL : LUT_STEP), 4);
    } else if (!kFp && vpl == 16) {
      IWQ_SEG4(LUT_STEP, ((kFp || KIND == KIND_INT_GENERIC) ? 4 : 16));
    } else if (!kFp && vpl == 8) {
      IWQ_SEG4(LUT_STEP, ((kFp || KIND == KIND_INT_GENERIC) ? 4 : 8));
    } else {
      IWQ_SEG4(LUT_STEP, 4);
    }
#undef IWQ_SEG4
    if (e != cudaSuccess) return e;
  } else {
    auto k = seg_tma_kernel<DT, KIND, SYM, GIVEN>;
    const size_t smem = kSmemHeader + p.lut_smem_bytes + (size_t)p.n_stages * kTileBytes;
    cudaError_t e = persistent_grid(k, kConsumers + 32, smem, tune.ctas_per_sm, p.total_units - p.unit_begin, &grid);
    if (e != cudaSuccess) return e;
    k<<<grid, kConsumers + 32, smem, st>>>(p);
  }
  return cudaGetLastError();
}

template <int DT, int KIND>
cudaError_t launch_seg(const LaunchParams& p, bool sym, bool given, int variant, const LaunchTuning& tune, cudaStream_t st) {
  if (sym) {
    return given ? launch_seg_one<DT, KIND, true, true>(p, variant, tune, st) : launch_seg_one<DT, KIND, true, false>(p, variant, tune, st);
  }
  return given ? launch_seg_one<DT, KIND, false, true>(p, variant, tune, st) : launch_seg_one<DT, KIND, false, false>(p, variant, tune, st);
}

template <int DT, int KIND, bool SYM>
static cudaError_t launch_row_one(LaunchParams p, int consumers, const LaunchTuning& tune, cudaStream_t st) {
  auto k = row_tma_kernel<DT, KIND, SYM>;
  const size_t smem = kSmemHeader + p.lut_smem_bytes + (size_t)p.n_stages * p.stage_bytes;
  int grid = 0;
  cudaError_t e = persistent_grid(k, consumers + 32, smem, tune.ctas_per_sm, p.total_units, &grid);
  if (e != cudaSuccess) return e;
  k<<<grid, consumers + 32, smem, st>>>(p);
  return cudaGetLastError();
}

template <int DT, int KIND>
cudaError_t launch_row(const LaunchParams& p, bool sym, int consumers, const LaunchTuning& tune, cudaStream_t st) {
  return sym ? launch_row_one<DT, KIND, true>(p, consumers, tune, st) : launch_row_one<DT, KIND, false>(p, consumers, tune, st);
}

#define IWQ_INSTANTIATE_KIND(DT, KIND)                                                                                       \
  template cudaError_t launch_seg<DT, KIND>(const LaunchParams&, bool, bool, int, const LaunchTuning&, cudaStream_t);       \
  template cudaError_t launch_row<DT, KIND>(const LaunchParams&, bool, int, const LaunchTuning&, cudaStream_t);   \
  template cudaError_t launch_col<DT, KIND>(const LaunchParams&, bool, const LaunchTuning&, cudaStream_t);            \
  template cudaError_t launch_col_tma<DT, KIND>(const LaunchParams&, const TensorMapPair&, const TensorMapPair*, bool, const LaunchTuning&, cudaStream_t); \
  template cudaError_t launch_fused<DT, KIND>(const LaunchParams&, const FusedParams&, bool, const LaunchTuning&, cudaStream_t, int*);

#endif  // IWQ_INSTANTIATE

}  // namespace iwq
