// iwq_host.cu -- C ABI of libiwq_b200.so (declared in include/iwq_b200.h): argument validation,
// launch-constant construction (including the per-binade minifloat table), work-list building and
// kernel dispatch.  No torch types; plain pointers, sizes and a cudaStream_t.
#include <cuda.h>  // CUtensorMap + enums only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint (no libcuda link)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "iwq_consts.h"
#include "iwq_launch.cuh"

namespace iwq {

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
static int fail_cuda(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return IWQ_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------
// tuning knobs (environment, read once): ring depth, CTAs per SM, forced kernel variant
// ------------------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return (s && *s) ? atoi(s) : dflt;
}
struct Knobs {
  int seg_stages = env_int("IWQ_SEG_STAGES", 4);
  int seg_ctas = env_int("IWQ_SEG_CTAS", 0);
  int row_ctas = env_int("IWQ_ROW_CTAS", 0);
  int row_smem_kb = env_int("IWQ_ROW_SMEM_KB", 0);  // per-CTA ring budget; 0 = automatic
  int variant = env_int("IWQ_VARIANT", -1);          // -1: use cfg->variant
  int row_warps = env_int("IWQ_ROW_WARPS", 0);       // consumer warps of the row kernels; 0 = automatic
  int col_tile = env_int("IWQ_COL_TILE", 0);         // quant_dim = 1: columns per tile (64/128/256), -1 = two-sweep kernel
  int col_ctas = env_int("IWQ_COL_CTAS", 0);
  int fused_batch_mb = env_int("IWQ_FUSED_BATCH_MB", 16);  // pertensor_fused_kernel: bytes of weights per batch
  int fused_lookahead = env_int("IWQ_FUSED_LOOKAHEAD", 2); // ... and how many batches the range pass runs ahead
  int row_tma_out = env_int("IWQ_ROW_TMA_OUT", 1);   // row kernel: bulk stores instead of STG.128 when the ring has >= 3 stages
  int wide_seg4 = env_int("IWQ_WIDE_SEG4", 1);       // 1: segments of 2048/4096/8192 elements go through seg4_tma_kernel
  int col_mode = env_int("IWQ_COL_MODE", 3);         // 3: CTA-wide ring of 128 x 256 tensor-map tiles (col_wide_kernel; group 32 / 64 / 128),
                                                     // 2: warp-private 2-D tensor-map boxes (col_tma_kernel), 0: warp-private cp.async ring
                                                     // (col_async_kernel), 1: register-resident (col_reg_kernel)
  int col_dbl_tile = env_int("IWQ_COL_DBL_TILE", 256); // col_double_kernel: columns per tile (256; 64 = second sweep from L2, measured 35-48 % SLOWER:
                                                     // the 16 per-lane scale computations per tile are then spread over 4x fewer elements)
  int col_box = env_int("IWQ_COL_BOX", 0);           // col_tma_kernel: columns per box (32 / 64 / 128 / 256); 0 = 4096 / group (8 KB boxes)
  int col_hint = env_int("IWQ_COL_HINT", 2);         // col_tma_kernel: L2 policy of the tensor loads (1 evict_first, 2 evict_last)
  int col_stages = env_int("IWQ_COL_STAGES", 0);     // ring depth of col_async_kernel (1..4); 0 = automatic
  int pt_slab_mb = env_int("IWQ_PT_SLAB_MB", 0);     // per-tensor: range pass -> apply pass per slab of whole matrices of at most this
                                                     // many MB (the apply pass then finds the slab in L2); 0 = one range pass over
                                                     // everything, then one apply pass
  int pt_reverse = env_int("IWQ_PT_REVERSE", 1);     // per-tensor apply pass walks its tiles from the end (most recently read first)
  int pt_hint = env_int("IWQ_PT_HINT", 0);           // LaunchParams::l2_hint bits of the per-tensor apply pass
  int seg_hint = env_int("IWQ_SEG_HINT", 2);         // LaunchParams::l2_hint bits of the one-pass segment kernels (2: bulk loads evict_last,
                                                     // +1.3 % on the in-place stream: the bulk store then hits the line it was loaded into)
  int mm_tma = env_int("IWQ_MM_TMA", 1);             // per-tensor range pass: 1 = bulk-copy ring (minmax_tma_kernel), 0 = 128-bit global loads
  int mm_stages = env_int("IWQ_MM_STAGES", 2);       // ... its ring depth (2 x 48 KB measured best: 7.26 TB/s vs 6.73 with 4, 6.9 for the global loads)
  int mm_hint = env_int("IWQ_MM_HINT", 0);           // ... L2 policy of its bulk loads (1 evict_first, 2 evict_last)
  int ws_cache = env_int("IWQ_WS_CACHE", 0);         // 1: skip the job-table upload when an identical table is already in the workspace
                                                     // (set by iwq_workspace_cache; the caller guarantees nobody else writes the workspace)
  int seg_vpl = env_int("IWQ_SEG_VPL", 0);           // seg4: 16-byte vectors per lane (4 / 8 / 16); 0 = automatic
  int seg_map = env_int("IWQ_SEG_MAP", 0);           // LaunchParams::tile_map of the seg4 kernels (1: contiguous run of tiles per CTA)
  int fp_mode = env_int("IWQ_FP_MODE", 2);           // seg4 minifloat kernel: 0 step table, 1 interleaved copies of it, 2 direct table
  int fp_rep = env_int("IWQ_FP_REP", 4);             // seg4 minifloat kernel: log2 of the interleaved copies of the step table
                                                     // in shared memory (4 = 16 copies: no bank conflicts); 0 = one copy
  int fp_rep_kb = env_int("IWQ_FP_REP_KB", 40);      // ... as long as the copies fit this many KB (else fewer copies)
  int fp_stages = env_int("IWQ_FP_STAGES", 0);       // ring depth of the seg4 minifloat kernel; 0 = IWQ_SEG_STAGES
  int host_copy_only = env_int("IWQ_HOST_COPY_ONLY", 0);  // diagnostics: the host-buffer pipeline without its kernels (copy ceiling)
};
static Knobs& knobs_mut() {
  static Knobs k;
  return k;
}
static const Knobs& knobs() { return knobs_mut(); }

// Ring depth of the segment kernels: the wanted depth, clamped so that header + table + ring fit the 227 KB of one CTA (and the
// per-stage range slots of seg4 fit the header).
static int fit_seg_stages(int want, size_t lut_bytes, size_t extra_bytes = 0) {
  int st = want < 2 ? 2 : want;
  if (st > kSeg4MaxStages) st = kSeg4MaxStages;
  while (st > 2 && kSmemHeader + lut_bytes + extra_bytes + (size_t)st * kTileBytes > 227u * 1024u) --st;
  return st;
}

constexpr int kSeg4AutoVpl = 4;  // default vectors per lane of seg4_tma_kernel (IWQ_SEG_VPL overrides)

static int effective_variant(const iwq_config_t* cfg) { return knobs().variant >= 0 ? knobs().variant : cfg->variant; }

static int build_consts(const iwq_config_t* cfg, QuantConsts* c, int* kind) {
  std::string msg;
  int rc = build_quant_consts(cfg, c, kind, &msg);
  if (rc) return fail(rc, msg);
  return IWQ_OK;
}

// ------------------------------------------------------------------------------------------------
// minifloat LUT cache: one small device buffer per (device, dtype, E, M, flavour), built on first use
// ------------------------------------------------------------------------------------------------
static std::mutex g_lut_mu;
static std::map<std::tuple<int, int, int, int, int, int, int, int, int, int, int, int>, FpStep*> g_luts;

static int get_fp_lut(const iwq_config_t* cfg, const QuantConsts& qc, cudaStream_t st, const FpStep** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail_cuda(e, "cudaGetDevice");
  const bool ap = qc.approx != 0;
  auto key = std::make_tuple(dev, cfg->dtype, cfg->exp_bits, cfg->mant_bits, qc.flavour_rne, qc.approx, ap ? qc.hi_start : 0,
                             ap ? qc.hi_field : 0, ap ? qc.tail_pad : 0, (int)qc.lut_shift, (int)qc.lut_lo_bits, (int)qc.lut_direct);
  std::lock_guard<std::mutex> lk(g_lut_mu);
  auto it = g_luts.find(key);
  if (it != g_luts.end()) {
    *out = it->second;
    return IWQ_OK;
  }
  if (!qc.lut_direct && qc.lut_n > 4096) return fail(IWQ_ERR_UNSUPPORTED, "minifloat step table too large");
  // approx == 2: [step table in info form, padded to 1 KB][value table, 16 KB] (build_double_tables)
  const size_t step_bytes = ((size_t)qc.lut_n * sizeof(FpStep) + 1023) & ~(size_t)1023;
  const size_t total = qc.lut_direct ? (((size_t)qc.lut_n * 2 + 7) & ~(size_t)7)
                     : qc.approx == 2 ? step_bytes + kDblValEntries * sizeof(uint16_t) : (size_t)qc.lut_n * sizeof(FpStep);
  std::vector<uint8_t> host(total, 0);
  FpStep* hsteps = reinterpret_cast<FpStep*>(host.data());
  bool ok = true;
  if (qc.lut_direct) build_direct_table(cfg->dtype, qc, reinterpret_cast<uint16_t*>(host.data()));
  else ok = qc.approx == 2 ? build_double_tables(cfg->dtype, qc, hsteps, reinterpret_cast<uint16_t*>(host.data() + step_bytes))
                           : build_step_table(cfg->dtype, qc, hsteps);
  if (!ok) return fail(IWQ_ERR_UNSUPPORTED, "minifloat format changes value more than once inside a half-step interval");
  FpStep* buf = nullptr;
  e = cudaMalloc((void**)&buf, total);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(minifloat step table)");
  e = cudaMemcpyAsync(buf, host.data(), total, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // one-time: later calls may come on any stream
  if (e != cudaSuccess) {
    cudaFree(buf);
    return fail_cuda(e, "minifloat step table upload");
  }
  g_luts[key] = buf;
  *out = buf;
  return IWQ_OK;
}

// ------------------------------------------------------------------------------------------------
// dispatch over (dtype, kind)
// ------------------------------------------------------------------------------------------------
static cudaError_t dispatch_seg(int dt, int kind, const LaunchParams& p, bool sym, bool given, int variant, const LaunchTuning& t,
                                cudaStream_t st) {
  if (dt == IWQ_F16) {
    switch (kind) {
      case KIND_INT_FAST: return launch_seg<IWQ_F16, KIND_INT_FAST>(p, sym, given, variant, t, st);
      case KIND_FP: return launch_seg<IWQ_F16, KIND_FP>(p, sym, given, variant, t, st);
      case KIND_BFP: return launch_seg<IWQ_F16, KIND_BFP>(p, sym, given, variant, t, st);
      default: return launch_seg<IWQ_F16, KIND_INT_GENERIC>(p, sym, given, variant, t, st);
    }
  }
  switch (kind) {
    case KIND_INT_FAST: return launch_seg<IWQ_BF16, KIND_INT_FAST>(p, sym, given, variant, t, st);
    case KIND_INT_WIDE: return launch_seg<IWQ_BF16, KIND_INT_WIDE>(p, sym, given, variant, t, st);
    case KIND_FP: return launch_seg<IWQ_BF16, KIND_FP>(p, sym, given, variant, t, st);
    case KIND_BFP: return launch_seg<IWQ_BF16, KIND_BFP>(p, sym, given, variant, t, st);
    default: return launch_seg<IWQ_BF16, KIND_INT_GENERIC>(p, sym, given, variant, t, st);
  }
}
static cudaError_t dispatch_row(int dt, int kind, const LaunchParams& p, bool sym, int consumers, const LaunchTuning& t, cudaStream_t st) {
  if (dt == IWQ_F16) {
    switch (kind) {
      case KIND_INT_FAST: return launch_row<IWQ_F16, KIND_INT_FAST>(p, sym, consumers, t, st);
      case KIND_FP: return launch_row<IWQ_F16, KIND_FP>(p, sym, consumers, t, st);
      case KIND_BFP: return launch_row<IWQ_F16, KIND_BFP>(p, sym, consumers, t, st);
      default: return launch_row<IWQ_F16, KIND_INT_GENERIC>(p, sym, consumers, t, st);
    }
  }
  switch (kind) {
    case KIND_INT_FAST: return launch_row<IWQ_BF16, KIND_INT_FAST>(p, sym, consumers, t, st);
    case KIND_INT_WIDE: return launch_row<IWQ_BF16, KIND_INT_WIDE>(p, sym, consumers, t, st);
    case KIND_FP: return launch_row<IWQ_BF16, KIND_FP>(p, sym, consumers, t, st);
    case KIND_BFP: return launch_row<IWQ_BF16, KIND_BFP>(p, sym, consumers, t, st);
    default: return launch_row<IWQ_BF16, KIND_INT_GENERIC>(p, sym, consumers, t, st);
  }
}

static cudaError_t dispatch_col(int dt, int kind, const LaunchParams& p, bool sym, const LaunchTuning& t, cudaStream_t st) {
  if (dt == IWQ_F16) {
    switch (kind) {
      case KIND_INT_FAST: return launch_col<IWQ_F16, KIND_INT_FAST>(p, sym, t, st);
      case KIND_FP: return launch_col<IWQ_F16, KIND_FP>(p, sym, t, st);
      case KIND_BFP: return launch_col<IWQ_F16, KIND_BFP>(p, sym, t, st);
      default: return launch_col<IWQ_F16, KIND_INT_GENERIC>(p, sym, t, st);
    }
  }
  switch (kind) {
    case KIND_INT_FAST: return launch_col<IWQ_BF16, KIND_INT_FAST>(p, sym, t, st);
    case KIND_INT_WIDE: return launch_col<IWQ_BF16, KIND_INT_WIDE>(p, sym, t, st);
    case KIND_FP: return launch_col<IWQ_BF16, KIND_FP>(p, sym, t, st);
    case KIND_BFP: return launch_col<IWQ_BF16, KIND_BFP>(p, sym, t, st);
    default: return launch_col<IWQ_BF16, KIND_INT_GENERIC>(p, sym, t, st);
  }
}

static cudaError_t dispatch_fused(int dt, int kind, const LaunchParams& p, const FusedParams& f, bool sym, const LaunchTuning& t, cudaStream_t st, int* grid) {
  if (dt == IWQ_F16) {
    switch (kind) {
      case KIND_INT_FAST: return launch_fused<IWQ_F16, KIND_INT_FAST>(p, f, sym, t, st, grid);
      case KIND_FP: return launch_fused<IWQ_F16, KIND_FP>(p, f, sym, t, st, grid);
      case KIND_BFP: return cudaErrorInvalidValue;
      default: return launch_fused<IWQ_F16, KIND_INT_GENERIC>(p, f, sym, t, st, grid);
    }
  }
  switch (kind) {
    case KIND_INT_FAST: return launch_fused<IWQ_BF16, KIND_INT_FAST>(p, f, sym, t, st, grid);
    case KIND_INT_WIDE: return launch_fused<IWQ_BF16, KIND_INT_WIDE>(p, f, sym, t, st, grid);
    case KIND_FP: return launch_fused<IWQ_BF16, KIND_FP>(p, f, sym, t, st, grid);
    case KIND_BFP: return cudaErrorInvalidValue;
    default: return launch_fused<IWQ_BF16, KIND_INT_GENERIC>(p, f, sym, t, st, grid);
  }
}

// quant_dim = 1 (segments down the rows): one launch of col_kernel over all matrices
// Tile shape of col_wide_kernel.  Shared memory = header + minifloat table + partial ranges + the published constants of both
// consumer groups + the ring; candidates in order of preference (largest bulk copies first): 128 x 256 (64 KB), 64 x 256 (32 KB,
// group <= 64), 128 x 128 and 64 x 128 (256-byte rows); the first one whose ring holds at least three stages wins.
struct ColWideShape {
  uint32_t rows, cols, stages;
};
static size_t col_wide_fixed(int64_t g, int kind, bool sym, uint32_t lut_bytes, uint32_t tr, uint32_t cc) {
  const size_t words = kind == KIND_INT_GENERIC ? 16 : (kind == KIND_BFP ? 4 : 8);  // PreparedPair<..>::kWords
  return (size_t)kColWideHeader + lut_bytes + (sym ? kColWidePart / 2 : kColWidePart) + (size_t)kColWideGroups * (tr / (uint32_t)g) * (cc / 2) * words * 4;
}
static bool col_wide_shape(int64_t g, int kind, bool sym, uint32_t lut_bytes, ColWideShape* out, bool full_tile_only = false) {
  const uint32_t cand[4][2] = {{128, 256}, {64, 256}, {128, 128}, {64, 128}};
  for (int i = 0; i < 4; ++i) {
    const uint32_t tr = cand[i][0], cc = cand[i][1];
    if (tr % (uint32_t)g != 0) continue;
    if (knobs().col_box == 64 && tr != 64 && g <= 64) continue;   // experiments: force 64-row tiles
    if (knobs().col_box == 128 && cc != 128) continue;            // experiments: force 128-column tiles
    if (full_tile_only && (tr != 128 || cc != 256)) continue;
    const size_t fixed = col_wide_fixed(g, kind, sym, lut_bytes, tr, cc), tile = (size_t)tr * cc * 2;
    int stg = knobs().col_stages > 0 ? knobs().col_stages : (tile >= 65536 ? 3 : 6);
    if (stg > 8) stg = 8;  // ColWideCtx ctxs[8]
    while (stg > 1 && fixed + (size_t)stg * tile > 227u * 1024u) --stg;
    if (stg >= 3 || (knobs().col_stages > 0 && stg >= knobs().col_stages)) {
      out->rows = tr;
      out->cols = cc;
      out->stages = (uint32_t)stg;
      return true;
    }
  }
  return false;
}

static int run_col(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, LaunchParams& p, int kind, void* workspace, size_t ws_bytes, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// work-list construction
// ------------------------------------------------------------------------------------------------
enum Path { PATH_SEG, PATH_ROW };

struct Plan {
  Path path;
  int lps, vpl;            // seg path
  uint32_t seg_elems;      // elements per segment (group, or cols)
  uint32_t stage_bytes;    // row path
  int stages, consumers;   // row path
};

static bool seg_shape(uint64_t seg_elems, int* lps, int* vpl, bool seg4_ok, bool minifloat) {
  // seg4 kernel only: segments of 2048 / 4096 elements span 2 / 4 warps of a 16 KB tile.  8192 (8 warps) works too: for
  // INT the row kernel is faster there (6073 vs 5924 GB/s on the 70B shapes; IWQ_WIDE_SEG4=2 forces it), for the
  // minifloat element stage the row kernel's 2-4 consumer warps are too few and seg4 wins
  if (seg4_ok && (seg_elems == 2048 || seg_elems == 4096 || (seg_elems == 8192 && (minifloat || knobs().wide_seg4 >= 2)))) {
    *lps = (int)(seg_elems / 32);
    *vpl = 1;
    return true;
  }
  // group = 8 * lps * vpl with lps in {1,2,4,8,16,32}; vpl in {1,2,4} only once lps == 32
  if (seg_elems < 8 || seg_elems > 1024 || (seg_elems & (seg_elems - 1))) return false;
  uint64_t vecs = seg_elems / 8;
  if (vecs <= 32) {
    *lps = (int)vecs;
    *vpl = 1;
  } else {
    *lps = 32;
    *vpl = (int)(vecs / 32);
  }
  return true;
}

static int check_matrix(const iwq_matrix_t* m) {
  if (!m->w || !m->out) return fail(IWQ_ERR_INVALID, "null weight/output pointer");
  if (m->rows <= 0 || m->cols <= 0) return fail(IWQ_ERR_INVALID, "rows and cols must be positive");
  if (((uintptr_t)m->w | (uintptr_t)m->out) & 15) return fail(IWQ_ERR_ALIGN, "weight/output pointers must be 16-byte aligned");
  if ((uint64_t)m->rows * (uint64_t)m->cols * 2 >= (1ull << 45)) return fail(IWQ_ERR_INVALID, "matrix too large");
  return IWQ_OK;
}

// decide the kernel path for one (cfg.group, cols); all matrices of a batched call must agree
static int make_plan(const iwq_config_t* cfg, int64_t cols, uint32_t lut_bytes, Plan* pl) {
  const int g = cfg->group;
  uint64_t seg;
  if (g > 0) {
    if (cols % g != 0) return fail(IWQ_ERR_INDIVISIBLE, "cols is not divisible by the group size");
    seg = (uint64_t)g;
  } else if (g == IWQ_GROUP_PER_CHANNEL) {
    seg = (uint64_t)cols;
  } else if (g == IWQ_GROUP_PER_TENSOR) {
    seg = 0;
  } else {
    return fail(IWQ_ERR_INVALID, "Invalid w_group_size");
  }
  memset(pl, 0, sizeof(*pl));
  pl->seg_elems = (uint32_t)seg;
  if (g == IWQ_GROUP_PER_TENSOR) {
    pl->path = PATH_SEG;
    pl->lps = 1;
    pl->vpl = 1;
    return IWQ_OK;
  }
  const int v = effective_variant(cfg);
  const bool minifloat = cfg->scheme == IWQ_SCHEME_FP_CODEC || cfg->scheme == IWQ_SCHEME_FP_RNE;
  if (seg_shape(seg, &pl->lps, &pl->vpl, (v < 1 || v > 2) && knobs().wide_seg4 != 0, minifloat)) {
    pl->path = PATH_SEG;
    return IWQ_OK;
  }
  if (seg % 8 != 0) return fail(IWQ_ERR_UNSUPPORTED, "segment length must be a multiple of 8 elements (16-byte vectors)");
  const uint64_t row_bytes = seg * 2;
  const uint32_t stage = (uint32_t)((row_bytes + 127) & ~127ull);
  const uint64_t max_ring = 200 * 1024 - (uint64_t)lut_bytes;
  if (2ull * stage > max_ring) return fail(IWQ_ERR_UNSUPPORTED, "row longer than 51200 elements: not implemented for per-channel granularity");
  pl->path = PATH_ROW;
  pl->stage_bytes = stage;
  const uint64_t vecs = row_bytes / 16;
  // Small CTAs, many per SM: the kernel pays one named barrier per row, which other resident CTAs hide.
  // Measured on B200 (profiles/r01_row_kernel_sweep.txt): 2 consumer warps x ~28-36 KB ring for 8 KB rows,
  // 4 warps x 72 KB for 16-56 KB rows.
  int warps = vecs <= 128 ? 1 : (int)(vecs / 256);
  if (warps < (vecs <= 128 ? 1 : 2)) warps = vecs <= 128 ? 1 : 2;
  if (warps > 4) warps = 4;
  if (minifloat) {
    // the table-lookup element stage needs many more warps to hide its shared-memory latency: measured on the 70B
    // shapes (FP4 per-channel) 3865 GB/s with 4 consumer warps, 4497 with 8, 4757 with 16 -- INT is the other way round
    warps = (int)(vecs / 128);
    if (warps < 2) warps = 2;
    if (warps > 16) warps = 16;
  }
  if (knobs().row_warps > 0) warps = knobs().row_warps > 16 ? 16 : knobs().row_warps;
  pl->consumers = 32 * warps;
  uint64_t budget = knobs().row_smem_kb > 0 ? (uint64_t)knobs().row_smem_kb * 1024 : 72 * 1024;
  if (budget > max_ring) budget = max_ring;
  int st = (int)(budget / stage);
  if (st < 2) st = 2;
  // three stages let the row kernel send its results through bulk stores (one stage stays parked behind its store)
  if (st == 2 && knobs().row_tma_out != 0 && stage >= 20 * 1024 && 3ull * stage <= 110 * 1024 && 3ull * stage <= max_ring) st = 3;
  if (st > 4 && knobs().row_smem_kb <= 0) st = 4;
  if (st > kMaxStages) st = kMaxStages;
  pl->stages = st;
  // rings of which exactly two fit an SM (24-36 KB rows, e.g. the 13824-wide rows of LLaMA-2-13B): 8 consumer warps per CTA instead of
  // 4 -- 6039 against 5654 GB/s (profiles/r02_sweep_row24.jsonl); with three or more resident CTAs (11008-wide rows) and with a single one
  // (28672-wide rows) 4 warps stay ahead
  if (!minifloat && knobs().row_warps <= 0 && warps == 4) {
    const uint64_t cta_smem = kSmemHeader + (uint64_t)lut_bytes + (uint64_t)st * stage;
    if ((227ull * 1024) / cta_smem == 2) pl->consumers = 32 * 8;
  }
  return IWQ_OK;
}

static Job make_job(const iwq_matrix_t* m, uint32_t first_unit) {
  Job j;
  memset(&j, 0, sizeof(j));
  j.src = (const uint8_t*)m->w;
  j.dst = (uint8_t*)m->out;
  j.scales = (uint8_t*)m->scales;
  j.zeros = (uint8_t*)m->zeros;
  j.codes = (uint8_t*)m->codes;
  j.bytes = (uint64_t)m->rows * (uint64_t)m->cols * 2;
  j.first_unit = first_unit;
  j.row_bytes = (uint32_t)(m->cols * 2);
  return j;
}


// Fill LaunchParams for `n` matrices; `jobs_host` receives the table (uploaded by the caller when n > 1).
static int fill_params(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, const Plan& pl, LaunchParams* p, std::vector<Job>* jobs_host) {
  jobs_host->clear();
  uint64_t units = 0;
  for (int i = 0; i < n; ++i) {
    int rc = check_matrix(&ms[i]);
    if (rc) return rc;
    if (units >= 0xFFFFFFFFull) return fail(IWQ_ERR_INVALID, "too much work for one launch");
    Job j = make_job(&ms[i], (uint32_t)units);
    if (cfg->group == IWQ_GROUP_PER_TENSOR && (j.bytes & 15))
      return fail(IWQ_ERR_UNSUPPORTED, "per-tensor granularity needs rows * cols to be a multiple of 8 (16-byte vectors); pad the tensor with one of its own values");
    if (pl.path == PATH_SEG) {
      units += (j.bytes + kTileBytes - 1) / kTileBytes;
    } else {
      j.row_bytes = pl.seg_elems * 2;  // a "row" is one segment (per-channel row, or an odd-sized group)
      units += j.bytes / j.row_bytes;
    }
    jobs_host->push_back(j);
  }
  if (units >= 0xFFFFFFFFull) return fail(IWQ_ERR_INVALID, "too much work for one launch");
  p->n_jobs = n;
  p->total_units = (uint32_t)units;
  p->single = (*jobs_host)[0];
  p->jobs = nullptr;
  p->lps = pl.lps;
  p->vpl = pl.vpl;
  p->stage_bytes = pl.stage_bytes;
  p->n_stages = pl.path == PATH_SEG ? fit_seg_stages(knobs().seg_stages, p->lut_smem_bytes, p->lut_extra_bytes) : pl.stages;
  // measured (profiles/README.md): bulk stores pay off for long rows (22 KB: +3.4 %, 32 KB: +2 %), not for 10..16 KB ones
  p->row_tma_out = (pl.path == PATH_ROW && pl.stages >= 3 && knobs().row_tma_out != 0 && (knobs().row_tma_out > 1 || pl.seg_elems * 2ull >= 20 * 1024)) ? 1u : 0u;
  return IWQ_OK;
}

struct WsCacheEntry {
  void* ws;
  cudaStream_t st;
  std::vector<uint8_t> bytes;
};
static std::mutex g_ws_mu;
static std::vector<WsCacheEntry> g_ws_cache;

static int upload_jobs(const std::vector<Job>& jobs, void* workspace, size_t ws_bytes, LaunchParams* p, cudaStream_t st) {
  if (jobs.size() <= 1) return IWQ_OK;
  const size_t need = jobs.size() * sizeof(Job);
  if (!workspace || ws_bytes < need) return fail(IWQ_ERR_INVALID, "workspace too small (see iwq_workspace_bytes)");
  if ((uintptr_t)workspace & 15) return fail(IWQ_ERR_ALIGN, "workspace must be 16-byte aligned");
  // An in-stream H2D copy has to wait for the previous kernel of the stream, and the next kernel for the copy: ~35 us of idle GPU
  // between back-to-back launches (0.6 % of a 5.4 ms pass).  When the caller declares that the workspace is touched by nobody but
  // this library (iwq_workspace_cache(1); the Python layer owns its workspaces and does), a job table that is byte-identical to the
  // one this stream last uploaded to this workspace is not uploaded again.
  if (knobs().ws_cache != 0) {
    std::lock_guard<std::mutex> lk(g_ws_mu);
    auto& cache = g_ws_cache;  // a handful of (workspace, stream) pairs: one per device and thread in practice
    using Entry = WsCacheEntry;
    Entry* hit = nullptr;
    for (auto& en : cache)
      if (en.ws == workspace && en.st == st) hit = &en;
    if (hit && hit->bytes.size() == need && memcmp(hit->bytes.data(), jobs.data(), need) == 0) {
      p->jobs = (const Job*)workspace;
      return IWQ_OK;
    }
    cudaError_t e = cudaMemcpyAsync(workspace, jobs.data(), need, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(job table)");
    // any other (workspace, stream) entry that overlaps this workspace is stale now
    for (size_t i = 0; i < cache.size();) {
      if (cache[i].ws == workspace && cache[i].st != st) cache.erase(cache.begin() + (long)i); else ++i;
    }
    if (!hit) {
      if (cache.size() >= 64) cache.clear();
      cache.push_back(Entry{workspace, st, {}});
      hit = &cache.back();
    }
    hit->bytes.assign(reinterpret_cast<const uint8_t*>(jobs.data()), reinterpret_cast<const uint8_t*>(jobs.data()) + need);
    p->jobs = (const Job*)workspace;
    return IWQ_OK;
  }
  cudaError_t e = cudaMemcpyAsync(workspace, jobs.data(), need, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(job table)");
  p->jobs = (const Job*)workspace;
  return IWQ_OK;
}

static int run_minmax(const iwq_config_t* cfg, const LaunchParams& base, float* range, cudaStream_t st, bool init = true, bool strided = false) {
  LaunchParams p = base;
  p.ranges_out = range;
  p.l2_hint = strided ? 8u : 0u;
  if (init) {
    const int n2 = 2 * p.n_jobs;
    range_init_kernel<<<(n2 + 255) / 256, 256, 0, st>>>(range, n2);
  }
  int grid = 0;
  cudaError_t e;
  const bool sym = p.qc.symmetric != 0;
  const bool tma = knobs().mm_tma != 0;
  if (tma) {
    p.n_stages = fit_seg_stages(knobs().mm_stages, 0);
    p.tile_map = strided ? 0u : 1u;
    p.l2_hint = (uint32_t)knobs().mm_hint & 3u;
  }
  const size_t tma_smem = kSmemHeader + (size_t)p.n_stages * kTileBytes;
#define IWQ_MM(DT, SYM)                                                                \
  do {                                                                                 \
    if (tma) {                                                                         \
      auto k = minmax_tma_kernel<DT, SYM>;                                             \
      e = persistent_grid(k, kConsumers + 32, tma_smem, 0, p.total_units - p.unit_begin, &grid); \
      if (e == cudaSuccess) {                                                          \
        k<<<grid, kConsumers + 32, tma_smem, st>>>(p);                                 \
        e = cudaGetLastError();                                                        \
      }                                                                                \
      break;                                                                           \
    }                                                                                  \
    auto k = minmax_kernel<DT, SYM>;                                                   \
    e = persistent_grid(k, kConsumers, 0, 0, p.total_units - p.unit_begin, &grid);     \
    if (e == cudaSuccess) {                                                            \
      k<<<grid, kConsumers, 0, st>>>(p);                                               \
      e = cudaGetLastError();                                                          \
    }                                                                                  \
  } while (0)
  if (cfg->dtype == IWQ_F16) {
    if (sym) IWQ_MM(IWQ_F16, true); else IWQ_MM(IWQ_F16, false);
  } else {
    if (sym) IWQ_MM(IWQ_BF16, true); else IWQ_MM(IWQ_BF16, false);
  }
#undef IWQ_MM
  if (e != cudaSuccess) return fail_cuda(e, "minmax kernel launch");
  return IWQ_OK;
}

static int run_batched(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, void* workspace, size_t ws_bytes, const float* given_range,
                       float* range_ws, int32_t* status, cudaStream_t st, bool minmax_only) {
  if (!cfg || !ms || n < 1) return fail(IWQ_ERR_INVALID, "null config/matrices or n < 1");
  LaunchParams p;
  memset(&p, 0, sizeof(p));
  int kind = 0;
  int rc = build_consts(cfg, &p.qc, &kind);
  if (rc) return rc;
  auto bind_lut = [&]() -> int {  // (re)build / fetch the device step table for the current p.qc
    int r = get_fp_lut(cfg, p.qc, st, &p.lut);
    if (r) return r;
    p.lut_smem_bytes = p.qc.lut_direct ? (uint32_t)(((size_t)p.qc.lut_n * 2 + 1023) & ~(size_t)1023)
                                       : (uint32_t)((((size_t)p.qc.lut_n * sizeof(FpStep)) << p.qc.lut_rep_shift) + 1023) & ~(uint32_t)1023;
    p.lut_extra_bytes = p.qc.approx == 2 ? (uint32_t)(kDblValEntries * sizeof(uint16_t)) : 0u;
    return IWQ_OK;
  };
  if (kind == KIND_FP) {
    rc = bind_lut();
    if (rc) return rc;
  }
  if (p.qc.approx == 2 && cfg->quant_dim == 0) {
    // double-approximate decode along the rows (quant_linear.py:484-488): groups of four = the same offset of four
    // consecutive segments (of a row, or of neighbouring rows when a row holds fewer than four)
    if (given_range || minmax_only) return fail(IWQ_ERR_INVALID, "range entry points are for per-tensor granularity");
    const int g = cfg->group;
    if (g != 32 && g != 64 && g != 128 && g != 256) return fail(IWQ_ERR_UNSUPPORTED, "double-approximate decode with quant_dim = 0 is implemented for group 32 / 64 / 128 / 256");
    std::vector<Job> jobs;
    uint64_t units = 0;
    const uint64_t quads_per_unit = 8ull * (256 / g);
    for (int i = 0; i < n; ++i) {
      rc = check_matrix(&ms[i]);
      if (rc) return rc;
      if (ms[i].codes) return fail(IWQ_ERR_UNSUPPORTED, "the double-approximate decode does not emit codes");
      if (ms[i].cols % g != 0) return fail(IWQ_ERR_INDIVISIBLE, "cols is not divisible by the group size");
      // four consecutive segments (row-major segment order) are 4 * group contiguous elements whatever the row length
      if (((uint64_t)ms[i].rows * (uint64_t)(ms[i].cols / g)) % 4 != 0)
        return fail(IWQ_ERR_UNSUPPORTED, "double-approximate decode needs the number of segments to be a multiple of 4");
      Job j = make_job(&ms[i], (uint32_t)units);
      units += (j.bytes / (8ull * g) + quads_per_unit - 1) / quads_per_unit;
      if (units >= 0xFFFFFFFFull) return fail(IWQ_ERR_INVALID, "too much work for one launch");
      jobs.push_back(j);
    }
    p.n_jobs = n;
    p.total_units = (uint32_t)units;
    p.single = jobs[0];
    p.jobs = nullptr;
    p.col_group = (uint32_t)g;
    p.status = status;
    rc = upload_jobs(jobs, workspace, ws_bytes, &p, st);
    if (rc) return rc;
    cudaError_t e2 = cfg->dtype == IWQ_F16 ? launch_seg_double<IWQ_F16>(p, st) : launch_seg_double<IWQ_BF16>(p, st);
    if (e2 != cudaSuccess) return fail_cuda(e2, "double-approximate (quant_dim 0) kernel launch");
    return IWQ_OK;
  }
  if (cfg->quant_dim == 1 && cfg->group != IWQ_GROUP_PER_TENSOR) {
    if (given_range || minmax_only) return fail(IWQ_ERR_INVALID, "range entry points are for per-tensor granularity");
    p.status = status;
    return run_col(cfg, ms, n, p, kind, workspace, ws_bytes, st);
  }
  if (cfg->quant_dim != 0 && cfg->quant_dim != 1) return fail(IWQ_ERR_INVALID, "quant_dim must be 0 or 1");
  Plan pl;
  rc = make_plan(cfg, ms[0].cols, p.lut_smem_bytes, &pl);
  if (rc) return rc;
  for (int i = 1; i < n; ++i) {
    Plan pi;
    rc = make_plan(cfg, ms[i].cols, p.lut_smem_bytes, &pi);
    if (rc) return rc;
    if (pi.path != pl.path || pi.lps != pl.lps || pi.vpl != pl.vpl || pi.stage_bytes != pl.stage_bytes)
      return fail(IWQ_ERR_UNSUPPORTED, "matrices of one batched call must share the segment shape (batch per distinct cols for per-channel)");
  }
  std::vector<Job> jobs;
  rc = fill_params(cfg, ms, n, pl, &p, &jobs);
  if (rc) return rc;
  rc = upload_jobs(jobs, workspace, ws_bytes, &p, st);
  if (rc) return rc;
  p.status = status;
  if (kind == KIND_FP && knobs().row_tma_out < 2) p.row_tma_out = 0;  // the minifloat rows measured 5 % slower with bulk stores
  const bool sym = p.qc.symmetric != 0;
  const bool per_tensor = cfg->group == IWQ_GROUP_PER_TENSOR;
  LaunchTuning tune;
  tune.stages = p.n_stages;
  cudaError_t e;
  // kernel variant: 0 = seg4 (TMA in + TMA out, 4 vectors per lane; needs group >= 32), 1 = direct global
  // loads/stores, 2 = TMA in + register stores (one vector per lane-iteration; also serves group 8 / 16)
  int variant = effective_variant(cfg);
  if (variant < 0 || variant > 2) variant = 0;
  if (variant == 0 && !per_tensor && pl.path == PATH_SEG && pl.seg_elems < 32) variant = 2;
  if (variant == 0 && pl.path == PATH_SEG) {
    p.seg_vpl = 4;
    p.lps = per_tensor ? 1 : (int)(pl.seg_elems / 32);
    p.vpl = 1;
    if (kind == KIND_FP && knobs().fp_mode == 2 && p.qc.approx != 2) {
      // seg4 minifloat kernel, DIRECT table: one 16-bit value per bit pattern (2 loads + 1 pack per pair instead of
      // 2 loads + 3 packs + compare + select), as long as it fits the shared-memory budget
      QuantConsts q = p.qc;
      set_lut_direct(cfg->dtype, &q);
      if ((size_t)q.lut_n * 2 <= (size_t)knobs().fp_rep_kb * 1024) {
        p.qc = q;
        rc = bind_lut();
        if (rc) return rc;
      }
      tune.stages = p.n_stages = fit_seg_stages(knobs().fp_stages >= 2 ? knobs().fp_stages : knobs().seg_stages, p.lut_smem_bytes, p.lut_extra_bytes);
    } else if (kind == KIND_FP && knobs().fp_mode == 1 && knobs().fp_rep > 0) {
      // seg4 minifloat kernel: interleaved copies of the step table (no shared-memory bank conflicts in the lookups)
      const QuantConsts base = p.qc;
      for (int rs = knobs().fp_rep > 4 ? 4 : knobs().fp_rep; rs > 0; --rs) {
        QuantConsts q = base;
        set_lut_replication(cfg->dtype, &q, rs);
        if ((((size_t)q.lut_n * sizeof(FpStep)) << rs) <= (size_t)knobs().fp_rep_kb * 1024) {
          p.qc = q;
          break;
        }
      }
      if (p.qc.lut_rep_shift > 0) {
        rc = bind_lut();
        if (rc) return rc;
      }
      tune.stages = p.n_stages = fit_seg_stages(knobs().fp_stages >= 2 ? knobs().fp_stages : knobs().seg_stages, p.lut_smem_bytes, p.lut_extra_bytes);
    }
    // vectors per lane: more of a segment per lane = fewer lanes repeating the segment's parameter arithmetic.  Segments that span
    // warps (2048+ elements) and the step-table minifloat forms stay at 4.
    int vpl = 4;
    {
      // measured (tools/r2_exp16.sh / r2_exp17.sh, fresh data): the direct-table minifloat stage gains 5-13 % from 8 vectors per lane
      // (FP4 E2M1 5.72 -> 6.39 TB/s, FP6 E3M2 5.60 -> 6.35, FP8 E4M3 5.45 -> 5.72); the INT / BFP stages are HBM-bound either way in
      // bench-length runs (6.23-6.34 vs 6.17-6.32 TB/s) and keep 4 (in minute-long runs under the power cap 8 is ~7 % faster)
      // INT stage, bench regime, fresh process per measurement, alternating (profiles/r02_bench20_vpl.txt): group 128 -- 16 vectors
      // per lane (a lane owns a whole segment: no shuffle, no repeated parameters) 6327-6382 vs 6270-6280 GB/s at 4; group 256 -- 8:
      // 6270-6298 vs 6041-6068; group 64 -- 4: 6350-6386 vs 6063-6089 at 8; BFP and the per-tensor apply pass: no gain, 4
      int auto_vpl = kind == KIND_FP ? 8 : kSeg4AutoVpl;
      if ((kind == KIND_INT_FAST || kind == KIND_INT_WIDE) && !per_tensor) auto_vpl = pl.seg_elems == 128 ? 16 : (pl.seg_elems >= 256 ? 8 : 4);
      const int want = knobs().seg_vpl == 8 || knobs().seg_vpl == 16 ? knobs().seg_vpl : (knobs().seg_vpl == 4 ? 4 : auto_vpl);
      const bool fp_step = kind == KIND_FP && !p.qc.lut_direct;
      if (kind != KIND_INT_GENERIC && !fp_step) {
        for (int v = want; v > 4; v >>= 1) {
          // (lanes per segment = seg / (8 v) must stay within one warp: 2048 / 4096-element per-channel rows qualify at v = 16)
          if (per_tensor || (pl.seg_elems % (8u * (uint32_t)v) == 0 && pl.seg_elems / (8u * (uint32_t)v) <= 32u)) {
            vpl = v;
            break;
          }
        }
      }
    }
    p.seg_vpl = (uint32_t)vpl;
    p.lps = per_tensor ? 1 : (int)(pl.seg_elems / (8u * (uint32_t)vpl));

  }
  if (per_tensor) {
    const float* rng = given_range;
    if (!rng && !range_ws) return fail(IWQ_ERR_INVALID, "per-tensor granularity needs a device range buffer of 2 floats per matrix");
    tune.ctas_per_sm = knobs().seg_ctas;
    const bool seg4 = variant == 0;
    if (!rng && !minmax_only && seg4 && knobs().pt_slab_mb > 0) {
      // SLAB ORDER: range pass and apply pass alternate over slabs of whole matrices that fit L2, so that the apply pass
      // finds its slab in L2 (4 B/weight of HBM traffic instead of 6); a matrix larger than the slab is a slab of its own
      // (its apply pass still starts with the tiles the range pass read last).  quant_linear.py:898-900
      const uint64_t slab_bytes = (uint64_t)knobs().pt_slab_mb << 20;
      const uint32_t end_units = p.total_units;
      const int n2 = 2 * n;
      range_init_kernel<<<(n2 + 255) / 256, 256, 0, st>>>(range_ws, n2);
      p.ranges = range_ws;
      e = cudaSuccess;
      for (int j0 = 0; j0 < n && e == cudaSuccess;) {
        int j1 = j0 + 1;
        uint64_t acc = jobs[j0].bytes;
        while (j1 < n && acc + jobs[j1].bytes <= slab_bytes) acc += jobs[j1++].bytes;
        p.unit_begin = jobs[j0].first_unit;
        p.total_units = j1 < n ? jobs[j1].first_unit : end_units;
        p.reverse = 0;
        rc = run_minmax(cfg, p, range_ws, st, false, true);
        if (rc) return rc;
        p.reverse = 1;
        p.l2_hint = (uint32_t)knobs().pt_hint & 23u;
        e = dispatch_seg(cfg->dtype, kind, p, sym, true, variant, tune, st);
        j0 = j1;
      }
    } else {
      if (!rng) {
        rc = run_minmax(cfg, p, range_ws, st);
        if (rc) return rc;
        rng = range_ws;
      }
      if (minmax_only) return IWQ_OK;
      p.ranges = rng;
      p.reverse = (seg4 && knobs().pt_reverse != 0) ? 1u : 0u;
      p.l2_hint = (uint32_t)knobs().pt_hint & 23u;
      p.tile_map = (uint32_t)knobs().seg_map;
      e = dispatch_seg(cfg->dtype, kind, p, sym, true, variant, tune, st);
    }
  } else if (pl.path == PATH_SEG) {
    tune.ctas_per_sm = knobs().seg_ctas;
    p.l2_hint = (uint32_t)knobs().seg_hint & 23u;
    p.tile_map = (uint32_t)knobs().seg_map;
    e = dispatch_seg(cfg->dtype, kind, p, sym, false, variant, tune, st);
  } else {
    tune.ctas_per_sm = knobs().row_ctas;
    e = dispatch_row(cfg->dtype, kind, p, sym, pl.consumers, tune, st);
  }
  if (e != cudaSuccess) return fail_cuda(e, "fake-quant kernel launch");
  return IWQ_OK;
}

// Tile shape for quant_dim = 1 segments of `g` rows.  lpr lanes cover a strip row (16 * lpr bytes).
// Default (col_async_kernel): tile = 32 * lpr columns (4 warps), *rpl = 0, ring of *stages x g * 16 * lpr bytes
// per warp.  IWQ_COL_MODE=1 (col_reg_kernel): tile = 64 * lpr columns (8 warps), *rpl = g * lpr / 32 vectors per
// lane held in registers.  Returns 0 for the two-sweep col_kernel (other group sizes, per-channel columns).
static uint32_t col_tile_for(int64_t g, int* rpl, int* stages) {
  if (knobs().col_tile < 0) return 0;
  if (g != 32 && g != 64 && g != 128 && g != 256 && g != 512) return 0;
  uint32_t lpr = g <= 128 ? 4 : g == 256 ? 2 : 1;
  if (knobs().col_tile == 128 && g <= 128) lpr = 2;
  if (knobs().col_tile == 64 && g <= 256) lpr = 1;
  if (knobs().col_mode == 1) {
    const int64_t r = g * lpr / 32;
    if (r != 4 && r != 8 && r != 16) return 0;
    *rpl = (int)r;
    return 64u * lpr;
  }
  *rpl = 0;
  // measured on B200 (profiles/r01_col_kernel_notes.md): a single stage with 4..6 CTAs per SM beats a deeper
  // per-warp ring with fewer resident warps
  int st = knobs().col_stages > 0 ? knobs().col_stages : 1;
  if (st > 4) st = 4;
  *stages = st;
  return 32u * lpr;
}

// 2-D tensor map of a row-major [rows, cols] 16-bit matrix with a box of [box_rows x box_cols] elements (col_tma_kernel)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_map(TensorMap2D* out, const void* base, int64_t rows, int64_t cols, uint32_t box_rows, uint32_t box_cols) {
  static EncodeTiledFn fn = nullptr;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!fn) {
      void* f = nullptr;
      cudaDriverEntryPointQueryResult qr;
      cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
      if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !f) return fail(IWQ_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
      fn = (EncodeTiledFn)f;
    }
  }
  static_assert(sizeof(TensorMap2D) == sizeof(CUtensorMap), "tensor map size");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  alignas(64) CUtensorMap m;
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(IWQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
  memcpy(out, &m, sizeof(m));
  return IWQ_OK;
}

static cudaError_t dispatch_col_tma(int dt, int kind, const LaunchParams& p, const TensorMapPair& single, const TensorMapPair* maps, bool sym, const LaunchTuning& t,
                                    cudaStream_t st) {
  if (dt == IWQ_F16) {
    switch (kind) {
      case KIND_INT_FAST: return launch_col_tma<IWQ_F16, KIND_INT_FAST>(p, single, maps, sym, t, st);
      case KIND_FP: return launch_col_tma<IWQ_F16, KIND_FP>(p, single, maps, sym, t, st);
      case KIND_BFP: return launch_col_tma<IWQ_F16, KIND_BFP>(p, single, maps, sym, t, st);
      default: return launch_col_tma<IWQ_F16, KIND_INT_GENERIC>(p, single, maps, sym, t, st);
    }
  }
  switch (kind) {
    case KIND_INT_FAST: return launch_col_tma<IWQ_BF16, KIND_INT_FAST>(p, single, maps, sym, t, st);
    case KIND_INT_WIDE: return launch_col_tma<IWQ_BF16, KIND_INT_WIDE>(p, single, maps, sym, t, st);
    case KIND_FP: return launch_col_tma<IWQ_BF16, KIND_FP>(p, single, maps, sym, t, st);
    case KIND_BFP: return launch_col_tma<IWQ_BF16, KIND_BFP>(p, single, maps, sym, t, st);
    default: return launch_col_tma<IWQ_BF16, KIND_INT_GENERIC>(p, single, maps, sym, t, st);
  }
}

// tile shape of col_tma_kernel: 8 KB boxes (g rows x 4096 / g columns; 64-byte rows at least) -> columns per tile, 0 = not applicable
static uint32_t col_tma_tile(int64_t g) {
  if (knobs().col_mode != 2 && knobs().col_mode != 3) return 0;
  if (g != 32 && g != 64 && g != 128 && g != 256) return 0;
  const int b = knobs().col_box;
  if ((b == 32 || b == 64 || b == 128 || b == 256) && (int64_t)b * g * 2 <= 48 * 1024 && (int64_t)b * g >= 2048) return (uint32_t)b;
  return g == 32 ? 128u : g == 64 ? 64u : 32u;
}

static int run_col(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, LaunchParams& p, int kind, void* workspace, size_t ws_bytes, cudaStream_t st) {
  std::vector<Job> jobs;
  uint64_t units = 0;
  int rpl = 0, stages = 0;
  const uint32_t tma_tile = (cfg->group > 0 && p.qc.approx != 2) ? col_tma_tile(cfg->group) : 0u;
  const uint32_t ctile = tma_tile ? tma_tile : ((cfg->group > 0 && p.qc.approx != 2) ? col_tile_for(cfg->group, &rpl, &stages) : 0u);
  // col_wide_kernel: 128 x 256 tiles through a CTA-wide ring (every matrix must hold at least one full tile)
  bool wide = tma_tile && knobs().col_mode == 3 && (cfg->group == 32 || cfg->group == 64 || cfg->group == 128);
  for (int i = 0; i < n && wide; ++i) wide = ms[i].rows >= kColWideRows && ms[i].cols >= kColWideCols;
  ColWideShape shape = {0, 0, 0};
  if (wide && kind == KIND_FP && knobs().fp_mode == 2 && p.qc.approx != 2) {
    // DIRECT minifloat table (one 16-bit value per bit pattern: one 2-byte load per element instead of an 8-byte load, three
    // byte permutes, a compare and a select) when a tile shape with at least three stages fits next to it
    QuantConsts q = p.qc;
    set_lut_direct(cfg->dtype, &q);
    const uint32_t direct_bytes = (uint32_t)(((size_t)q.lut_n * 2 + 1023) & ~(size_t)1023);
    // (measured: the direct table pays only with full 128 x 256 tiles -- FP4 at group 128, every format on bf16; with smaller
    //  tiles the step table and full tiles win: profiles/r02_sweep_col18.jsonl)
    if (col_wide_shape(cfg->group, kind, p.qc.symmetric != 0, direct_bytes, &shape, true)) {
      p.qc = q;
      int rc = get_fp_lut(cfg, p.qc, st, &p.lut);
      if (rc) return rc;
      p.lut_smem_bytes = direct_bytes;
    }
  }
  if (wide && shape.rows == 0) wide = col_wide_shape(cfg->group, kind, p.qc.symmetric != 0, p.lut_smem_bytes, &shape);
  const int64_t wide_rows = wide ? (int64_t)shape.rows : 0;
  const uint32_t dbl_tile = knobs().col_dbl_tile == 64 ? 64u : 256u;
  const uint32_t unit_cols = wide ? shape.cols : (ctile ? ctile : (p.qc.approx == 2 ? dbl_tile : (uint32_t)kColTile));
  for (int i = 0; i < n; ++i) {
    int rc = check_matrix(&ms[i]);
    if (rc) return rc;
    const int64_t rows = ms[i].rows, cols = ms[i].cols;
    if (ms[i].codes) return fail(IWQ_ERR_UNSUPPORTED, "code emission is implemented for quant_dim = 0 only");
    int64_t g;
    if (cfg->group > 0) {
      if (rows % cfg->group != 0) return fail(IWQ_ERR_INDIVISIBLE, "rows is not divisible by the group size (quant_dim = 1)");
      g = cfg->group;
    } else if (cfg->group == IWQ_GROUP_PER_CHANNEL) {
      g = rows;
    } else {
      return fail(IWQ_ERR_INVALID, "Invalid w_group_size");
    }
    if (cols % 8 != 0) return fail(IWQ_ERR_UNSUPPORTED, "quant_dim = 1 needs cols to be a multiple of 8 (16-byte vectors)");
    if (g >= (1ll << 31)) return fail(IWQ_ERR_INVALID, "segment too long");
    Job j = make_job(&ms[i], (uint32_t)units);
    uint64_t row_blocks = (uint64_t)(rows / g);
    if (p.qc.approx == 2) {
      // the reference groups 4 codes along the flattened transpose; they share a column only when the number of
      // row blocks is a multiple of 4 (true for every LLaMA / OPT shape at group 32..256)
      if (row_blocks % 4 != 0) return fail(IWQ_ERR_UNSUPPORTED, "double-approximate decode needs (rows / group) % 4 == 0");
      row_blocks /= 4;
    }
    if (wide) row_blocks = (uint64_t)((rows + wide_rows - 1) / wide_rows);
    units += row_blocks * (uint64_t)((cols + unit_cols - 1) / unit_cols);
    if (units >= 0xFFFFFFFFull) return fail(IWQ_ERR_INVALID, "too much work for one launch");
    jobs.push_back(j);
  }
  p.n_jobs = n;
  p.total_units = (uint32_t)units;
  p.single = jobs[0];
  p.jobs = nullptr;
  p.col_group = cfg->group > 0 ? (uint32_t)cfg->group : 0u;
  p.col_tile = p.qc.approx == 2 ? dbl_tile : ctile;
  p.col_rpl = (uint32_t)rpl;
  if (ctile && rpl == 0) {
    p.stage_bytes = (uint32_t)cfg->group * (ctile / 2);  // g rows x 16 * lpr bytes, lpr = ctile / 32
    p.n_stages = stages;
  }
  int rc = upload_jobs(jobs, workspace, ws_bytes, &p, st);
  if (rc) return rc;
  LaunchTuning tune;
  tune.stages = 0;
  tune.ctas_per_sm = knobs().col_ctas;
  if (tma_tile) {
    // ---- col_tma_kernel: tensor maps (input / output) per matrix; behind the job table in the workspace, or in the kernel
    //      parameters for a single matrix ----
    p.stage_bytes = (uint32_t)cfg->group * tma_tile * 2u;
    p.n_stages = knobs().col_stages == 1 ? 1 : 2;
    if (p.stage_bytes >= 16384 && knobs().col_stages == 0) p.n_stages = 1;  // 16 KB boxes: one stage, more resident warps
    p.l2_hint = (uint32_t)knobs().col_hint & 3u;
    std::vector<TensorMapPair> maps((size_t)n);
    if (wide) {
      // load boxes of TR rows x CC columns, store boxes of TR / 8 rows x CC columns (one per consumer warp)
      p.col_wide = 1;
      p.col_tile = shape.cols;
      p.col_rows = shape.rows;
      p.stage_bytes = shape.rows * shape.cols * 2u;
      p.n_stages = (int)shape.stages;
      for (int i = 0; i < n; ++i) {
        rc = encode_map(&maps[i].in, ms[i].w, ms[i].rows, ms[i].cols, shape.rows, shape.cols);
        if (rc) return rc;
        rc = encode_map(&maps[i].out, ms[i].out, ms[i].rows, ms[i].cols, shape.rows / kColWideWG, shape.cols);
        if (rc) return rc;
      }
    }
    for (int i = 0; i < n && !wide; ++i) {
      rc = encode_map(&maps[i].in, ms[i].w, ms[i].rows, ms[i].cols, (uint32_t)cfg->group, tma_tile);
      if (rc) return rc;
      if (ms[i].out == ms[i].w) maps[i].out = maps[i].in;
      else {
        rc = encode_map(&maps[i].out, ms[i].out, ms[i].rows, ms[i].cols, (uint32_t)cfg->group, tma_tile);
        if (rc) return rc;
      }
    }
    const TensorMapPair* dmaps = nullptr;
    if (n > 1) {
      const size_t off = ((size_t)n * sizeof(Job) + 127) & ~(size_t)127;
      if (!workspace || ws_bytes < off + (size_t)n * sizeof(TensorMapPair)) return fail(IWQ_ERR_INVALID, "workspace too small (see iwq_workspace_bytes)");
      if ((uintptr_t)workspace & 127) return fail(IWQ_ERR_ALIGN, "workspace must be 128-byte aligned for quant_dim = 1");
      cudaError_t e2 = cudaMemcpyAsync((uint8_t*)workspace + off, maps.data(), (size_t)n * sizeof(TensorMapPair), cudaMemcpyHostToDevice, st);
      if (e2 != cudaSuccess) return fail_cuda(e2, "cudaMemcpyAsync(tensor maps)");
      dmaps = (const TensorMapPair*)((uint8_t*)workspace + off);
    }
    cudaError_t e2 = dispatch_col_tma(cfg->dtype, kind, p, maps[0], dmaps, p.qc.symmetric != 0, tune, st);
    if (e2 != cudaSuccess) return fail_cuda(e2, "column-group (tensor map) kernel launch");
    return IWQ_OK;
  }
  // the light INT / BFP element stages run best with 4 CTAs (16 warps) per SM at group >= 128; the minifloat
  // stage (two table lookups per pair) wants every warp it can get
  if (tune.ctas_per_sm == 0 && ctile && rpl == 0 && kind != KIND_FP && cfg->group >= 128) tune.ctas_per_sm = 4;
  cudaError_t e;
  if (p.qc.approx == 2)
    e = cfg->dtype == IWQ_F16 ? launch_col_double<IWQ_F16>(p, st) : launch_col_double<IWQ_BF16>(p, st);
  else
    e = dispatch_col(cfg->dtype, kind, p, p.qc.symmetric != 0, tune, st);
  if (e != cudaSuccess) return fail_cuda(e, "column-group kernel launch");
  return IWQ_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer pipeline
// ------------------------------------------------------------------------------------------------
struct HostPipe {
  int device = -1;
  static constexpr int kBuf = 4;
  void* buf[kBuf] = {nullptr, nullptr, nullptr, nullptr};
  size_t cap = 0;
  cudaStream_t stream[kBuf] = {nullptr, nullptr, nullptr, nullptr};
  int32_t* status = nullptr;
  int32_t* status_host = nullptr;
};
static std::mutex g_pipe_mu;
static HostPipe g_pipes[16];

static int pipe_prepare(HostPipe* hp, int device, size_t chunk) {
  cudaError_t e;
  if (hp->device != device) {
    hp->device = device;
    for (int i = 0; i < HostPipe::kBuf; ++i) {
      e = cudaStreamCreateWithFlags(&hp->stream[i], cudaStreamNonBlocking);
      if (e != cudaSuccess) return fail_cuda(e, "cudaStreamCreate");
    }
    e = cudaMalloc((void**)&hp->status, sizeof(int32_t));
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(status)");
    e = cudaMallocHost((void**)&hp->status_host, sizeof(int32_t));
    if (e != cudaSuccess) return fail_cuda(e, "cudaMallocHost(status)");
  }
  if (hp->cap < chunk) {
    for (int i = 0; i < HostPipe::kBuf; ++i) {
      if (hp->buf[i]) cudaFree(hp->buf[i]);
      hp->buf[i] = nullptr;
    }
    hp->cap = 0;
    for (int i = 0; i < HostPipe::kBuf; ++i) {
      e = cudaMalloc(&hp->buf[i], chunk);
      if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(staging)");
    }
    hp->cap = chunk;
  }
  return IWQ_OK;
}

}  // namespace iwq

using namespace iwq;

extern "C" {

int iwq_abi_version(void) { return IWQ_ABI_VERSION; }

// Which kernel and tile shape a [rows, cols] matrix would get under `cfg` (no CUDA call: testable on a GPU-less box).
// out[0] path: 0 seg4 / seg / direct (see out[1]), 1 row_tma_kernel, 2 col_async_kernel, 3 col_reg_kernel, 4 col_kernel
//              (two sweeps), 5 col_double_kernel, 6 col_tma_kernel (2-D tensor-map boxes)
// out[1] kernel variant of the segment path (0 seg4_tma, 1 seg_direct, 2 seg_tma), out[2] lanes per segment,
// out[3] ring stages, out[4] consumer warps (row path), out[5] bytes per stage (row / column paths),
// out[6] row kernel sends its results through bulk stores, out[7] columns per tile (column paths)
int iwq_debug_plan(const iwq_config_t* cfg, int64_t rows, int64_t cols, int32_t* out) {
  if (!cfg || !out) return fail(IWQ_ERR_INVALID, "null argument");
  memset(out, 0, 8 * sizeof(int32_t));
  QuantConsts qc;
  int kind = 0;
  int rc = build_consts(cfg, &qc, &kind);
  if (rc) return rc;
  const uint32_t lut_bytes = kind == KIND_FP ? (uint32_t)(((size_t)qc.lut_n * sizeof(FpStep) + 1023) & ~(size_t)1023) : 0u;
  if (cfg->quant_dim == 1 && cfg->group != IWQ_GROUP_PER_TENSOR) {
    if (qc.approx == 2) {
      out[0] = 5;
      out[7] = knobs().col_dbl_tile == 64 ? 64 : 256;
      return IWQ_OK;
    }
    int rpl = 0, stages = 0;
    if (cfg->group > 0 && col_tma_tile(cfg->group) && knobs().col_mode == 3 && (cfg->group == 32 || cfg->group == 64 || cfg->group == 128) && rows >= kColWideRows &&
        cols >= kColWideCols) {
      ColWideShape shape = {0, 0, 0};
      uint32_t lb = lut_bytes;
      if (kind == KIND_FP && knobs().fp_mode == 2 && qc.approx != 2) {
        QuantConsts q = qc;
        set_lut_direct(cfg->dtype, &q);
        const uint32_t direct_bytes = (uint32_t)(((size_t)q.lut_n * 2 + 1023) & ~(size_t)1023);
        if (col_wide_shape(cfg->group, kind, qc.symmetric != 0, direct_bytes, &shape, true)) lb = direct_bytes;
      }
      if (shape.rows != 0 || col_wide_shape(cfg->group, kind, qc.symmetric != 0, lb, &shape)) {
        out[0] = 7;  // col_wide_kernel
        out[3] = (int32_t)shape.stages;
        out[4] = (int32_t)shape.rows;
        out[5] = (int32_t)(shape.rows * shape.cols * 2u);
        out[6] = lb != lut_bytes ? 1 : 0;  // direct minifloat table
        out[7] = (int32_t)shape.cols;
        return IWQ_OK;
      }
    }
    if (cfg->group > 0 && col_tma_tile(cfg->group)) {
      out[0] = 6;  // col_tma_kernel
      out[3] = (knobs().col_stages == 1 || (cfg->group == 256 && knobs().col_stages == 0)) ? 1 : 2;
      out[5] = (int32_t)((uint32_t)cfg->group * col_tma_tile(cfg->group) * 2u);
      out[7] = (int32_t)col_tma_tile(cfg->group);
      return IWQ_OK;
    }
    const uint32_t ctile = cfg->group > 0 ? col_tile_for(cfg->group, &rpl, &stages) : 0u;
    out[0] = ctile ? (rpl ? 3 : 2) : 4;
    out[3] = stages;
    out[5] = (ctile && !rpl) ? (int32_t)((uint32_t)cfg->group * (ctile / 2)) : 0;
    out[7] = (int32_t)(ctile ? ctile : (uint32_t)kColTile);
    return IWQ_OK;
  }
  Plan pl;
  rc = make_plan(cfg, cols, lut_bytes, &pl);
  if (rc) return rc;
  const bool per_tensor = cfg->group == IWQ_GROUP_PER_TENSOR;
  if (pl.path == PATH_SEG) {
    int variant = effective_variant(cfg);
    if (variant < 0 || variant > 2) variant = 0;
    if (variant == 0 && !per_tensor && pl.seg_elems < 32) variant = 2;
    const int st = fit_seg_stages(knobs().seg_stages, lut_bytes);
    out[0] = 0;
    out[1] = variant;
    out[2] = variant == 0 ? (per_tensor ? 1 : (int32_t)(pl.seg_elems / 32)) : pl.lps;
    out[3] = st;
    return IWQ_OK;
  }
  out[0] = 1;
  out[3] = pl.stages;
  out[4] = pl.consumers / 32;
  out[5] = (int32_t)pl.stage_bytes;
  out[6] = (pl.stages >= 3 && knobs().row_tma_out != 0 && kind != KIND_FP && (knobs().row_tma_out > 1 || pl.seg_elems * 2ull >= 20 * 1024)) ? 1 : 0;
  (void)rows;
  return IWQ_OK;
}
const char* iwq_last_error(void) { return g_err.c_str(); }

// Tuning knobs are read from the environment once (IWQ_*); this sets one at run time so that a sweep can run inside one
// process (tools/sweep_r2.py).  `name` is the environment variable's name without the IWQ_ prefix, lower case.
int iwq_debug_set_knob(const char* name, int value) {
  if (!name) return fail(IWQ_ERR_INVALID, "null knob name");
  Knobs& k = knobs_mut();
  const std::string n(name);
#define IWQ_KNOB(field)   \
  if (n == #field) {      \
    k.field = value;      \
    return IWQ_OK;        \
  }
  IWQ_KNOB(seg_stages) IWQ_KNOB(seg_ctas) IWQ_KNOB(row_ctas) IWQ_KNOB(row_smem_kb) IWQ_KNOB(variant) IWQ_KNOB(row_warps) IWQ_KNOB(col_tile)
  IWQ_KNOB(col_ctas) IWQ_KNOB(fused_batch_mb) IWQ_KNOB(fused_lookahead) IWQ_KNOB(row_tma_out) IWQ_KNOB(wide_seg4) IWQ_KNOB(col_mode)
  IWQ_KNOB(col_stages) IWQ_KNOB(pt_slab_mb) IWQ_KNOB(pt_reverse) IWQ_KNOB(pt_hint) IWQ_KNOB(seg_hint) IWQ_KNOB(fp_mode) IWQ_KNOB(fp_rep) IWQ_KNOB(fp_rep_kb)
  IWQ_KNOB(fp_stages) IWQ_KNOB(host_copy_only) IWQ_KNOB(seg_map) IWQ_KNOB(mm_tma) IWQ_KNOB(mm_stages) IWQ_KNOB(mm_hint) IWQ_KNOB(col_box) IWQ_KNOB(col_hint) IWQ_KNOB(seg_vpl) IWQ_KNOB(ws_cache) IWQ_KNOB(col_dbl_tile)
#undef IWQ_KNOB
  return fail(IWQ_ERR_INVALID, "unknown knob: " + n);
}
int iwq_workspace_cache(int enable) {
  std::lock_guard<std::mutex> lk(g_ws_mu);
  g_ws_cache.clear();  // every call forgets what was uploaded (call it again after freeing / reallocating a workspace)
  knobs_mut().ws_cache = enable ? 1 : 0;
  return IWQ_OK;
}
// job table, then (128-byte aligned) one pair of tensor maps per matrix for the quant_dim = 1 kernels
size_t iwq_workspace_bytes(int n) {
  const size_t m = (size_t)(n < 1 ? 1 : n);
  return ((m * sizeof(Job) + 127) & ~(size_t)127) + m * sizeof(TensorMapPair);
}

int iwq_fake_quant(const iwq_config_t* cfg, const iwq_matrix_t* m, float* range_ws, int32_t* status, void* stream) {
  return run_batched(cfg, m, 1, nullptr, 0, nullptr, range_ws, status, (cudaStream_t)stream, false);
}

int iwq_fake_quant_batched(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, void* workspace, size_t workspace_bytes, int32_t* status,
                           void* stream) {
  if (cfg && cfg->group == IWQ_GROUP_PER_TENSOR)
    return fail(IWQ_ERR_INVALID, "per-tensor granularity: use iwq_minmax + iwq_fake_quant_with_range (or iwq_fake_quant per matrix)");
  return run_batched(cfg, ms, n, workspace, workspace_bytes, nullptr, nullptr, status, (cudaStream_t)stream, false);
}

int iwq_fake_quant_per_tensor_batched(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, float* range, void* workspace, size_t workspace_bytes,
                                      int32_t* status, void* stream) {
  if (!cfg || cfg->group != IWQ_GROUP_PER_TENSOR) return fail(IWQ_ERR_INVALID, "iwq_fake_quant_per_tensor_batched is for per-tensor granularity");
  if (!range) return fail(IWQ_ERR_INVALID, "null range buffer");
  return run_batched(cfg, ms, n, workspace, workspace_bytes, nullptr, range, status, (cudaStream_t)stream, false);
}

int iwq_minmax(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, float* range, void* workspace, size_t workspace_bytes, void* stream) {
  if (!cfg || cfg->group != IWQ_GROUP_PER_TENSOR) return fail(IWQ_ERR_INVALID, "iwq_minmax is for per-tensor granularity");
  if (!range) return fail(IWQ_ERR_INVALID, "null range buffer");
  return run_batched(cfg, ms, n, workspace, workspace_bytes, nullptr, range, nullptr, (cudaStream_t)stream, true);
}

int iwq_fake_quant_with_range(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, const float* range, void* workspace,
                              size_t workspace_bytes, int32_t* status, void* stream) {
  if (!cfg || cfg->group != IWQ_GROUP_PER_TENSOR) return fail(IWQ_ERR_INVALID, "iwq_fake_quant_with_range is for per-tensor granularity");
  if (!range) return fail(IWQ_ERR_INVALID, "null range buffer");
  return run_batched(cfg, ms, n, workspace, workspace_bytes, range, nullptr, status, (cudaStream_t)stream, false);
}

static long long g_fused_trace_off = -1;
static int g_fused_trace_nb = 0;
// diagnostics (IWQ_FUSED_TRACE=1): offset of the [n_batches][4] uint64 globaltimer stamps inside the workspace of the
// last fused call, or -1; *n_batches receives the batch count
long long iwq_debug_fused_trace(int* n_batches) {
  if (n_batches) *n_batches = g_fused_trace_nb;
  return g_fused_trace_off;
}

size_t iwq_fused_workspace_bytes(int n) {
  const size_t m = (size_t)(n < 1 ? 1 : n);
  // jobs | batch_tile0 | batch_job0 | arrive | local_range | local slots [2][1][2n] | local flags | 2 x 16 peer pointers
  return m * sizeof(Job) + 2 * (m + 1) * 4 + m * 4 + 2 * m * 4 + 4 * m * 4 + m * 4 + 2 * 16 * 8 + 256 + m * 32 + 2 * m * 4;
}

int iwq_fake_quant_per_tensor_fused(const iwq_config_t* cfg, const iwq_matrix_t* ms, int n, void* workspace, size_t workspace_bytes,
                                    const iwq_peers_t* peers, int32_t* status, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!cfg || !ms || n < 1) return fail(IWQ_ERR_INVALID, "null config/matrices or n < 1");
  if (cfg->group != IWQ_GROUP_PER_TENSOR || cfg->quant_dim != 0) return fail(IWQ_ERR_INVALID, "iwq_fake_quant_per_tensor_fused is for per-tensor granularity");
  if (!workspace || workspace_bytes < iwq_fused_workspace_bytes(n)) return fail(IWQ_ERR_INVALID, "workspace too small (see iwq_fused_workspace_bytes)");
  if ((uintptr_t)workspace & 15) return fail(IWQ_ERR_ALIGN, "workspace must be 16-byte aligned");
  const int world = peers ? peers->world : 1, rank = peers ? peers->rank : 0;
  if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(IWQ_ERR_INVALID, "bad world / rank");
  if (peers && (!peers->slot_ptrs || !peers->flag_ptrs || peers->epoch == 0)) return fail(IWQ_ERR_INVALID, "peers: null pointer arrays or epoch 0");
  LaunchParams p;
  memset(&p, 0, sizeof(p));
  int kind = 0;
  int rc = build_consts(cfg, &p.qc, &kind);
  if (rc) return rc;
  if (kind == KIND_BFP) return fail(IWQ_ERR_INVALID, "BFP needs a positive group size");
  if (kind == KIND_FP) {
    rc = get_fp_lut(cfg, p.qc, st, &p.lut);
    if (rc) return rc;
    p.lut_smem_bytes = (uint32_t)(((size_t)p.qc.lut_n * sizeof(FpStep) + 1023) & ~(size_t)1023);
  }
  Plan pl;
  memset(&pl, 0, sizeof(pl));
  pl.path = PATH_SEG;
  pl.lps = 1;
  pl.vpl = 1;
  std::vector<Job> jobs;
  rc = fill_params(cfg, ms, n, pl, &p, &jobs);
  if (rc) return rc;
  p.n_stages = 3;  // the fused kernel parks one stage behind its bulk store across item changes; its own ring stays at 3 tiles
  for (int i = 0; i < n; ++i)
    if (ms[i].codes) return fail(IWQ_ERR_UNSUPPORTED, "the fused per-tensor entry point does not emit codes");
  p.status = status;
  // batches of whole matrices
  const uint64_t batch_bytes = (uint64_t)(knobs().fused_batch_mb > 0 ? knobs().fused_batch_mb : 16) << 20;
  std::vector<uint32_t> b_tile0, b_job0;
  uint64_t acc = 0;
  int in_batch = 0;
  for (int i = 0; i < n; ++i) {
    if (acc == 0) {
      b_tile0.push_back(jobs[i].first_unit);
      b_job0.push_back((uint32_t)i);
      in_batch = 0;
    }
    acc += jobs[i].bytes;
    if (acc >= batch_bytes || ++in_batch >= kFusedMaxBatchJobs) acc = 0;
  }
  const int nb = (int)b_tile0.size();
  if (nb > 2048) return fail(IWQ_ERR_UNSUPPORTED, "too many batches for one fused launch (raise IWQ_FUSED_BATCH_MB or split the call)");
  b_tile0.push_back(p.total_units);
  b_job0.push_back((uint32_t)n);
  // one host blob -> one H2D copy: tables, zeroed counters / local flags, -inf ranges, peer pointer arrays
  const size_t m = (size_t)n;
  size_t off = 0;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) / align * align; size_t o = off; off += bytes; return o; };
  const size_t o_jobs = take(m * sizeof(Job), 16), o_bt = take((m + 1) * 4, 4), o_bj = take((m + 1) * 4, 4), o_arr = take(m * 4, 4);
  const size_t o_rng = take(2 * m * 4, 4), o_slots = take(4 * m * 4, 16), o_flags = take(m * 4, 4), o_sp = take(16 * 8, 8), o_fp = take(16 * 8, 8);
  const size_t o_ctr = take(2 * m * 4, 4);
  const bool want_trace = env_int("IWQ_FUSED_TRACE", 0) != 0;
  const size_t o_trace = want_trace ? take((size_t)nb * 4 * 8, 8) : 0;
  g_fused_trace_off = want_trace ? (long long)o_trace : -1;
  g_fused_trace_nb = nb;
  if (off > workspace_bytes) return fail(IWQ_ERR_INVALID, "workspace too small (see iwq_fused_workspace_bytes)");
  std::vector<uint8_t> blob(off, 0);
  memcpy(blob.data() + o_jobs, jobs.data(), m * sizeof(Job));
  memcpy(blob.data() + o_bt, b_tile0.data(), b_tile0.size() * 4);
  memcpy(blob.data() + o_bj, b_job0.data(), b_job0.size() * 4);
  for (size_t i = 0; i < 2 * m; ++i) reinterpret_cast<float*>(blob.data() + o_rng)[i] = -INFINITY;
  uint8_t* ws = (uint8_t*)workspace;
  FusedParams f;
  memset(&f, 0, sizeof(f));
  void** sp = reinterpret_cast<void**>(blob.data() + o_sp);
  void** fp = reinterpret_cast<void**>(blob.data() + o_fp);
  if (peers) {
    for (int r = 0; r < world; ++r) {
      if (!peers->slot_ptrs[r] || !peers->flag_ptrs[r]) return fail(IWQ_ERR_INVALID, "peers: null buffer pointer");
      sp[r] = peers->slot_ptrs[r];
      fp[r] = peers->flag_ptrs[r];
    }
    f.epoch = peers->epoch;
  } else {
    sp[0] = ws + o_slots;  // zeroed by the blob copy; epoch 1 of a fresh flag array
    fp[0] = ws + o_flags;
    f.epoch = 1;
  }
  cudaError_t e = cudaMemcpyAsync(workspace, blob.data(), off, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpyAsync(fused tables)");
  p.jobs = (const Job*)(ws + o_jobs);
  f.local_range = (float*)(ws + o_rng);
  f.arrive = (uint32_t*)(ws + o_arr);
  f.batch_tile0 = (const uint32_t*)(ws + o_bt);
  f.batch_job0 = (const uint32_t*)(ws + o_bj);
  f.peer_slots = (float* const*)(ws + o_sp);
  f.peer_flags = (uint32_t* const*)(ws + o_fp);
  f.n_batches = nb;
  f.world = world;
  f.rank = rank;
  f.spin_limit = (peers && peers->spin_limit) ? peers->spin_limit : 2000000u;
  f.tile_ctr = (uint32_t*)(ws + o_ctr);
  f.trace = want_trace ? (unsigned long long*)(ws + o_trace) : nullptr;
  f.poll_ns = (uint32_t)env_int("IWQ_FUSED_POLL_NS", 200);
  f.debug = (uint32_t)env_int("IWQ_FUSED_DEBUG", 0);
  f.lookahead = knobs().fused_lookahead > 0 ? (knobs().fused_lookahead < kFusedRing ? knobs().fused_lookahead : kFusedRing - 1) : 2;
  LaunchTuning tune;
  tune.stages = p.n_stages;
  tune.ctas_per_sm = knobs().seg_ctas;
  int grid = 0;
  e = dispatch_fused(cfg->dtype, kind, p, f, p.qc.symmetric != 0, tune, st, &grid);
  if (e != cudaSuccess) return fail_cuda(e, "fused per-tensor kernel launch");
  return IWQ_OK;
}

int iwq_fp_roundtrip(int dtype, int flavour, int exp_bits, int mant_bits, const void* x, void* y, void* codes, int64_t n, void* stream) {
  iwq_config_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.dtype = dtype;
  cfg.scheme = flavour;
  cfg.exp_bits = exp_bits;
  cfg.mant_bits = mant_bits;
  cfg.symmetric = 1;
  cfg.group = 128;
  QuantConsts qc;
  int kind = 0;
  int rc = build_consts(&cfg, &qc, &kind);
  if (rc) return rc;
  if (kind != KIND_FP) return fail(IWQ_ERR_INVALID, "flavour must be IWQ_SCHEME_FP_CODEC or IWQ_SCHEME_FP_RNE");
  if (n <= 0 || (n & 1)) return fail(IWQ_ERR_INVALID, "n must be positive and even");
  if (((uintptr_t)x | (uintptr_t)y) & 3) return fail(IWQ_ERR_ALIGN, "x and y must be 4-byte aligned");
  const int64_t pairs = n / 2;
  int grid = (int)((pairs + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IWQ_F16)
    fp_roundtrip_kernel<IWQ_F16><<<grid, 256, 0, st>>>((const uint32_t*)x, (uint32_t*)y, (uint16_t*)codes, pairs, qc);
  else
    fp_roundtrip_kernel<IWQ_BF16><<<grid, 256, 0, st>>>((const uint32_t*)x, (uint32_t*)y, (uint16_t*)codes, pairs, qc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "fp_roundtrip kernel launch");
  return IWQ_OK;
}

int iwq_fp_scale(int exp_bits, int mant_bits, const void* x, const void* scale, void* y, int64_t n, void* stream) {
  iwq_config_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.dtype = IWQ_F16;
  cfg.scheme = IWQ_SCHEME_FP_RNE;
  cfg.exp_bits = exp_bits;
  cfg.mant_bits = mant_bits;
  cfg.symmetric = 1;
  cfg.group = 128;
  QuantConsts qc;
  int kind = 0;
  int rc = build_consts(&cfg, &qc, &kind);
  if (rc) return rc;
  if (!x || !scale || !y) return fail(IWQ_ERR_INVALID, "null pointer");
  if (n <= 0 || (n & 1)) return fail(IWQ_ERR_INVALID, "n must be positive and even");
  if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)scale) & 3) return fail(IWQ_ERR_ALIGN, "x, scale and y must be 4-byte aligned");
  const int64_t pairs = n / 2;
  int grid = (int)((pairs + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  fp_scale_kernel<IWQ_F16><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)x, (const uint32_t*)scale, (uint32_t*)y, pairs, qc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "fp_scale kernel launch");
  return IWQ_OK;
}

int iwq_selftest_division(int dtype, unsigned long long* mismatches, void* stream) {
  if (!mismatches) return fail(IWQ_ERR_INVALID, "null counter");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IWQ_F16)
    division_selftest_kernel<IWQ_F16><<<148 * 4, 256, 0, st>>>(mismatches);
  else if (dtype == IWQ_BF16)
    division_selftest_kernel<IWQ_BF16><<<148 * 4, 256, 0, st>>>(mismatches);
  else
    return fail(IWQ_ERR_INVALID, "dtype must be IWQ_F16 or IWQ_BF16");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail_cuda(e, "division selftest launch");
  return IWQ_OK;
}

// One chunk pipeline over ALL matrices of the call: chunk i goes through staging buffer / stream i % kBuf
// (H2D -> kernel -> D2H in stream order), so the H2D of later chunks overlaps the D2H of earlier ones on the two
// copy engines and the ring is drained only once, at the end of the call -- not once per matrix.
namespace {
struct DeviceGuard {  // restores the caller's current device on every exit path
  int prev = -1;
  bool armed = false;
  cudaError_t enter(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return e;
    armed = true;
    return prev == device ? cudaSuccess : cudaSetDevice(device);
  }
  ~DeviceGuard() {
    if (armed) cudaSetDevice(prev);
  }
};
}  // namespace

static int host_pipeline(const iwq_config_t* cfg, const iwq_host_matrix_t* ms, int n, int device, size_t chunk_bytes, int32_t* host_status) {
  if (!cfg || !ms || n < 1) return fail(IWQ_ERR_INVALID, "null argument");
  if (cfg->group == IWQ_GROUP_PER_TENSOR) return fail(IWQ_ERR_UNSUPPORTED, "host pipeline: per-tensor granularity needs the whole matrix on the device");
  if (cfg->quant_dim != 0) return fail(IWQ_ERR_UNSUPPORTED, "host pipeline: quant_dim = 1 needs whole columns on the device");
  if (device < 0 || device >= 16) return fail(IWQ_ERR_INVALID, "device ordinal out of range");
  if (chunk_bytes == 0) chunk_bytes = 32ull << 20;
  auto rows_per_chunk = [&](const iwq_host_matrix_t& m) {
    const size_t row_bytes = (size_t)m.cols * 2;
    int64_t r = (int64_t)(chunk_bytes / row_bytes);
    if (r < 1) r = 1;
    if (row_bytes % 16 != 0) r = (r + 7) / 8 * 8;  // keep chunk starts 16-byte aligned
    return r > m.rows ? m.rows : r;
  };
  auto segs_per_row = [&](const iwq_host_matrix_t& m) -> int64_t { return cfg->group > 0 ? m.cols / cfg->group : 1; };
  size_t max_chunk = 0, max_meta = 0;
  bool want_meta = false;
  for (int i = 0; i < n; ++i) {
    if (!ms[i].host_w || !ms[i].host_out) return fail(IWQ_ERR_INVALID, "null host pointer");
    if (ms[i].rows <= 0 || ms[i].cols <= 0) return fail(IWQ_ERR_INVALID, "rows and cols must be positive");
    if (cfg->group > 0 && ms[i].cols % cfg->group != 0) return fail(IWQ_ERR_INDIVISIBLE, "cols is not divisible by the group size");
    const int64_t cr = rows_per_chunk(ms[i]);
    const size_t bytes = (size_t)cr * (size_t)ms[i].cols * 2;
    if (bytes > max_chunk) max_chunk = bytes;
    if (ms[i].host_scales || ms[i].host_zeros) {
      want_meta = true;
      const size_t mb = ((size_t)(cr * segs_per_row(ms[i])) * 2 + 255) & ~(size_t)255;
      if (mb > max_meta) max_meta = mb;
    }
  }
  DeviceGuard guard;
  cudaError_t e = guard.enter(device);
  if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
  std::lock_guard<std::mutex> lk(g_pipe_mu);
  HostPipe* hp = &g_pipes[device];
  // staging: [weights chunk | scales of the chunk | zero points of the chunk] per ring slot
  const size_t slot_bytes = ((max_chunk + 255) & ~(size_t)255) + (want_meta ? 2 * max_meta : 0);
  int rc = pipe_prepare(hp, device, slot_bytes);
  if (rc) return rc;
  const size_t meta_off = (max_chunk + 255) & ~(size_t)255;
  e = cudaMemsetAsync(hp->status, 0, sizeof(int32_t), hp->stream[0]);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemsetAsync");
  e = cudaStreamSynchronize(hp->stream[0]);
  if (e != cudaSuccess) return fail_cuda(e, "cudaStreamSynchronize");
  // on any error below: let the queued copies / kernels finish before returning -- D2H copies still in flight would
  // otherwise keep writing into the caller's buffers after the error return
  auto drain = [&]() {
    for (int b = 0; b < HostPipe::kBuf; ++b) cudaStreamSynchronize(hp->stream[b]);
  };
  const bool copy_only = knobs().host_copy_only != 0;  // diagnostics: the same copies without the kernels
  int ci = 0;
  for (int i = 0; i < n; ++i) {
    const int64_t rows = ms[i].rows, cols = ms[i].cols;
    const size_t row_bytes = (size_t)cols * 2;
    const int64_t chunk_rows = rows_per_chunk(ms[i]), spr = segs_per_row(ms[i]);
    for (int64_t r0 = 0; r0 < rows; r0 += chunk_rows, ++ci) {
      const int64_t nr = rows - r0 < chunk_rows ? rows - r0 : chunk_rows;
      const size_t bytes = (size_t)nr * row_bytes;
      const int b = ci % HostPipe::kBuf;
      cudaStream_t st = hp->stream[b];
      uint8_t* dbuf = (uint8_t*)hp->buf[b];
      e = cudaMemcpyAsync(dbuf, (const uint8_t*)ms[i].host_w + (size_t)r0 * row_bytes, bytes, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) { drain(); return fail_cuda(e, "cudaMemcpyAsync(H2D)"); }
      iwq_matrix_t m;
      memset(&m, 0, sizeof(m));
      m.w = dbuf;
      m.out = dbuf;
      m.rows = nr;
      m.cols = cols;
      if (ms[i].host_scales) m.scales = dbuf + meta_off;
      if (ms[i].host_zeros) m.zeros = dbuf + meta_off + max_meta;
      if (!copy_only) {
        rc = run_batched(cfg, &m, 1, nullptr, 0, nullptr, nullptr, hp->status, st, false);
        if (rc) { drain(); return rc; }
      }
      e = cudaMemcpyAsync((uint8_t*)ms[i].host_out + (size_t)r0 * row_bytes, dbuf, bytes, cudaMemcpyDeviceToHost, st);
      const size_t meta_bytes = (size_t)(nr * spr) * 2, meta_at = (size_t)(r0 * spr) * 2;
      if (e == cudaSuccess && ms[i].host_scales)
        e = cudaMemcpyAsync((uint8_t*)ms[i].host_scales + meta_at, dbuf + meta_off, meta_bytes, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess && ms[i].host_zeros)
        e = cudaMemcpyAsync((uint8_t*)ms[i].host_zeros + meta_at, dbuf + meta_off + max_meta, meta_bytes, cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) { drain(); return fail_cuda(e, "cudaMemcpyAsync(D2H)"); }
    }
  }
  for (int b = 0; b < HostPipe::kBuf; ++b) {
    e = cudaStreamSynchronize(hp->stream[b]);
    if (e != cudaSuccess) { drain(); return fail_cuda(e, "cudaStreamSynchronize"); }
  }
  e = cudaMemcpy(hp->status_host, hp->status, sizeof(int32_t), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpy(status)");
  if (host_status) *host_status = *hp->status_host;
  return IWQ_OK;
}

int iwq_fake_quant_host(const iwq_config_t* cfg, const void* host_w, void* host_out, int64_t rows, int64_t cols, int device,
                        size_t chunk_bytes, int32_t* host_status) {
  iwq_host_matrix_t m;
  memset(&m, 0, sizeof(m));
  m.host_w = host_w;
  m.host_out = host_out;
  m.rows = rows;
  m.cols = cols;
  return host_pipeline(cfg, &m, 1, device, chunk_bytes, host_status);
}

int iwq_fake_quant_host_batched(const iwq_config_t* cfg, const iwq_host_matrix_t* ms, int n, int device, size_t chunk_bytes,
                                int32_t* host_status) {
  return host_pipeline(cfg, ms, n, device, chunk_bytes, host_status);
}

}  // extern "C"
