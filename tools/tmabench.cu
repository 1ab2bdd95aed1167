// tmabench.cu -- ceiling probe for the seg4 streaming pattern: global --bulk load--> shared stage --bulk store--> global with NO
// arithmetic and no consumer warps; one lane per CTA drives a ring of `stages` tiles of `tile` bytes.  Tells how fast an in-place
// (or out-of-place) read+write stream can go on this part as a function of tile size, ring depth, CTAs per SM, L2 policies and the
// tile -> CTA mapping, i.e. what roofline.frac the real kernel can be asked to reach.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmabench tools/tmabench.cu
//   run:   tools/tmabench [GB]          (prints one JSON line per variant)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_normal() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void tma_load(void* s, const void* g, uint32_t bytes, uint64_t* bar, uint64_t pol, bool hint) {
  if (hint) asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(s)), "l"(g), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
  else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s)), "l"(g), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store(void* g, const void* s, uint32_t bytes, uint64_t pol, bool hint) {
  if (hint) asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes), "l"(pol) : "memory");
  else asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}

struct P {
  uint8_t* src;
  uint8_t* dst;
  uint64_t n_tiles;
  uint32_t tile, stages;
  int mode;      // 0 read+write, 1 read only, 2 write only
  int map;       // 0: tile = cta + k * grid (sliding window), 1: contiguous chunk per CTA, 2: sliding window in blocks of 4 tiles per CTA
  int lhint, shint;  // 0 none, 1 evict_first, 2 evict_last, 3 evict_normal
  int split;     // pieces per tile for the stores (1 = one store per tile, 8 = eight stores of tile/8 like the warps of seg4)
};

// blockDim.x = 32 * W: every warp is an independent "virtual CTA" with its own ring (W = 1: the CTA-wide ring of seg4_tma_kernel,
// driven by one lane; W = 8 and tile = 2 KB: warp-private rings, every consumer warp fetching its own slice)
__global__ void __launch_bounds__(1024) ring(const __grid_constant__ P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t W = blockDim.x >> 5, warp = threadIdx.x >> 5;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem) + warp * 16;
  uint8_t* data = smem + 1024 + (size_t)warp * p.stages * p.tile;
  if ((threadIdx.x & 31) != 0) return;
  const uint32_t vid = blockIdx.x * W + warp, nv = gridDim.x * W;
  for (uint32_t i = 0; i < p.stages; ++i) mbar_init(&full[i], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint64_t lp = p.lhint == 1 ? pol_first() : (p.lhint == 2 ? pol_last() : pol_normal());
  const uint64_t sp = p.shint == 1 ? pol_first() : (p.shint == 2 ? pol_last() : pol_normal());
  uint64_t per = (p.n_tiles + nv - 1) / nv;
  uint64_t n_mine;
  if (p.map == 1) {
    const uint64_t b = per * vid;
    n_mine = b >= p.n_tiles ? 0 : (p.n_tiles - b < per ? p.n_tiles - b : per);
  } else {
    n_mine = p.n_tiles > vid ? (p.n_tiles - vid + nv - 1) / nv : 0;
    if (p.map == 2) n_mine = (p.n_tiles / 4 > vid ? (p.n_tiles / 4 - vid + nv - 1) / nv : 0) * 4;
  }
  auto tile_of = [&](uint64_t k) -> uint64_t {
    if (p.map == 1) return per * vid + k;
    if (p.map == 2) return ((k >> 2) * nv + vid) * 4 + (k & 3);
    return vid + k * nv;
  };
  // software pipeline: all S stages are loaded up front; tile k's store is issued as soon as its load has landed, and once the
  // PREVIOUS tile's store has finished reading shared memory (wait_group.read 1) that stage is reloaded with tile k-1+S
  const uint32_t S = p.stages;
  auto issue_load = [&](uint64_t k) {
    const uint32_t st = (uint32_t)(k % S);
    mbar_expect(&full[st], p.tile);
    tma_load(data + (size_t)st * p.tile, p.src + tile_of(k) * p.tile, p.tile, &full[st], lp, p.lhint != 0);
  };
  if (p.mode != 2)
    for (uint64_t k = 0; k < n_mine && k < S; ++k) issue_load(k);
  for (uint64_t k = 0; k < n_mine; ++k) {
    const uint32_t st = (uint32_t)(k % S);
    if (p.mode != 2) mbar_wait(&full[st], (uint32_t)((k / S) & 1));
    if (p.mode == 1) {  // read only: the stage is free as soon as the data has landed
      if (k + S < n_mine) issue_load(k + S);
      continue;
    }
    const uint32_t piece = p.tile / p.split;
    for (int q = 0; q < p.split; ++q) tma_store(p.dst + tile_of(k) * p.tile + (size_t)q * piece, data + (size_t)st * p.tile + (size_t)q * piece, piece, sp, p.shint != 0);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    if (p.mode == 0 && k >= 1 && k - 1 + S < n_mine) issue_load(k - 1 + S);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// plain vectorised in-place (or out-of-place) copy: every thread moves U 16-byte vectors per iteration (all loads, then all stores);
// a CTA owns a contiguous block of threads * U * 16 bytes per iteration, blocks are handed out grid-strided.
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int U>
__global__ void plain(const uint8_t* src, uint8_t* dst, uint64_t bytes) {
  const uint64_t blk = (uint64_t)blockDim.x * U * 16;
  const uint64_t n_blk = bytes / blk;
  for (uint64_t b = blockIdx.x; b < n_blk; b += gridDim.x) {
    const uint8_t* s = src + b * blk + threadIdx.x * 16;
    uint8_t* d = dst + b * blk + threadIdx.x * 16;
    uint4 v[U];
#pragma unroll
    for (int i = 0; i < U; ++i) v[i] = ldg128(s + (size_t)i * blockDim.x * 16);
#pragma unroll
    for (int i = 0; i < U; ++i) stg128(d + (size_t)i * blockDim.x * 16, v[i]);
  }
}
template <int U>
static float run_plain(const uint8_t* src, uint8_t* dst, uint64_t bytes, int threads, int grid, cudaEvent_t e0, cudaEvent_t e1) {
  float tot = 0;
  for (int it = 0; it < 6; ++it) {
    CK(cudaEventRecord(e0));
    plain<U><<<grid, threads>>>(src, dst, bytes);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (it) tot += ms;
  }
  return tot / 5;
}

int main(int argc, char** argv) {
  const double gb = argc > 1 ? atof(argv[1]) : 8.0;
  const size_t bytes = ((size_t)(gb * 1e9) >> 20) << 20;
  uint8_t *a, *b;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 1, bytes));
  CK(cudaMemset(b, 2, bytes));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  struct V { uint32_t tile, stages; int ctas, mode, map, lhint, shint, split, inplace, warps; };
  std::vector<V> vs;
  if (argc > 2) {
    // variants from the command line: "tile,stages,ctas,mode,map,lhint,shint,split,inplace,warps;..."
    const char* q = argv[2];
    while (*q) {
      V v;
      int n = 0;
      if (sscanf(q, "%u,%u,%d,%d,%d,%d,%d,%d,%d,%d%n", &v.tile, &v.stages, &v.ctas, &v.mode, &v.map, &v.lhint, &v.shint, &v.split, &v.inplace, &v.warps, &n) != 10) break;
      vs.push_back(v);
      q += n;
      if (*q == ';') ++q;
    }
  } else {
    for (int rep = 0; rep < 2; ++rep) vs.push_back({16384, 3, 3, 0, 0, 0, 0, 8, 1, 1});
  }
  for (const V& v : vs) {
    P p;
    p.src = a;
    p.dst = v.inplace ? a : b;
    p.tile = v.tile;
    p.stages = v.stages;
    p.n_tiles = bytes / v.tile;
    p.mode = v.mode;
    p.map = v.map;
    p.lhint = v.lhint;
    p.shint = v.shint;
    p.split = v.split;
    const size_t sm = 1024 + (size_t)v.tile * v.stages * v.warps;
    const int grid = sms * v.ctas;
    // occupancy is set by shared memory: pad the request so that exactly `ctas` CTAs fit
    size_t want = (227 * 1024) / v.ctas - 1024;
    if (want < sm) { printf("{\"skip\": \"smem\"}\n"); continue; }
    want = want > 227 * 1024 ? 227 * 1024 : want;
    float best = 1e30f, tot = 0;
    const int iters = 5;
    for (int it = 0; it < iters + 1; ++it) {
      CK(cudaEventRecord(e0));
      ring<<<grid, 32 * v.warps, want>>>(p);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it) { tot += ms; best = ms < best ? ms : best; }
    }
    const double traffic = (v.mode == 0 ? 2.0 : 1.0) * (double)bytes;
    printf("{\"tile\": %u, \"stages\": %u, \"ctas\": %d, \"mode\": %d, \"map\": %d, \"lhint\": %d, \"shint\": %d, \"split\": %d, \"inplace\": %d, \"warps\": %d, \"ms\": %.4f, \"GBps_mean\": %.1f, \"GBps_best\": %.1f}\n",
           v.tile, v.stages, v.ctas, v.mode, v.map, v.lhint, v.shint, v.split, v.inplace, v.warps, tot / iters, traffic / (tot / iters) / 1e6, traffic / best / 1e6);
    fflush(stdout);
  }
  if (argc > 3) {
    for (int inplace = 1; inplace >= 0; --inplace)
      for (int threads : {256, 512, 1024})
        for (int per_sm : {1, 2, 4, 8}) {
          if (threads * per_sm > 2048) continue;
          const int grid = sms * per_sm;
          uint8_t* d = inplace ? a : b;
          const float m1 = run_plain<1>(a, d, bytes, threads, grid, e0, e1), m2 = run_plain<2>(a, d, bytes, threads, grid, e0, e1), m4 = run_plain<4>(a, d, bytes, threads, grid, e0, e1),
                      m8 = run_plain<8>(a, d, bytes, threads, grid, e0, e1);
          printf("{\"plain\": 1, \"inplace\": %d, \"threads\": %d, \"ctas\": %d, \"GBps_U1\": %.1f, \"GBps_U2\": %.1f, \"GBps_U4\": %.1f, \"GBps_U8\": %.1f}\n", inplace, threads, per_sm,
                 2.0 * bytes / m1 / 1e6, 2.0 * bytes / m2 / 1e6, 2.0 * bytes / m4 / 1e6, 2.0 * bytes / m8 / 1e6);
          fflush(stdout);
        }
  }
  // reference points: cudaMemcpy D2D (out of place) and memset
  for (int it = 0; it < 3; ++it) {
    CK(cudaEventRecord(e0));
    CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"cudaMemcpyD2D_GBps\": %.1f}\n", 2.0 * bytes / ms / 1e6);
  }
  return 0;
}
