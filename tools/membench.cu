// membench.cu -- access-pattern probe for the quant_dim = 1 (column-grouped) kernels: in-place copy kernels that
// touch a [rows, cols] fp16 matrix exactly like the candidate kernels do, without any arithmetic.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu
//   run:   tools/membench rows cols
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg128(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// mode 0: flat streaming (tile = 16 KB contiguous), 4 x 16 B per thread in flight
// mode 1: column tile g x C, warp w owns a strip of C/8 columns, lpr lanes per strip row (col_reg pattern)
// mode 2: column tile g x 256, warp w takes rows w, w+8, ... reading 512 contiguous bytes per row (col_kernel pattern)
// order 0: column tile fastest; order 1: row block fastest
template <int MODE, int RPL>
__global__ void __launch_bounds__(256) probe(uint8_t* w, uint32_t rows, uint32_t cols, uint32_t g, uint32_t C, int order, int touch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row_bytes = (size_t)cols * 2;
  if (MODE == 0) {
    const size_t n_tiles = (size_t)rows * row_bytes / 16384;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      uint8_t* base = w + t * 16384 + threadIdx.x * 16;
      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = ldg128(base + i * 4096);
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[i].x ^= touch; stg128(base + i * 4096, v[i]); }
    }
    return;
  }
  const uint32_t n_ct = cols / C, n_rb = rows / g;
  const uint32_t total = n_ct * n_rb;
  for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
    uint32_t rb, ct;
    if (order == 0) { rb = t / n_ct; ct = t - rb * n_ct; } else { ct = t / n_rb; rb = t - ct * n_rb; }
    if (MODE == 1) {
      const uint32_t lpr = C >> 6, rps = 32 / lpr;
      const uint32_t sub = lane & (lpr - 1), rph = lane / lpr;
      const uint32_t col = ct * C + warp * (C >> 3) + sub * 8;
      uint8_t* base = w + ((size_t)rb * g + rph) * row_bytes + (size_t)col * 2;
      const size_t step = (size_t)rps * row_bytes;
      uint4 v[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) v[i] = ldg128(base + i * step);
#pragma unroll
      for (int i = 0; i < RPL; ++i) { v[i].x ^= touch; stg128(base + i * step, v[i]); }
    } else {
      uint8_t* base = w + ((size_t)rb * g + warp) * row_bytes + (size_t)(ct * 256 + lane * 8) * 2;
      const size_t step = 8 * row_bytes;
      for (uint32_t k = 0; k < g / 8; k += RPL) {
        uint4 v[RPL];
#pragma unroll
        for (int i = 0; i < RPL; ++i) v[i] = ldg128(base + (k + i) * step);
#pragma unroll
        for (int i = 0; i < RPL; ++i) { v[i].x ^= touch; stg128(base + (k + i) * step, v[i]); }
      }
    }
  }
}

// L2 residency probe: read-only pass over `bytes` (sum kept alive through a never-taken store)
__global__ void __launch_bounds__(256) read_only(const uint8_t* w, size_t bytes, uint32_t* sink) {
  uint32_t acc = 0;
  for (size_t off = ((size_t)blockIdx.x * 256 + threadIdx.x) * 16; off < bytes; off += (size_t)gridDim.x * 256 * 16) {
    const uint4 v = *reinterpret_cast<const uint4*>(w + off);
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <class F>
static void run(const char* name, size_t bytes, F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch();
  cudaEventRecord(e0);
  for (int i = 0; i < 10; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 10;
  printf("%-44s %8.3f ms  %7.0f GB/s (read+write)  %s\n", name, ms, 2.0 * bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
  const uint32_t rows = argc > 1 ? atoi(argv[1]) : 8192, cols = argc > 2 ? atoi(argv[2]) : 8192;
  const int n_mat = argc > 3 ? atoi(argv[3]) : 8;
  const size_t bytes = (size_t)rows * cols * 2;
  uint8_t* w;
  cudaMalloc(&w, bytes * n_mat);
  cudaMemset(w, 0, bytes * n_mat);
  printf("matrix [%u, %u] x %d (%.1f MB each)\n", rows, cols, n_mat, bytes / 1e6);
  if (argc > 4) {  // tools/membench rows cols n l2: how much of a region just read is still in L2 when it is read again?
    uint32_t* sink;
    cudaMalloc(&sink, 4);
    const size_t total = bytes * n_mat;
    for (size_t mb : {8, 16, 24, 32, 48, 64, 96, 128, 192}) {
      const size_t sz = mb << 20;
      if (sz > total) break;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      float best = 1e9f;
      for (int rep = 0; rep < 5; ++rep) {
        read_only<<<148 * 8, 256>>>(w + total - (256ull << 20), 256ull << 20, sink);  // evict: 256 MB of something else
        read_only<<<148 * 8, 256>>>(w, sz, sink);                                        // warm
        cudaEventRecord(e0);
        read_only<<<148 * 8, 256>>>(w, sz, sink);                                        // re-read (timed)
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("re-read of %4zu MB right after reading it: %7.3f ms  %7.0f GB/s\n", mb, best, sz / best / 1e6);
    }
    {  // read-only streaming ceiling over the whole buffer (the per-tensor range pass is such a stream)
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      for (int ctas : {2, 4, 8}) {
        read_only<<<148 * ctas, 256>>>(w, total, sink);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) read_only<<<148 * ctas, 256>>>(w, total, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("read-only stream of %zu MB, %d CTAs/SM: %7.0f GB/s\n", total >> 20, ctas, total * 5 / ms / 1e6);
      }
    }
    return 0;
  }
  auto each = [&](auto f) { for (int m = 0; m < n_mat; ++m) f(w + (size_t)m * bytes); };
  const size_t tot = bytes * n_mat;
  for (int ctas : {1, 2, 3, 4}) {
    char nm[128];
    const int grid = 148 * ctas;
    snprintf(nm, sizeof nm, "flat 16KB tiles, %d CTA/SM", ctas);
    run(nm, tot, [&] { each([&](uint8_t* p) { probe<0, 4><<<grid, 256>>>(p, rows, cols, 128, 256, 0, 0); }); });
  }
  for (int order : {0, 1}) for (int ctas : {1, 2, 3, 4}) {
    char nm[128];
    const int grid = 148 * ctas;
    snprintf(nm, sizeof nm, "strip 64B x8 rows, RPL16, %d CTA/SM, order %d", ctas, order);
    run(nm, tot, [&] { each([&](uint8_t* p) { probe<1, 16><<<grid, 256>>>(p, rows, cols, 128, 256, order, 0); }); });
    snprintf(nm, sizeof nm, "strip 32B x16 rows, RPL8, %d CTA/SM, order %d", ctas, order);
    run(nm, tot, [&] { each([&](uint8_t* p) { probe<1, 8><<<grid, 256>>>(p, rows, cols, 128, 128, order, 0); }); });
    snprintf(nm, sizeof nm, "row 512B, 4 in flight, %d CTA/SM, order %d", ctas, order);
    run(nm, tot, [&] { each([&](uint8_t* p) { probe<2, 4><<<grid, 256>>>(p, rows, cols, 128, 256, order, 0); }); });
    snprintf(nm, sizeof nm, "row 512B, 16 in flight, %d CTA/SM, order %d", ctas, order);
    run(nm, tot, [&] { each([&](uint8_t* p) { probe<2, 16><<<grid, 256>>>(p, rows, cols, 128, 256, order, 0); }); });
  }
  return 0;
}
