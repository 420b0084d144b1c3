// Profiling harness (not part of the library): how many bytes per clock can every SM pull from L2 through
// TMA when all CTAs stream the SAME weight tiles -- unicast vs cluster multicast (2, 4 CTAs).  Decides whether
// the conv-chain kernel's weight stream (416 KB per 128-pixel tile) can be shared across a cluster.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I deepfly3d_b200/csrc -I include tools/l2_probe.cu \
//        -L deepfly3d_b200 -ldf3d_b200 -Xlinker -rpath='$ORIGIN/../deepfly3d_b200' -o tools/l2_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv_gemm.cuh"
#include "sm100.cuh"

using namespace df3d;
using namespace df3d::sm100;

#define CK(x)                                                   \
  do {                                                          \
    cudaError_t e = (x);                                        \
    if (e != cudaSuccess) {                                     \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));     \
      return 1;                                                 \
    }                                                           \
  } while (0)

constexpr int kSlots = 4, kSlotBytes = 32768, kUnit = 16384, kChunks = 13;

struct Params {
  CUtensorMap tm;  // (K, 128) bf16, box (64, 128 / csz)
  int csz, iters, distinct;
  unsigned long long* cycles;
};

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + kSlots * kSlotBytes;
  auto full = [&](int s) { return bar + 8u * s; };
  auto empty = [&](int s) { return bar + 8u * (kSlots + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int csz = p.csz;
  const uint32_t rank = csz > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), csz);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (csz > 1) cluster_sync();
  const int rows = 128 / csz;
  const uint16_t mask = (uint16_t)((1u << csz) - 1u);
  const int kofs = p.distinct ? (int)(blockIdx.x % 8) * kChunks * 128 : 0;  // distinct: 8 different weight sets
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    uint32_t u = 0, ph = 0;
    for (int it = 0; it < p.iters; ++it)
      for (int c = 0; c < kChunks; ++c) {
        mbar_wait(empty(u), ph ^ 1u);
        mbar_arrive_expect_tx(full(u), kSlotBytes);
        const uint32_t dst = base + u * kSlotBytes + rank * rows * 128;
        if (csz > 1) {
          tma_load_2d_mc(dst, &p.tm, full(u), kofs + (2 * c) * 64, rank * rows, mask);
          tma_load_2d_mc(dst + kUnit, &p.tm, full(u), kofs + (2 * c + 1) * 64, rank * rows, mask);
        } else {
          tma_load_2d(dst, &p.tm, full(u), kofs + (2 * c) * 64, 0);
          tma_load_2d(dst + kUnit, &p.tm, full(u), kofs + (2 * c + 1) * 64, 0);
        }
        if (++u == kSlots) {
          u = 0;
          ph ^= 1u;
        }
      }
  } else if (warp == 1 && lane == 0) {
    uint32_t u = 0, ph = 0;
    for (int it = 0; it < p.iters; ++it)
      for (int c = 0; c < kChunks; ++c) {
        mbar_wait(full(u), ph);
        if (csz > 1) {
          for (int r = 0; r < csz; ++r) mbar_arrive_remote(empty(u), r);
        } else {
          mbar_arrive(empty(u));
        }
        if (++u == kSlots) {
          u = 0;
          ph ^= 1u;
        }
      }
  }
  __syncthreads();
  if (csz > 1) cluster_sync();
  if (threadIdx.x == 0) p.cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 200;
  if (tma_init()) {
    printf("tma_init failed\n");
    return 1;
  }
  const int K = kChunks * 128 * 8;  // 8 weight sets for the "distinct" variant
  __nv_bfloat16* w;
  unsigned long long* cyc;
  CK(cudaMalloc(&w, (size_t)K * 128 * 2));
  CK(cudaMemset(w, 0, (size_t)K * 128 * 2));
  CK(cudaMalloc(&cyc, 148 * 8));
  const int smem = kSlots * kSlotBytes + 1024 + 256;
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int distinct = 0; distinct < 2; ++distinct)
    for (int csz : {1, 2, 4, 8}) {
      if (distinct && csz > 1) continue;
      Params p;
      memset(&p, 0, sizeof(p));
      if (make_tmap_wgt(&p.tm, w, K, 128, 128 / csz)) {
        printf("tmap failed: %s\n", df3d_last_error());
        return 1;
      }
      p.csz = csz;
      p.iters = iters;
      p.distinct = distinct;
      p.cycles = cyc;
      const int grid = (148 / csz) * csz;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csz;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        CK(cudaLaunchKernelEx(&cfg, stream_kernel, p));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        unsigned long long h[148];
        CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < grid; ++i) avg += (double)h[i];
        avg /= grid;
        const double bytes = (double)iters * kChunks * kSlotBytes;
        printf("distinct %d cluster %d grid %d rep %d: %.3f ms, %.1f B/clk/SM delivered, %.2f TB/s chip delivered\n", distinct,
               csz, grid, rep, ms, bytes / avg, bytes * grid / (ms * 1e-3) / 1e12);
      }
    }
  return 0;
}
