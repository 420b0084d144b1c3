// Profiling harness (not part of the library): runs one conv-chain launch on synthetic data with
// the kernel's time-stamp hook enabled and prints the per-tile timeline of CTA 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I deepfly3d_b200/csrc tools/chain_probe.cu \
//        -L deepfly3d_b200 -ldf3d_b200 -Xlinker -rpath=$PWD/deepfly3d_b200 -o gpurun_out/chain_probe
//   gpurun_out/chain_probe [head_taps=9] [n_images=224] [variant: 0=T1 1=T3 2=head+c3]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv_chain.cuh"
#include "conv_gemm.cuh"

using namespace df3d;

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e = (x);                                                      \
    if (e != cudaSuccess) {                                                   \
      printf("%s failed: %s\n", #x, cudaGetErrorString(e));                   \
      return 1;                                                               \
    }                                                                         \
  } while (0)

int main(int argc, char** argv) {
  const int taps = argc > 1 ? atoi(argv[1]) : 9;
  const int B = argc > 2 ? atoi(argv[2]) : 224;
  const int variant = argc > 3 ? atoi(argv[3]) : 0;
  const int nores = argc > 6 ? atoi(argv[6]) : 0;  // 1: no residual slabs at all (timing experiment)
  const int H = 64, W = 64;
  if (tma_init() || conv_chain_configure()) {
    printf("init failed: %s\n", df3d_last_error());
    return 1;
  }
  const size_t px = (size_t)B * H * W;
  __nv_bfloat16 *t1, *res, *y, *t1n, *half, *wts;
  float* aff;
  unsigned long long* dbg;
  CK(cudaMalloc(&t1, px * 128 * 2));
  CK(cudaMalloc(&res, px * 256 * 2));
  CK(cudaMalloc(&y, px * 256 * 2));
  CK(cudaMalloc(&t1n, px * 128 * 2));
  CK(cudaMalloc(&half, px / 4 * 256 * 2));
  CK(cudaMalloc(&wts, (size_t)(9 * 128 * 128 + 4 * 256 * 256) * 2));
  CK(cudaMalloc(&aff, 4096 * 4));
  CK(cudaMalloc(&dbg, 4 * 4096 * 8));
  CK(cudaMemset(t1, 0, px * 128 * 2));
  CK(cudaMemset(res, 0, px * 256 * 2));
  CK(cudaMemset(half, 0, px / 4 * 256 * 2));
  CK(cudaMemset(wts, 0, (size_t)(9 * 128 * 128 + 4 * 256 * 256) * 2));
  CK(cudaMemset(aff, 0, 4096 * 4));
  CK(cudaMemset(dbg, 0, 4 * 4096 * 8));

  ChainParams p;
  memset(&p, 0, sizeof(p));
  const int halo = (argc > 5 ? atoi(argv[5]) : 1) && taps == 9;
  const int tw = halo ? 8 : 16, th = halo ? 16 : 8, nb = 1;
  int rc = make_tmap_act(&p.tmA, t1, 128, W, H, B, tw, th, nb);
  if (halo) {
    p.halo = 1;
    rc |= make_tmap_box(&p.tmHalo, t1, 128, W, H, B, tw + 2, th + 2, 1);
  }
  p.taps = taps;
  p.kc_per_tap = 2;
  p.H = H;
  p.W = W;
  p.B = B;
  p.tw = tw;
  p.th = th;
  p.nb = nb;
  p.tiles_x = W / tw;
  p.tiles_y = H / th;
  p.tiles_b = B;
  p.dbg = dbg;
  p.dbg_exec = argc > 4 ? atoi(argv[4]) : 0;
  int n = 0;
  auto stage = [&](int K, int N, size_t woff) -> ChainStage& {
    ChainStage& st = p.st[n++];
    rc |= make_tmap_wgt(&st.tmB, wts + woff, K, N, 64);
    st.n = N;
    st.kblocks = K / 64;
    st.scale1 = aff;
    st.shift1 = aff;
    st.scale2 = aff;
    st.shift2 = aff;
    return st;
  };
  const char* names[8];
  if (taps == 9) {  // head 3x3
    ChainStage& s0 = stage(9 * 128, 128, 0);
    s0.relu1 = 1;
    s0.x_src = 1;
    names[0] = "3x3";
  }
  {  // conv3 + residual (+ up)
    ChainStage& s1 = stage(128, 256, 9 * 128 * 128);
    s1.unit_scale = 1;
    s1.has_res = 1;
    rc |= make_tmap_act(&s1.tmRes, res, 256, W, H, B, tw, th, nb);
    if (variant == 0 && nores < 2) {
      s1.has_res2 = 1;
      rc |= make_tmap_box(&s1.tmRes2, half, 256, W / 2, H / 2, B, tw / 2, th / 2, nb);
    }
    names[n - 1] = "c3";
    if (variant == 1) {
      s1.x_src = 1;  // r -> fc
      ChainStage& s2 = stage(256, 256, 9 * 128 * 128 + 256 * 128);
      s2.relu1 = 1;
      s2.x_src = 1;
      names[n - 1] = "fc";
      ChainStage& s3 = stage(256, 256, 9 * 128 * 128 + 256 * 128 + 256 * 256);
      s3.unit_scale = 1;
      s3.has_res = 1;
      rc |= make_tmap_act(&s3.tmRes, res, 256, W, H, B, tw, th, nb);
      s3.out_raw = y;
      rc |= make_tmap_quarter(&s3.tmOutQ, y, 256, W, H, B, tw, th, nb);
      s3.x_src = 2;
      names[n - 1] = "merged";
    } else {
      s1.out_raw = y;
      rc |= make_tmap_quarter(&s1.tmOutQ, y, 256, W, H, B, tw, th, nb);
      s1.x_src = variant == 2 ? 0 : 2;
    }
  }
  if (variant != 2) {
    ChainStage& s2 = stage(256, 128, 9 * 128 * 128 + 256 * 128 + 2 * 256 * 256);
    s2.relu1 = 1;
    s2.out_raw = t1n;
    rc |= make_tmap_quarter(&s2.tmOutQ, t1n, 128, W, H, B, tw, th, nb);
    names[n - 1] = "c1'";
  }
  p.n_chain = n;
  if (rc) {
    printf("tensor map failed: %s\n", df3d_last_error());
    return 1;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    if (launch_conv_chain(p, 148, 0)) {
      printf("launch failed: %s\n", df3d_last_error());
      return 1;
    }
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const int tiles = p.tiles_x * p.tiles_y * p.tiles_b;
    printf("run %d: %.3f ms, %d tiles, %.2f us/tile/SM\n", it, ms, tiles, ms * 1e3 / ((tiles + 147) / 148));
  }
  std::vector<unsigned long long> h(4 * 4096);
  CK(cudaMemcpy(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost));
  // MMA warp stamps: head(0) start/issued, then per tile: for each stage i >= 1 (ready, issued), with the next
  // head's (start, issued) behind stage head_after.  Epilogue stamps: (ready, done) per stage in its own order.
  const int per = 2 * n;  // stamps per tile in steady state, both roles
  const int per_m = per + ((p.dbg_exec & 1) ? n - 1 : 0);
  printf("MMA warp, tiles 20..23: deltas in clock cycles between consecutive stamps (period = %d stamps per tile)\n", per);
  for (int t = 20; t < 24; ++t) {
    printf("  tile %d:", t);
    for (int j = 0; j < per_m; ++j) {
      const size_t k = 2 + (size_t)t * per_m + j;
      printf(" %llu", h[k] - h[k - 1]);
    }
    printf("   => %llu per tile\n", h[2 + (size_t)(t + 1) * per_m] - h[2 + (size_t)t * per_m]);
  }
  for (int g = 0; g < 2; ++g) {
    printf("epilogue group %d, tiles 20..23: (loop overhead, barrier wait, busy) per stage in program order\n", g);
    for (int t = 20; t < 24; ++t) {
      printf("  tile %d:", t);
      for (int j = 0; j < n; ++j) {
        const size_t k = 4096 * (1 + g) + 3 + (size_t)t * 3 * n + 3 * j;
        printf(" (%llu, %llu, %llu)", h[k] - h[k - 1], h[k + 1] - h[k], h[k + 2] - h[k + 1]);
      }
      printf("\n");
    }
  }
  if (p.dbg_exec & 4) {
    // group 0's slabs in program order: per tile (steady state) stage 1 has n1/128 slabs of this group, the head and the
    // other stages n/128 each
    int per_tile = 0;
    for (int i = 0; i < n; ++i) per_tile += p.st[i].n >> 7;
    printf("epilogue group 0, slabs of tiles 20..21 (%d slabs per tile): (wait for the slab, tcgen05.ld + math, operand hand-over, store) and the gap to the next slab\n", per_tile);
    for (int sidx = 20 * per_tile; sidx < 22 * per_tile; ++sidx) {
      const size_t k = 3 * 4096 + (size_t)sidx * 5;
      printf("  slab %d: (%llu, %llu, %llu, %llu) gap %llu\n", sidx, h[k + 1] - h[k], h[k + 2] - h[k + 1], h[k + 3] - h[k + 2], h[k + 4] - h[k + 3],
             h[k + 5] - h[k + 4]);
    }
  }
  return 0;
}
