// tcgen05 convolution-chain kernel (see conv_chain.cuh).
//
// One persistent CTA per SM, 12 warps, 128-pixel tiles; the two SMs of a TPC work as a CTA PAIR
// (cluster of 2, tcgen05 cta_group::2): every MMA is M = 256 -- 128 pixels of each CTA's own tile -- by
// N = 128, and the weights, the operand the two tiles share, are split between the CTAs: each CTA fetches
// and holds only 64 of the 128 weight rows of an MMA.  Measured on B200 (tools/chain_probe, tools/l2_probe):
// the single-CTA form is bound by the 128 B/clk shared-memory port (an M128 x N128 SS MMA reads 8 KB per
// 64 clocks by itself, plus the TMA writes of the weight stream); the pair halves the weight traffic.
//   warp 0  TMA producer of the operand ring, in the MMA warp's issue order: per 64-wide K block of the head
//           one A tile (a 4-D box per tap and channel block; the 3x3 halo is the TMA's out-of-bounds
//           zero fill) + this CTA's half of the head's weights; per later stage 16 KB slots of weights.
//           Completion is signalled on the LEADER CTA's barriers (the leader issues the MMAs of the pair)
//   warp 1  leader CTA (rank 0): MMA issuer; both CTAs: TMEM owner (warp-uniform loop, one elected lane issues)
//   warp 2  TMA producer of the residual slabs (64 channels x 128 pixels, + the half-resolution slab
//           of the up-sample branch), a ring that prefetches across stages and tiles
//   warp 3  idle.  Warps 0..3 (warpgroup 0) give registers back (setmaxnreg 104) so that the epilogue
//           warpgroups can take 200 each: the heavy epilogues spilled at the 168 registers a 384-thread
//           block allows, and with 227 KB of shared memory (no L1) a spill reload is an L2 round trip
//   warps 4..11  epilogue: two groups of four warps (one TMEM lane quarter each), group g takes the
//           64-channel slabs sl = g, g+2 of every stage; inside a warp, lanes l and l+16 share two pixel
//           rows and split their channels (see epi_slab)
//
// Stage i of a tile:
//   MMA      D(acc_i) = A_i * W_i^T;  A_0 from shared memory (TMA), A_i (i >= 1) from tensor memory.  The MMAs
//            over K block k of stage i+1 wait only for slab k of stage i's epilogue (one barrier per slab),
//            so they overlap the epilogue of the remaining slabs
//   epilogue (each CTA for its own tile, out of its own tensor memory; arrivals are counted on the leader's
//            barriers: 4 warps of each CTA per slab, 8 per stage) tcgen05.ld acc_i -> scale/shift
//            (+ residuals) -> ReLU / bf16 rounding
//            -> stored stages with a residual: bf16(v) written in place into the residual slab, which then
//               leaves with one TMA store per warp quarter; stored stages without: 256-bit stores from registers
//            -> optional next-BatchNorm + ReLU -> tcgen05.st of the bf16 operand of stage i+1 IN PLACE
//               over the first half of the accumulator columns it was computed from
// Tensor memory holds three regions, P = [0,128), Q = [128,384), R = [384,512); launch_conv_chain maps
// every accumulator (two 128-column halves for 256 channels) onto them so that a region is only
// rewritten after its last reader.  The chain of one tile is sequential, but the issue ORDER is
// software-pipelined: the head GEMM of tile t+1 is issued right behind stage `head_after` of tile t
// (the stage with the longest epilogue), so the tensor pipe works on the next 3x3 conv while the
// epilogue warps apply residual / BatchNorm / stores of this one.
#include <cstdlib>
#include <cstring>
#include <initializer_list>

#include "conv_chain.cuh"
#include "conv_gemm.cuh"  // tensor-map helpers
#include "epilogue.cuh"
#include "sm100.cuh"

namespace df3d {

using namespace sm100;

constexpr int kChainThreads = 384;  // warpgroup 0: producer / MMA / slab producer / idle; warpgroups 1, 2: epilogue
constexpr int kEpiWarp0c = 4;
constexpr int kUnitBytes = 16384;       // 128 rows x 64 bf16: one A tile
constexpr int kSubBytes = 8192;         // 64 rows x 64 bf16: this CTA's half of the weights of one N = 128, K = 64 block
constexpr int kMaxM = 12, kMaxSlabs = 8;  // ring slots, residual slabs
constexpr int kHaloPitch = 10;               // halo row = 8 tile pixels + one on each side
constexpr int kHaloBytes = 18 * kHaloPitch * 128;  // one 64-channel half of the halo of an 8 x 16 tile (45 KB)
constexpr int kColP = 0, kColQ = 128, kColR = 384;
constexpr int kChainSmemLimit = 232448;  // 227 KB opt-in maximum per CTA
constexpr int kChainBarBytes = 1280;  // barriers, TMEM slot, StageLite table
// specialised epilogues (see epi_slab)
enum { kEpiReluX = 0, kEpiReluOut, kEpiResOutAct, kEpiResUpOutAct, kEpiResOut, kEpiResUpOut, kEpiResX };

// Per-stage fields the MMA and epilogue warps read every tile, copied to shared memory once: indexed
// loads from the kernel-parameter constant bank miss the small constant cache and cost ~1000 cycles
// per stage when done in sequence.
struct StageLite {
  unsigned long long out;  // bf16 output base or 0
  unsigned long long pool_raw, pool_act;  // pooled outputs or 0
  int n, kblocks, has_res, x_src, kind, col[2][2], aff_off, unit, hz, hzd, pool_off, ssk, stg;
};
static_assert(sizeof(StageLite) == 88, "StageLite layout");
constexpr int kLiteStride = 96;
static_assert(8 * (2 * kMaxM + 2 * kMaxSlabs + 2 * kMaxChain) + 32 + 8 * 4 * kMaxChain + kMaxChain * kLiteStride <= kChainBarBytes, "barrier area");

// kProbe: the time-stamp probe of tools/chain_probe.cu (launch_conv_chain picks that instantiation when
// ChainParams::dbg is set); the production instantiation carries none of its branches.
template <bool kProbe>
__global__ void __launch_bounds__(kChainThreads, 1) conv_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const sm = smem_raw + (smem_base - smem_u32(smem_raw));  // same place as a __shared__ pointer: plain
                                                                    // loads/stores the compiler may schedule
  const int nh0 = p.st[0].n >> 7;
  // ONE ring of equal slots feeds the tensor pipe in issue order: a head K block (A tile + its weight
  // half-tiles) or 32 KB of later-stage weights per slot
  const uint32_t slot_bytes = (uint32_t)p.slot_bytes;
  const uint32_t m_base = smem_base;
  // halo mode (3x3 head on maps of at least 16 x 8): the head's activations are not streamed per tap; the
  // (th+2) x (tw+2) pixel halo of the tile is loaded once per 64-channel half and the nine taps are
  // row-shifted views of it (see issue_head); the ring then carries the head's weights only
  const uint32_t halo_base = m_base + (uint32_t)p.n_m * slot_bytes;
  const uint32_t s_base = halo_base + (p.halo ? 2u * kHaloBytes : 0u);
  const uint32_t aff_base = s_base + (uint32_t)p.n_slabs * (uint32_t)p.slab_bytes;
  const uint32_t bar_base = aff_base + (uint32_t)p.aff_bytes;
  auto mfull = [&](uint32_t s) { return bar_base + 8u * s; };
  auto mempty = [&](uint32_t s) { return bar_base + 8u * (kMaxM + s); };
  auto sfull = [&](uint32_t s) { return bar_base + 8u * (2 * kMaxM + s); };
  auto sempty = [&](uint32_t s) { return bar_base + 8u * (2 * kMaxM + kMaxSlabs + s); };
  auto accfull = [&](uint32_t i) { return bar_base + 8u * (2 * kMaxM + 2 * kMaxSlabs + i); };
  auto epidone = [&](uint32_t i) { return bar_base + 8u * (2 * kMaxM + 2 * kMaxSlabs + kMaxChain + i); };
  const uint32_t hfull = bar_base + 8u * (2 * kMaxM + 2 * kMaxSlabs + 2 * kMaxChain);
  const uint32_t hempty = hfull + 8u;
  const uint32_t tmem_slot = hfull + 16u;
  // operand K block sl (= 64-channel slab sl of the stage's output) written to tensor memory by all its warps:
  // the next stage's MMAs over that K block may go while the epilogue still works on the other slabs
  auto epislab = [&](uint32_t i, uint32_t sl) { return hfull + 32u + 8u * (4u * i + sl); };
  const StageLite* const lite0 = reinterpret_cast<const StageLite*>(sm + (bar_base - smem_base) + 8 * (2 * kMaxM + 2 * kMaxSlabs + 2 * kMaxChain) + 32 + 8 * 4 * kMaxChain);
  auto lite = [&](int i) -> const StageLite& { return *reinterpret_cast<const StageLite*>(reinterpret_cast<const uint8_t*>(lite0) + i * kLiteStride); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_b;
  // loop-invariant scalars of the parameter block, pinned in registers: left to itself the compiler re-reads them
  // from the constant bank inside the per-stage loops, and with a 4.4 KB parameter block those loads miss the
  // small constant cache (ncu: long-scoreboard / branch-resolving samples on LDCU c[0x0][...] in every role loop)
  int n_chain = p.n_chain, head_after = p.head_after;
  asm volatile("" : "+r"(n_chain), "+r"(head_after));
  // CTA pair: rank 0 (the leader) issues every MMA; pair-tile pt = tiles 2 pt (leader) and 2 pt + 1 (peer); a
  // tile past the end is a phantom (its loads are out of bounds = zeros, its stores are skipped)
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pt = (total_tiles + 1) >> 1;
  const int pt0 = (int)(blockIdx.x >> 1), pstride = (int)(gridDim.x >> 1);
  // the leader's copies of the barriers both CTAs signal (shared::cluster addresses)
  const uint32_t lead_bar = mapa_cluster(bar_base, 0);
  auto mfull_l = [&](uint32_t s) { return lead_bar + 8u * s; };
  auto epidone_l = [&](uint32_t i) { return lead_bar + 8u * (2 * kMaxM + 2 * kMaxSlabs + kMaxChain + i); };
  auto epislab_l = [&](uint32_t i, uint32_t sl) { return lead_bar + 8u * (2 * kMaxM + 2 * kMaxSlabs + 2 * kMaxChain) + 32u + 8u * (4u * i + sl); };
  const uint32_t hfull_l = lead_bar + 8u * (2 * kMaxM + 2 * kMaxSlabs + 2 * kMaxChain);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&p.tmA);
    for (int i = 0; i < n_chain; ++i) {
      prefetch_tensormap(&p.st[i].tmB);
      if (p.st[i].has_res) prefetch_tensormap(&p.st[i].tmRes);
      if (p.st[i].has_res2) prefetch_tensormap(&p.st[i].tmRes2);
      if ((p.st[i].has_res || p.st[i].stage_out) && p.st[i].out_raw) prefetch_tensormap(&p.st[i].tmOutQ);
      if (p.st[i].ss_kblocks) prefetch_tensormap(&p.st[i].tmA2);
    }
    for (int s = 0; s < p.n_m; ++s) {
      mbar_init(mfull(s), 1);
      mbar_init(mempty(s), 1);
    }
    for (int s = 0; s < p.n_slabs; ++s) {
      mbar_init(sfull(s), 1);
      mbar_init(sempty(s), 4);  // one arrive per warp of the epilogue group that read it
    }
    for (int i = 0; i < n_chain; ++i) {
      mbar_init(accfull(i), 1);
      mbar_init(epidone(i), 16);  // one arrive per epilogue warp of BOTH CTAs (only the leader's copy is used)
      for (int sl = 0; sl < 4; ++sl) mbar_init(epislab(i, sl), 8);  // the four warps of the slab's group, both CTAs
    }
    mbar_init(hfull, 1);
    mbar_init(hempty, 1);
    if (p.halo) prefetch_tensormap(&p.tmHalo);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair<512>(tmem_slot);
  if (warp == 2 && lane < n_chain) {
    const ChainStage& st = p.st[lane];
    StageLite* l = const_cast<StageLite*>(&lite(lane));
    l->out = reinterpret_cast<unsigned long long>(st.out_raw);
    l->n = st.n;
    l->kblocks = st.kblocks;
    l->has_res = st.has_res;
    l->x_src = st.x_src;
    l->kind = st.epi_kind;
    l->col[0][0] = st.col[0][0];
    l->col[0][1] = st.col[0][1];
    l->col[1][0] = st.col[1][0];
    l->col[1][1] = st.col[1][1];
    l->ssk = st.ss_kblocks;
    l->aff_off = st.aff_off;
    l->unit = st.unit_scale;
    l->hz = st.hz_stage;
    l->hzd = st.hz_delta;
    l->pool_raw = reinterpret_cast<unsigned long long>(st.pool_raw);
    l->pool_act = reinterpret_cast<unsigned long long>(st.pool_act);
    l->pool_off = st.pool_off;
    l->stg = st.stage_out;
  }
  // per-channel epilogue constants of every stage -> shared memory, only the arrays the stage uses:
  // [scale1 n (unless it is 1)][shift1 n][scale2 n][shift2 n (when the operand is relu(bn2(.)))]
  for (int i = 0; i < n_chain; ++i) {
    const ChainStage& st = p.st[i];
    uint32_t a0 = aff_base + 4u * (uint32_t)st.aff_off;
    if (!st.unit_scale) {
      for (int c = threadIdx.x; c < st.n; c += kChainThreads) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u * c), "f"(st.scale1[c]));
      a0 += 4u * st.n;
    }
    for (int c = threadIdx.x; c < st.n; c += kChainThreads) asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u * c), "f"(st.shift1[c]));
    if (st.x_src == 2) {
      for (int c = threadIdx.x; c < st.n; c += kChainThreads) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u * (st.n + c)), "f"(st.scale2[c]));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a0 + 4u * (2 * st.n + c)), "f"(st.shift2[c]));
      }
    }
    if (st.pool_raw) {
      const uint32_t p0 = aff_base + 4u * (uint32_t)st.pool_off;
      for (int c = threadIdx.x; c < st.n; c += kChainThreads) {
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(p0 + 4u * c), "f"(st.pool_scale[c]));
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(p0 + 4u * (st.n + c)), "f"(st.pool_shift[c]));
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / TMA completion
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  auto decode_tile = [&](int tile, int& x0, int& y0, int& n0) {
    int tx, ty, tb;
    if (p.tx_shift >= 0 && p.ty_shift >= 0) {
      tx = tile & (p.tiles_x - 1);
      ty = (tile >> p.tx_shift) & (p.tiles_y - 1);
      tb = tile >> (p.tx_shift + p.ty_shift);
    } else {
      tx = tile % p.tiles_x;
      tile /= p.tiles_x;
      ty = tile % p.tiles_y;
      tb = tile / p.tiles_y;
    }
    x0 = tx * p.tw;
    y0 = ty * p.th;
    n0 = tb * p.nb;
  };

  // register re-allocation between the warpgroups, inside the role branches so that ptxas budgets each role
  // separately (per scheduler: 104 + 2 x 200 = 504 registers per lane = the 3 x 168 the CTA was launched with: setmaxnreg moves registers inside the CTA's own pool, a larger sum blocks forever)
  if (warp < kEpiWarp0c) asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
  if (warp == 0) {
    // ------------------------------------------------------------------ ring producer (issue order of the MMA warp)
    if (lane == 0) {
      const int kb0 = p.st[0].kblocks;
      uint32_t u = 0, ph = 0;
      // `bytes` = what ONE CTA loads into the slot; the leader's barrier counts the bytes of both
      auto acquire = [&](uint32_t bytes) -> uint32_t {
        mbar_wait_cluster(mempty(u), ph ^ 1u);
        if (leader) mbar_arrive_expect_tx(mfull(u), 2u * bytes);
        return m_base + u * slot_bytes;
      };
      const int brow = (int)rank * 64;  // this CTA's 64 rows of every 128-row weight block
      auto advance = [&]() {
        if (++u == (uint32_t)p.n_m) {
          u = 0;
          ph ^= 1u;
        }
      };
      uint32_t heads = 0;
      auto load_head = [&](int tile) {
        int x0, y0, n0;
        decode_tile(tile, x0, y0, n0);
        if (p.halo) {  // the halo once, then one slot of weights per tap (both 64-channel K blocks, 64 rows each)
          mbar_wait_cluster(hempty, (heads & 1u) ^ 1u);
          if (leader) mbar_arrive_expect_tx(hfull, 4u * kHaloBytes);
          tma_load_4d_pair(halo_base, &p.tmHalo, hfull_l, 0, x0 - 1, y0 - 1, n0);
          tma_load_4d_pair(halo_base + kHaloBytes, &p.tmHalo, hfull_l, 64, x0 - 1, y0 - 1, n0);
          ++heads;
          for (int tap9 = 0; tap9 < 9; ++tap9) {
            const uint32_t dst = acquire(2 * kSubBytes);
            tma_load_2d_pair(dst, &p.st[0].tmB, mfull_l(u), (2 * tap9) * 64, brow);
            tma_load_2d_pair(dst + kSubBytes, &p.st[0].tmB, mfull_l(u), (2 * tap9 + 1) * 64, brow);
            advance();
          }
          return;
        }
        int tap = 0, kc = 0;
        for (int kb = 0; kb < kb0; ++kb) {
          int dx = 0, dy = 0;
          if (p.taps == 9) {
            dy = tap / 3 - 1;
            dx = tap - (tap / 3) * 3 - 1;
          }
          const uint32_t dst = acquire((uint32_t)(kUnitBytes + nh0 * kSubBytes));
          tma_load_4d_pair(dst, &p.tmA, mfull_l(u), kc * 64, x0 + dx, y0 + dy, n0);
          for (int h = 0; h < nh0; ++h)
            tma_load_2d_pair(dst + kUnitBytes + h * kSubBytes, &p.st[0].tmB, mfull_l(u), kb * 64, h * 128 + brow);
          advance();
          if (++kc == p.kc_per_tap) {
            kc = 0;
            ++tap;
          }
        }
      };
      // one slot = this CTA's 64 rows of both 128-row halves of one K block (256 outputs) or of two K blocks (128)
      auto load_weights = [&](int i, int tile) {
        const int kbn = lite(i).kblocks, nh = lite(i).n >> 7;
        // K blocks read from shared memory first (issue order of the MMA warp): one slot with this CTA's A tile of the
        // second activation tensor, one with its 64 rows of both 128-row weight halves
        if (const int ssk = lite(i).ssk) {
          int x0, y0, n0;
          decode_tile(tile, x0, y0, n0);
          for (int kb = 0; kb < ssk; ++kb) {
            uint32_t dst = acquire(kUnitBytes);
            tma_load_4d_pair(dst, &p.st[i].tmA2, mfull_l(u), kb * 64, x0, y0, n0);
            advance();
            dst = acquire(2 * kSubBytes);
            tma_load_2d_pair(dst, &p.st[i].tmB, mfull_l(u), (kbn + kb) * 64, brow);
            tma_load_2d_pair(dst + kSubBytes, &p.st[i].tmB, mfull_l(u), (kbn + kb) * 64, 128 + brow);
            advance();
          }
        }
        const int slots = (kbn * nh) >> 1;
        for (int sl = 0; sl < slots; ++sl) {
          const uint32_t dst = acquire(2 * kSubBytes);
          if (nh == 2) {
            tma_load_2d_pair(dst, &p.st[i].tmB, mfull_l(u), sl * 64, brow);
            tma_load_2d_pair(dst + kSubBytes, &p.st[i].tmB, mfull_l(u), sl * 64, 128 + brow);
          } else {
            tma_load_2d_pair(dst, &p.st[i].tmB, mfull_l(u), (2 * sl) * 64, brow);
            tma_load_2d_pair(dst + kSubBytes, &p.st[i].tmB, mfull_l(u), (2 * sl + 1) * 64, brow);
          }
          advance();
        }
      };
      // issue order (same walk in the MMA warp and in the epilogue warps): a virtual tile -1 runs only the
      // head of tile 0; step s of tile t is stage s+1 (s < head_after), the head of tile t+1
      // (s == head_after) or stage s (s > head_after)
      for (int t = -1, pt = pt0 - pstride;; ++t, pt += pstride) {
        const bool has_next = pt + pstride < n_pt;
        for (int sidx = 0; sidx < n_chain; ++sidx) {
          const bool head_step = (n_chain == 1) || (sidx == head_after);
          if (head_step) {
            if (has_next) load_head(2 * (pt + pstride) + (int)rank);
          } else if (t >= 0) {
            load_weights(sidx < head_after ? sidx + 1 : sidx, 2 * pt + (int)rank);
          }
        }
        if (!has_next) break;
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ residual slab producer
    if (lane == 0 && p.n_slabs > 0) {
      uint32_t u = 0, ph = 0;
      for (int pt = pt0; pt < n_pt; pt += pstride) {
        int x0, y0, n0;
        decode_tile(2 * pt + (int)rank, x0, y0, n0);
        for (int i = 0; i < n_chain; ++i) {
          const ChainStage& st = p.st[i];
          if (!st.has_res && !st.stage_out) continue;
          const int nsl = st.n >> 6;
          for (int sl = 0; sl < nsl; ++sl) {
            mbar_wait(sempty(u), ph ^ 1u);
            const uint32_t slab = s_base + u * (uint32_t)p.slab_bytes;
            if (st.stage_out) {  // blank slab: the epilogue stages its output in it
              mbar_arrive(sfull(u));
            } else {
              mbar_arrive_expect_tx(sfull(u), kUnitBytes + (st.has_res2 ? kUnitBytes / 4 : 0));
              tma_load_4d(slab, &st.tmRes, sfull(u), sl * 64, x0, y0, n0);
              if (st.has_res2) tma_load_4d(slab + kUnitBytes, &st.tmRes2, sfull(u), sl * 64, x0 >> 1, y0 >> 1, n0);
            }
            if (++u == (uint32_t)p.n_slabs) {
              u = 0;
              ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------------ MMA issuer (of the pair)
    // The whole warp runs the loops (warp-uniform control flow and operands: descriptors live in
    // uniform registers, no per-lane waterfall around each tcgen05 instruction); one elected lane issues.
    constexpr uint32_t idesc = umma_idesc_bf16(256, 128);  // 128 rows in each CTA
    const uint64_t desc_hi = umma_smem_desc_sw128(0);  // everything but the 14-bit start address
    const uint32_t n_m = (uint32_t)p.n_m;
    const int kb0 = p.st[0].kblocks;
    const int hz0 = p.st[0].hz_stage, hz0_delta = p.st[0].hz_delta;
    uint32_t mu = 0, mph = 0;
    unsigned long long* const dbg = (kProbe && blockIdx.x == 0 && lane == 0) ? p.dbg : nullptr;
    int di = 0;

    // Ring consumption with the barrier latency taken off the issue path: the readiness of the NEXT slot is
    // probed (mbarrier.try_wait) before the MMAs of the current one are issued, so the probe's round trip
    // overlaps the tensor work instead of sitting between two groups of MMAs.
    bool ring_ready = false;  // outcome of the early probe of slot `mu`
    auto ring_wait = [&]() {
      if (!ring_ready) mbar_wait_cluster(mfull(mu), mph);
      tc_fence_after();
    };
    auto ring_probe_next = [&]() {
      uint32_t nu = mu + 1, nph = mph;
      if (nu == n_m) {
        nu = 0;
        nph ^= 1u;
      }
      ring_ready = mbar_try_wait_cluster(mfull(nu), nph);
    };
    auto ring_advance = [&]() {
      if (++mu == n_m) {
        mu = 0;
        mph ^= 1u;
      }
    };

    // head GEMM of local tile number `t` (A and B from shared memory)
    const uint64_t desc_halo = (desc_hi & ~((uint64_t)0x3FFF << 32)) | ((uint64_t)((kHaloPitch * 128) >> 4) << 32);  // group pitch
    uint32_t heads = 0;
    auto issue_head = [&](int t) {
      const uint32_t d0 = tmem_base + (uint32_t)lite(0).col[t & 1][0], d1 = tmem_base + (uint32_t)lite(0).col[t & 1][1];
      if (kProbe && dbg && di < 4000) dbg[di++] = clock64();  // [head start]
      if (hz0 >= 0 && t - hz0_delta >= 0) mbar_wait_cluster(epidone(hz0), (uint32_t)(t - hz0_delta) & 1u);
      if (p.halo) {
        // A operand of tap (dy, dx) = the halo rows shifted by dy*10 + dx: pixel (r, c) of the 8-wide tile is
        // halo row (r+dy)*10 + (c+dx), i.e. 8-row groups 1280 B apart starting at a row offset.  The 128B
        // swizzle is a function of the absolute shared-memory address (what TMA wrote), so the shifted
        // start needs no descriptor base_offset (measured: bit-identical to nine separate TMA boxes).
        mbar_wait_cluster(hfull, heads & 1u);
        ++heads;
        tc_fence_after();
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          ring_wait();
          const int dy = tap / 3, dx = tap - dy * 3;
          const uint32_t a_addr = halo_base + (uint32_t)(dy * kHaloPitch + dx) * 128u;
          const uint64_t a0 = desc_halo | (uint64_t)((a_addr >> 4) & 0x3FFFu);
          const uint64_t a1 = desc_halo | (uint64_t)(((a_addr + kHaloBytes) >> 4) & 0x3FFFu);
          const uint32_t b_addr = m_base + mu * slot_bytes;
          const uint64_t b0 = desc_hi | (uint64_t)((b_addr >> 4) & 0x3FFFu);
          const uint64_t b1 = desc_hi | (uint64_t)(((b_addr + kSubBytes) >> 4) & 0x3FFFu);
          ring_probe_next();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_pair(d0, a0 + 2u * k, b0 + 2u * k, idesc, (tap | k) != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_pair(d0, a1 + 2u * k, b1 + 2u * k, idesc, 1u);
            umma_commit_pair(mempty(mu));
          }
          ring_advance();
        }
        if (elect_one()) {
          umma_commit_pair(hempty);  // the halo may be overwritten once these MMAs retire
          umma_commit_pair(accfull(0));
        }
        if (kProbe && dbg && di < 4000) dbg[di++] = clock64();  // [head issued]
        return;
      }
#pragma unroll 1
      for (int kb = 0; kb < kb0; ++kb) {
        ring_wait();
        const uint32_t a_addr = m_base + mu * slot_bytes;
        const uint64_t adesc = desc_hi | (uint64_t)((a_addr >> 4) & 0x3FFFu);
        const uint64_t b0 = desc_hi | (uint64_t)(((a_addr + kUnitBytes) >> 4) & 0x3FFFu);
        const uint64_t b1 = desc_hi | (uint64_t)(((a_addr + kUnitBytes + kSubBytes) >> 4) & 0x3FFFu);
        ring_probe_next();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 4 x (K = 16): +32 bytes inside the 128B swizzle atom
            umma_bf16_pair(d0, adesc + 2u * k, b0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          if (nh0 == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_pair(d1, adesc + 2u * k, b1 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_pair(mempty(mu));  // frees the ring stage (in both CTAs) once these MMAs retire
        }
        ring_advance();
      }
      if (elect_one()) umma_commit_pair(accfull(0));
      if (kProbe && dbg && di < 4000) dbg[di++] = clock64();  // [head issued]
    };

    // later stage i of local tile t: A from tensor memory (written in place by the previous epilogue), B from the ring
    auto issue_stage = [&](int i, int t) {
      const uint32_t par = (uint32_t)t & 1u;
      const StageLite& L = lite(i);
      const StageLite& Lp = lite(i - 1);
      const int kbn = L.kblocks, nh = L.n >> 7;
      const uint32_t c0 = tmem_base + (uint32_t)L.col[par][0], c1 = tmem_base + (uint32_t)L.col[par][1];
      const uint32_t x0c = tmem_base + (uint32_t)Lp.col[par][0], x1c = tmem_base + (uint32_t)Lp.col[par][1];
      const int hz = L.hz, hzd = L.hzd;
      if (hz >= 0 && t - hzd >= 0) mbar_wait_cluster(epidone(hz), (uint32_t)(t - hzd) & 1u);  // accumulator columns drained
      // K blocks whose A operand comes from shared memory (second activation tensor): independent of the previous
      // epilogue, so they go first and keep the tensor pipe busy while that epilogue runs
      const int ssk = L.ssk;
      for (int kb = 0; kb < ssk; ++kb) {
        ring_wait();
        const uint32_t ua = mu;
        const uint32_t a_addr = m_base + mu * slot_bytes;
        const uint64_t adesc = desc_hi | (uint64_t)((a_addr >> 4) & 0x3FFFu);
        ring_probe_next();
        ring_advance();
        ring_wait();
        const uint32_t b_addr = m_base + mu * slot_bytes;
        const uint64_t b0 = desc_hi | (uint64_t)((b_addr >> 4) & 0x3FFFu);
        const uint64_t b1 = desc_hi | (uint64_t)(((b_addr + kSubBytes) >> 4) & 0x3FFFu);
        ring_probe_next();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_pair(c0, adesc + 2u * k, b0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_pair(c1, adesc + 2u * k, b1 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_pair(mempty(ua));
          umma_commit_pair(mempty(mu));
        }
        ring_advance();
      }
      const uint32_t acc0 = ssk ? 1u : 0u;  // the tensor-memory K blocks accumulate on top of those
      const int slots = (kbn * nh) >> 1;
#pragma unroll 1
      for (int sl = 0; sl < slots; ++sl) {
        // the operand K block(s) of this slot are complete in tensor memory (both CTAs): per 64-channel slab of
        // the previous stage's epilogue, so these MMAs overlap the epilogue of its remaining slabs
        if (nh == 2) {
          mbar_wait_cluster(epislab(i - 1, sl), par);
        } else {
          mbar_wait_cluster(epislab(i - 1, 2 * sl), par);
          mbar_wait_cluster(epislab(i - 1, 2 * sl + 1), par);
        }
        tc_fence_after();
        if (kProbe && sl == 0 && dbg && di < 4000) dbg[di++] = clock64();  // [stage i operand ready]
        ring_wait();
        const uint32_t b_addr = m_base + mu * slot_bytes;
        const uint64_t b0 = desc_hi | (uint64_t)((b_addr >> 4) & 0x3FFFu);
        const uint64_t b1 = desc_hi | (uint64_t)(((b_addr + kSubBytes) >> 4) & 0x3FFFu);
        ring_probe_next();
        if (elect_one()) {
          if (nh == 2) {  // K block sl, both output halves
            const uint32_t xa = ((sl >> 1) ? x1c : x0c) + (sl & 1) * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(c0, xa + k * 8, b0 + 2u * k, idesc, (sl | k) != 0 ? 1u : acc0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(c1, xa + k * 8, b1 + 2u * k, idesc, (sl | k) != 0 ? 1u : acc0);
          } else {        // K blocks 2 sl and 2 sl + 1 of a 128-wide stage
            const uint32_t xa = sl ? x1c : x0c;  // K blocks 0,1 live in the first operand half, 2,3 in the second
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(c0, xa + k * 8, b0 + 2u * k, idesc, (sl | k) != 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_ts_pair(c0, xa + 64 + k * 8, b1 + 2u * k, idesc, 1u);
          }
          umma_commit_pair(mempty(mu));
        }
        ring_advance();
      }
      if (elect_one()) umma_commit_pair(accfull(i));  // accumulator ready; the consumed operand may be overwritten
      if (kProbe && dbg && di < 4000) dbg[di++] = clock64();  // [stage i issued]
      if (kProbe && (p.dbg_exec & 1)) {  // probe only: how long until the accumulator is complete (serialises this warp)
        mbar_wait_cluster(accfull(i), par);
        if (kProbe && dbg && di < 4000) dbg[di++] = clock64();
      }
    };

    for (int t = -1, pt = pt0 - pstride;; ++t, pt += pstride) {
      const bool has_next = pt + pstride < n_pt;
#pragma unroll 1
      for (int sidx = 0; sidx < n_chain; ++sidx) {
        const bool head_step = (n_chain == 1) || (sidx == head_after);
        if (head_step) {
          if (has_next) issue_head(t + 1);
        } else if (t >= 0) {
          issue_stage(sidx < head_after ? sidx + 1 : sidx, t);
        }
      }
      if (!has_next) break;
    }
  } else if (warp >= kEpiWarp0c) {
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int grp = (warp - kEpiWarp0c) >> 2; // slab parity this group handles
    // two pixel rows per lane (see epi_slab): lanes l and l + 16 share rows r16 and r16 + 16 of the quarter
    const uint32_t h = (uint32_t)lane >> 4;   // which 32 of a slab's 64 channels this lane owns
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint32_t row_off[2], sw[2], row2_off[2], sw2[2];
    int pw[2], phh[2], pn[2];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int m = q * 32 + 16 * w + (lane & 15);  // row of the tile = pixel
      row_off[w] = (uint32_t)m * 128u;
      sw[w] = (uint32_t)(m & 7);
      // position of the pixel inside the tile, and the row of its parent in the half-resolution residual slab
      // (box tw/2 x th/2 x nb)
      pw[w] = m % p.tw;
      phh[w] = (m / p.tw) % p.th;
      pn[w] = m / (p.tw * p.th);
      const int r2 = (pn[w] * (p.th >> 1) + (phh[w] >> 1)) * (p.tw >> 1) + (pw[w] >> 1);
      row2_off[w] = (uint32_t)r2 * 128u;
      sw2[w] = (uint32_t)(r2 & 7);
    }
    // this warp's quarter of the tile (32 consecutive pixels) as a TMA sub-box: offsets inside the tile
    const int rpq = 32 / p.tw;  // rows per quarter
    const int qx = 0, qy = (q * rpq) % p.th, qn = (q * rpq) / p.th;
    constexpr uint32_t kNoSlab = 0xffffffffu;
    uint32_t pending = kNoSlab;     // lane 0: slab whose TMA store may still be reading it
    uint32_t spos = 0, sphase = 0;  // ring position / phase of the next residual slab (slab 0 of the next stage with one)
    uint32_t n_slabs = (uint32_t)p.n_slabs, slab_bytes_r = (uint32_t)p.slab_bytes;
    int tile_th = p.th, img_B = p.B, img_H = p.H, img_W = p.W;
    asm volatile("" : "+r"(n_slabs), "+r"(slab_bytes_r), "+r"(tile_th), "+r"(img_B), "+r"(img_H), "+r"(img_W));
    unsigned long long* const dbg = (kProbe && blockIdx.x == 0 && lane == 0 && q == 0) ? p.dbg : nullptr;
    int di = 4096 * (1 + grp);
    const int dend = di + 4000;
    // probe, dbg_exec & 4: five stamps per slab of group 0 in a fourth region (loop top, slab ready, math done, operand
    // handed over, store issued)
    unsigned long long* const dbs = (kProbe && dbg && grp == 0 && (p.dbg_exec & 4)) ? p.dbg + 3 * 4096 : nullptr;
    int ds = 0;
    auto stamp = [&]() {
      if (kProbe && dbs && ds < 4000) dbs[ds++] = clock64();
    };

    // epilogue of stage i of the tile (local number t) at pixel offsets x0, y0, n0
    // pixel index of this lane's two rows for a tile at (x0, y0, n0), and whether they lie inside the batch
    // (partially filled multi-image tiles, phantom tiles): once per tile, not per stage
    struct TilePix {
      uint32_t pix[2];
      bool inb[2];
    };
    auto tile_pix = [&](int x0, int y0, int n0) {
      TilePix tp;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        tp.inb[w] = (n0 + pn[w]) < img_B && !(kProbe && (p.dbg_exec & 2));
        tp.pix[w] = (uint32_t)(((n0 + pn[w]) * img_H + (y0 + phh[w])) * img_W + (x0 + pw[w]));
      }
      return tp;
    };
    auto run_stage = [&](int i, int t, int x0, int y0, int n0, const TilePix& tp) {
      const StageLite& st = lite(i);
      const bool has_res = st.has_res != 0;
      const bool has_slab = has_res || st.stg != 0;  // works on a slab of the ring (residual, or blank for a staged output)
      const int x_src = st.x_src, kind = st.kind;
      const int nsl = st.n >> 6;
      uint8_t* out_row[2];
#pragma unroll
      for (int w = 0; w < 2; ++w)
        out_row[w] = reinterpret_cast<uint8_t*>(st.out) + (size_t)tp.pix[w] * (size_t)(st.n * 2);  // only used when st.out != 0
      const float4* const sc1 = reinterpret_cast<const float4*>(sm + (aff_base - smem_base)) + (st.aff_off >> 2);
      const float4* const sh1 = sc1 + (st.unit ? 0 : (st.n >> 2));
      const float4* const sc2 = sh1 + (st.n >> 2);
      const float4* const sh2 = sc2 + (st.n >> 2);
      if (kProbe && dbg && di < dend) dbg[di++] = clock64();  // [stage i entered]
      if (lane == 0) mbar_wait_cluster(accfull(i), (uint32_t)t & 1u);  // one polling lane (see mbar_wait_warp)
      __syncwarp();
      tc_fence_after();
      if (kProbe && dbg && di < dend) dbg[di++] = clock64();  // [stage i accumulator ready]
      const uint32_t t_lo = tmem_base + lane_base + (uint32_t)st.col[t & 1][0], t_hi = tmem_base + lane_base + (uint32_t)st.col[t & 1][1];
#pragma unroll 1
      for (int sl = grp; sl < nsl; sl += 2) {
        uint32_t su = 0, slab = s_base;
        stamp();
        const uint32_t t_slab = ((sl >> 1) ? t_hi : t_lo) + (sl & 1) * 64;
        if (has_slab) {
          su = spos + (uint32_t)sl;
          uint32_t sph = sphase;
          while (su >= n_slabs) {
            su -= n_slabs;
            sph ^= 1u;
          }
          mbar_wait_warp(sfull(su), sph);
          slab = s_base + su * slab_bytes_r;
        }
        stamp();
        EpiRow row[2];
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          row[w].rrow_s = slab + row_off[w];
          row[w].rrow2_s = slab + kUnitBytes + row2_off[w];
          row[w].sw = sw[w];
          row[w].sw2 = sw2[w];
          row[w].out = out_row[w] + sl * 128;
          row[w].store = tp.inb[w];
        }
        // constants of this lane's 32 channels of the slab (float4 units)
        const float4 *c1 = sc1 + sl * 16 + h * 8, *h1 = sh1 + sl * 16 + h * 8, *c2 = sc2 + sl * 16 + h * 8,
                     *h2 = sh2 + sl * 16 + h * 8;
        // the operand of the next stage replaces the first 32 of the 64 columns just read (in place)
        // (a tree of two-way branches on the stage's properties: a switch over `kind` compiles to a jump table in
        // the constant bank, whose load + indirect branch cost a few hundred cycles per slab when it misses)
        //                          UNIT   RES    RES2   RELU  XSRC OUT
        const bool up = (kind == kEpiResUpOutAct) | (kind == kEpiResUpOut);
        if (has_res) {
          if (x_src == 2) {
            if (up) epi_slab<true, true, true, false, 2, true>(t_slab, c1, h1, c2, h2, row, t_slab, h);    // kEpiResUpOutAct
            else    epi_slab<true, true, false, false, 2, true>(t_slab, c1, h1, c2, h2, row, t_slab, h);   // kEpiResOutAct
          } else if (x_src == 1) {
            epi_slab<true, true, false, false, 1, false>(t_slab, c1, h1, c2, h2, row, t_slab, h);          // kEpiResX
          } else {
            if (up) epi_slab<true, true, true, false, 0, true>(t_slab, c1, h1, c2, h2, row, t_slab, h);    // kEpiResUpOut
            else    epi_slab<true, true, false, false, 0, true>(t_slab, c1, h1, c2, h2, row, t_slab, h);   // kEpiResOut
          }
        } else {
          if (x_src)       epi_slab<false, false, false, true, 1, false>(t_slab, c1, h1, c2, h2, row, t_slab, h);  // kEpiReluX
          else if (has_slab) epi_slab<false, false, false, true, 0, true, false, true>(t_slab, c1, h1, c2, h2, row, t_slab, h);  // kEpiReluOut, staged
          else             epi_slab<false, false, false, true, 0, true>(t_slab, c1, h1, c2, h2, row, t_slab, h);   // kEpiReluOut
        }
        stamp();
        if (x_src) {  // this slab = one K block of the next stage's operand: hand it to the MMA warp first (the stores
                      // and the pooling below are off the chain's critical path)
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(epislab_l(i, sl));
        }
        stamp();
        if (has_slab && st.out) {
          // in-place output: this warp's 32 rows of the slab go out as one TMA store; the slab is handed back
          // to the producer once the store has read it -- one slab later, so that nobody waits for that
          fence_proxy_async();
          __syncwarp();
          if (st.pool_raw) {
            // 2x2 max-pool of the warp's own quarter (4 rows of the 8-wide tile = 2 x 4 pooling windows) straight
            // from the slab: 8 pooled pixels x 8 sixteen-byte chunks, two items per lane, stored directly (the 8
            // chunks of a pooled pixel are 128 contiguous bytes)
            const uint32_t R0 = 4u * (uint32_t)q;            // first tile row of the quarter
            const int rows_img = tile_th;                      // tile rows per image
            const int qn = n0 + (int)R0 / rows_img, qy = y0 + (int)R0 % rows_img;
            const float* const pc = reinterpret_cast<const float*>(sm + (aff_base - smem_base)) + st.pool_off;
#pragma unroll
            for (int it2 = 0; it2 < 2; ++it2) {
              const uint32_t item = (uint32_t)lane + 32u * it2;
              const uint32_t ppx = item >> 3, chunk = item & 7u;
              const uint32_t m0 = 32u * (uint32_t)q + (ppx >> 2) * 16u + 2u * (ppx & 3u);
              auto ldc = [&](uint32_t mm) { return lds128(slab + mm * 128u + ((chunk ^ (mm & 7u)) << 4)); };
              const uint4 a = ldc(m0), b = ldc(m0 + 1u), c = ldc(m0 + 8u), d = ldc(m0 + 9u);
              auto mx2 = [](uint32_t x, uint32_t y) {
                uint32_t r;
                asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
                return r;
              };
              uint4 mx;
              mx.x = mx2(mx2(a.x, b.x), mx2(c.x, d.x));
              mx.y = mx2(mx2(a.y, b.y), mx2(c.y, d.y));
              mx.z = mx2(mx2(a.z, b.z), mx2(c.z, d.z));
              mx.w = mx2(mx2(a.w, b.w), mx2(c.w, d.w));
              const int cc = sl * 64 + (int)chunk * 8;
              const float4 sa = *reinterpret_cast<const float4*>(pc + cc), sb = *reinterpret_cast<const float4*>(pc + cc + 4);
              const float4 ha = *reinterpret_cast<const float4*>(pc + st.n + cc), hb = *reinterpret_cast<const float4*>(pc + st.n + cc + 4);
              uint4 act;
              {
                float2 v;
                v = ffma2(bf2_unpack(mx.x), make_float2(sa.x, sa.y), make_float2(ha.x, ha.y));
                act.x = pack2_relu(v.x, v.y);
                v = ffma2(bf2_unpack(mx.y), make_float2(sa.z, sa.w), make_float2(ha.z, ha.w));
                act.y = pack2_relu(v.x, v.y);
                v = ffma2(bf2_unpack(mx.z), make_float2(sb.x, sb.y), make_float2(hb.x, hb.y));
                act.z = pack2_relu(v.x, v.y);
                v = ffma2(bf2_unpack(mx.w), make_float2(sb.z, sb.w), make_float2(hb.z, hb.w));
                act.w = pack2_relu(v.x, v.y);
              }
              if (qn < img_B && !(kProbe && (p.dbg_exec & 2))) {
                const int py = (qy >> 1) + (int)(ppx >> 2), px = (x0 >> 1) + (int)(ppx & 3u);
                const size_t off = ((((size_t)qn * (img_H >> 1) + py) * (img_W >> 1) + px) * (size_t)st.n + (size_t)cc) * 2;
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(st.pool_raw) + off) = mx;
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(st.pool_act) + off) = act;
              }
            }
          }
          if (lane == 0) {
            if (!(kProbe && (p.dbg_exec & 2))) tma_store_4d(&p.st[i].tmOutQ, slab + (uint32_t)q * 4096u, sl * 64, x0 + qx, y0 + qy, n0 + qn);
            tma_store_commit();
            if (pending != kNoSlab) {
              tma_store_wait_read<1>();
              mbar_arrive(sempty(pending));
            }
            pending = su;
          }
        } else if (has_slab) {  // slab consumed by this warp
          __syncwarp();
          if (lane == 0) mbar_arrive(sempty(su));
        }
        stamp();
      }
      if (has_slab) {
        spos += (uint32_t)nsl;
        while (spos >= n_slabs) {
          spos -= n_slabs;
          sphase ^= 1u;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(epidone_l(i));  // counted on the leader's barrier (accumulator columns drained)
      if (lane == 0 && pending != kNoSlab) {  // after the hand-over to the tensor pipe: release the last stored slab
        tma_store_wait_read<0>();
        mbar_arrive(sempty(pending));
        pending = kNoSlab;
      }
      if (kProbe && dbg && di < dend) dbg[di++] = clock64();  // [stage i epilogue done]
    };

    int x0 = 0, y0 = 0, n0 = 0;
    TilePix tp_cur = tile_pix(0, 0, 0), tp_next = tp_cur;
    for (int t = -1, pt = pt0 - pstride;; ++t, pt += pstride) {
      const bool has_next = pt + pstride < n_pt;
      int nx0 = 0, ny0 = 0, nn0 = 0;
      if (has_next) {
        decode_tile(2 * (pt + pstride) + (int)rank, nx0, ny0, nn0);
        tp_next = tile_pix(nx0, ny0, nn0);
      }
#pragma unroll 1
      for (int sidx = 0; sidx < n_chain; ++sidx) {
        const bool head_step = (n_chain == 1) || (sidx == head_after);
        if (head_step ? !has_next : (t < 0)) continue;
        const int i = head_step ? 0 : (sidx < head_after ? sidx + 1 : sidx);
        run_stage(i, head_step ? t + 1 : t, head_step ? nx0 : x0, head_step ? ny0 : y0, head_step ? nn0 : n0,
                  head_step ? tp_next : tp_cur);
      }
      if (!has_next) break;
      x0 = nx0;
      y0 = ny0;
      n0 = nn0;
      tp_cur = tp_next;
    }
    if (lane == 0) tma_store_wait_all();  // every output tile has left shared memory and is written
  }

  // neither CTA may leave (or free its tensor memory) while the other can still signal its barriers
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------ host
int make_tmap_quarter(CUtensorMap* out, const void* base, int C, int W, int H, int N, int tw, int th, int nb) {
  DF3D_REQUIRE(tw > 0 && 32 % tw == 0 && tw * th * nb == 128, DF3D_EINVAL, "make_tmap_quarter: bad tile %d x %d x %d", tw, th, nb);
  const int rows = 32 / tw;
  if (rows <= th) {
    DF3D_REQUIRE(th % rows == 0, DF3D_EINVAL, "make_tmap_quarter: bad tile %d x %d x %d", tw, th, nb);
    return make_tmap_box(out, base, C, W, H, N, tw, rows, 1);
  }
  DF3D_REQUIRE(rows % th == 0, DF3D_EINVAL, "make_tmap_quarter: bad tile %d x %d x %d", tw, th, nb);
  return make_tmap_box(out, base, C, W, H, N, tw, th, rows / th);
}

int conv_chain_configure() {
  DF3D_CUDA(cudaFuncSetAttribute(conv_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemLimit));
  DF3D_CUDA(cudaFuncSetAttribute(conv_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kChainSmemLimit));
  return DF3D_OK;
}

// Tensor-memory plan of a chain: accumulator columns per stage, where in the issue order the next
// tile's head GEMM goes, and the one epilogue each stage has to wait for before it may overwrite its
// accumulator columns (the operand dependency on the previous stage is implicit).  Rules:
//   * the operand of stage i+1 lives in place in the accumulator columns of stage i, so those columns
//     stay busy until stage i+1 has been issued (the tensor pipe executes in issue order);
//   * a hazard (stage j, delta) means "epilogue j of tile t - delta"; with delta = 1, j >= i keeps the
//     parity wait unambiguous (epilogue j of tile t cannot have completed yet).
static int plan_tmem(ChainParams& p) {
  const int n = p.n_chain;
  auto set = [&](int i, int lo, int hi, int hz, int delta) {
    p.st[i].col[0][0] = p.st[i].col[1][0] = lo;
    p.st[i].col[0][1] = p.st[i].col[1][1] = hi;
    p.st[i].hz_stage = hz;
    p.st[i].hz_delta = delta;
  };
  auto set2 = [&](int i, int lo_e, int hi_e, int lo_o, int hi_o, int hz, int delta) {  // even / odd tiles differ
    p.st[i].col[0][0] = lo_e;
    p.st[i].col[0][1] = hi_e;
    p.st[i].col[1][0] = lo_o;
    p.st[i].col[1][1] = hi_o;
    p.st[i].hz_stage = hz;
    p.st[i].hz_delta = delta;
  };
  int w[kMaxChain] = {0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) w[i] = p.st[i].n;
  auto is = [&](std::initializer_list<int> pat) {
    if ((int)pat.size() != n) return false;
    int i = 0;
    for (int v : pat)
      if (w[i++] != v) return false;
    return true;
  };
  if (is({128, 256, 128})) {  // bottleneck tail + next conv1: P | Q | R, next head behind conv3
    set(0, kColP, kColP, -1, 0);
    set(1, kColQ, kColQ + 128, -1, 0);
    set(2, kColR, kColR, 2, 1);
    p.head_after = 1;
  } else if (is({128, 256})) {  // bottleneck tail alone
    set(0, kColP, kColP, -1, 0);
    set(1, kColQ, kColQ + 128, 1, 1);
    p.head_after = 1;
  } else if (is({128, 256, 256, 256, 128})) {  // inter-stack chain: P | Q | P+R | Q | R, next head behind the merged conv
    set(0, kColP, kColP, -1, 0);
    set(1, kColQ, kColQ + 128, -1, 0);
    set(2, kColP, kColR, 4, 1);
    set(3, kColQ, kColQ + 128, -1, 0);
    set(4, kColR, kColR, -1, 0);
    p.head_after = 3;
  } else if (is({128, 256, 256, 128})) {
    // inter-stack chain with res.conv3 and fc merged: four 128-column blocks P, Q0, Q1, R whose roles rotate with
    // the tile parity so that the next head can go behind stage 2 (the heavy merged-skip epilogue):
    //   even tile: head P | fc Q0+Q1 | skip R+P | conv1 Q0 | next head Q1
    //   odd tile:  head Q1 | fc R+P  | skip Q0+Q1 | conv1 R | next head P
    // hazards: fc overwrites the previous tile's skip accumulator (its epilogue 2), skip the previous conv1's (3)
    const int Q0 = kColQ, Q1 = kColQ + 128;
    set2(0, kColP, kColP, Q1, Q1, -1, 0);
    set2(1, Q0, Q1, kColR, kColP, 2, 1);
    set2(2, kColR, kColP, Q0, Q1, 3, 1);
    set2(3, Q0, Q0, kColR, kColR, -1, 0);
    p.head_after = 2;
  } else if (is({128, 256, 256})) {  // last stack: P | Q | P+R; the next head has to wait for the fc epilogue
    set(0, kColP, kColP, 2, 1);
    set(1, kColQ, kColQ + 128, -1, 0);
    set(2, kColP, kColR, -1, 0);
    p.head_after = 2;
  } else if (is({256, 128})) {  // point-wise chains behind a stand-alone 3x3 conv
    set(0, kColQ, kColQ + 128, -1, 0);
    set(1, kColP, kColP, 1, 1);
    p.head_after = 1;
  } else if (is({256, 256, 256, 128})) {
    set(0, kColQ, kColQ + 128, -1, 0);
    set(1, kColP, kColR, 3, 1);
    set(2, kColQ, kColQ + 128, -1, 0);
    set(3, kColP, kColP, -1, 0);
    p.head_after = 3;
  } else if (is({256, 256})) {
    set(0, kColQ, kColQ + 128, -1, 0);
    set(1, kColP, kColR, 1, 1);
    p.head_after = 1;
  } else if (is({128}) || is({256})) {
    set(0, w[0] == 128 ? kColP : kColQ, w[0] == 128 ? kColP : kColQ + 128, 0, 1);
    p.head_after = 0;
  } else {
    DF3D_REQUIRE(false, DF3D_EUNSUPPORTED, "launch_conv_chain: no tensor-memory plan for this chain shape");
  }
  for (int i = 0; i < n; ++i)
    DF3D_REQUIRE(p.st[i].hz_stage < 0 || p.st[i].hz_delta == 0 || p.st[i].hz_stage >= i, DF3D_EINVAL,
                 "launch_conv_chain: ambiguous hazard in the tensor-memory plan");
  return DF3D_OK;
}

int launch_conv_chain(const ChainParams& p_in, int num_sms, cudaStream_t stream) {
  ChainParams p = p_in;
  const int total = p.tiles_x * p.tiles_y * p.tiles_b;
  if (total <= 0) return DF3D_OK;
  DF3D_REQUIRE(p.n_chain >= 1 && p.n_chain <= kMaxChain, DF3D_EINVAL, "launch_conv_chain: bad chain length %d", p.n_chain);
  DF3D_REQUIRE(p.tw * p.th * p.nb == 128, DF3D_EINVAL, "launch_conv_chain: tile must hold 128 pixels");
  int aff_floats = 0;
  bool any_slab = false, any_res2 = false;
  for (int i = 0; i < p.n_chain; ++i) {
    ChainStage& st = p.st[i];
    DF3D_REQUIRE(st.n == 128 || st.n == 256, DF3D_EUNSUPPORTED, "launch_conv_chain: stage %d has %d output channels (128 or 256)", i, st.n);
    DF3D_REQUIRE(i == 0 || st.kblocks * 64 == p.st[i - 1].n, DF3D_EINVAL, "launch_conv_chain: stage %d K does not match stage %d N", i, i - 1);
    DF3D_REQUIRE(st.ss_kblocks == 0 || (i > 0 && st.n == 256 && st.ss_kblocks <= 4), DF3D_EUNSUPPORTED,
                 "launch_conv_chain: shared-memory K blocks need a later stage with 256 outputs");
    DF3D_REQUIRE((i + 1 < p.n_chain) == (st.x_src != 0), DF3D_EINVAL, "launch_conv_chain: x_src must be set on every stage but the last");
    DF3D_REQUIRE(!st.has_res2 || (st.has_res && p.tw % 2 == 0 && p.th % 2 == 0), DF3D_EUNSUPPORTED,
                 "launch_conv_chain: the half-resolution residual needs a full-resolution residual and an even tile");
    DF3D_REQUIRE(st.x_src || st.out_raw, DF3D_EINVAL, "launch_conv_chain: the last stage must store its output");
    {
      const bool relu = st.relu1 != 0, res = st.has_res != 0, up = st.has_res2 != 0, out = st.out_raw != nullptr;
      const bool unit = st.unit_scale != 0;
      int kind = -1;
      if (!unit && !res && !up && relu && st.x_src == 1 && !out) kind = kEpiReluX;
      if (!unit && !res && !up && relu && st.x_src == 0 && out) kind = kEpiReluOut;
      if (unit && res && !up && !relu && st.x_src == 2 && out) kind = kEpiResOutAct;
      if (unit && res && up && !relu && st.x_src == 2 && out) kind = kEpiResUpOutAct;
      if (unit && res && !up && !relu && st.x_src == 0 && out) kind = kEpiResOut;
      if (unit && res && up && !relu && st.x_src == 0 && out) kind = kEpiResUpOut;
      if (unit && res && !up && !relu && st.x_src == 1 && !out) kind = kEpiResX;
      DF3D_REQUIRE(kind >= 0, DF3D_EUNSUPPORTED,
                   "launch_conv_chain: stage %d has no specialised epilogue (unit %d res %d up %d relu %d x %d out %d)", i,
                   (int)unit, (int)res, (int)up, (int)relu, st.x_src, (int)out);
      st.epi_kind = kind;
    }
    st.aff_off = aff_floats;
    aff_floats += st.n * ((st.unit_scale ? 1 : 2) + (st.x_src == 2 ? 2 : 0));
    if (st.pool_raw) {
      DF3D_REQUIRE(st.has_res && st.out_raw && st.pool_act && st.pool_scale && st.pool_shift && p.tw == 8 && p.th % 4 == 0 &&
                       p.H % 2 == 0 && p.W % 2 == 0,
                   DF3D_EUNSUPPORTED, "launch_conv_chain: the pooled output needs a stored stage with a residual on 8-wide tiles");
      st.pool_off = aff_floats;
      aff_floats += 2 * st.n;
    }
    // stored stages without a residual leave through a blank slab + TMA store when the caller provided the store map
    st.stage_out = (st.epi_kind == kEpiReluOut && p.tw * p.th * p.nb == 128 && !getenv("DF3D_CHAIN_NO_STAGE_OUT")) ? 1 : 0;
    any_slab |= st.has_res != 0 || st.stage_out != 0;
    any_res2 |= st.has_res2 != 0;
  }
  if (int e = plan_tmem(p)) return e;
  // the slab producer walks tile by tile, stage by stage; the epilogue pulls the next head forward, which
  // only keeps the same order when the head has no residual or is issued behind the last stage anyway
  DF3D_REQUIRE(!p.st[0].has_res || p.head_after == p.n_chain - 1, DF3D_EUNSUPPORTED,
               "launch_conv_chain: a head with a residual must be issued behind the last stage");
  // shared-memory budget: constants, barriers, residual slabs, the rest is the operand ring
  p.aff_bytes = (aff_floats * 4 + 255) & ~255;
  p.slab_bytes = kUnitBytes + (any_res2 ? kUnitBytes / 4 : 0);
  // ring slot (per CTA): two 64-row weight sub-tiles; without the halo a head K block = A tile + its sub-tiles
  p.slot_bytes = 2 * kSubBytes;
  if (!p.halo) p.slot_bytes = kUnitBytes + (p.st[0].n >> 7) * kSubBytes;
  if (p.halo) {
    DF3D_REQUIRE(p.taps == 9 && p.kc_per_tap == 2 && p.st[0].n == 128 && p.tw == 8 && p.th == 16 && p.nb == 1, DF3D_EINVAL,
                 "launch_conv_chain: halo mode needs a 3x3 head with 128 input and output channels on 8 x 16 tiles");
  }
  p.n_slabs = any_slab ? 4 : 0;
  if (const char* env = getenv("DF3D_CHAIN_NS")) {  // profiling knob: residual slab depth
    const int v = atoi(env);
    if (any_slab && v >= 2 && v <= kMaxSlabs) p.n_slabs = v;
  }
  int fixed = 0;
  for (;;) {
    fixed = 1024 + p.aff_bytes + kChainBarBytes + p.n_slabs * p.slab_bytes + (p.halo ? 2 * kHaloBytes : 0);
    p.n_m = (kChainSmemLimit - fixed) / p.slot_bytes;
    if (p.n_m >= 3 || p.n_slabs <= 2) break;
    --p.n_slabs;  // wide heads: trade slab depth for ring depth
  }
  if (p.n_m > kMaxM) p.n_m = kMaxM;
  if (const char* env = getenv("DF3D_CHAIN_NM")) {  // profiling knob: ring depth
    const int v = atoi(env);
    if (v >= 2 && v < p.n_m) p.n_m = v;
  }
  DF3D_REQUIRE(p.n_m >= 2, DF3D_EUNSUPPORTED, "launch_conv_chain: shared-memory budget too small (%d ring slots)", p.n_m);
  const int smem = fixed + p.n_m * p.slot_bytes;
  // tile decode without divisions when the tile grid is a power of two (it is for every hourglass level)
  auto log2_exact = [](int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
  };
  p.tx_shift = log2_exact(p.tiles_x);
  p.ty_shift = log2_exact(p.tiles_y);
  // clusters of two CTAs (one TPC): an even grid, one pair per two tiles
  const int pairs = (total + 1) / 2;
  const int grid = 2 * (pairs < num_sms / 2 ? pairs : num_sms / 2);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kChainThreads);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (p.dbg) {
    DF3D_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel<true>, p));
  } else {
    DF3D_CUDA(cudaLaunchKernelEx(&cfg, conv_chain_kernel<false>, p));
  }
  DF3D_LAUNCH_CHECK("conv_chain_kernel");
  return DF3D_OK;
}

}  // namespace df3d
