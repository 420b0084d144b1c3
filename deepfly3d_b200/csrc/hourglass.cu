// Stacked-hourglass forward as a static plan of sm_100a kernels (host-side executor + C ABI).
// Replaces the network forward + decode inside df2d.inference.inference_folder (reference call
// site df3d/core.py:177-185).
//
// df3d_hg_create parses the float32 parameter blob, folds every BatchNorm into a per-channel
// (scale, shift) pair, packs every conv weight to bf16 [CoutPad][taps*CinPad] (K-major) and
// uploads both once.  The forward is a fixed list of launches over a chunk of images:
//   stem_im2col -> conv_gemm / conv_chain (tcgen05) x N, maxpool_bn_relu -> score head with the arg-max in
//   its epilogue -> key decode.
// Each conv's epilogue applies the *next* BatchNorm + ReLU (pre-activation bottlenecks), the
// residual add, the hourglass' nearest x2 up-sample + add and the bf16 rounding.  Unfused (DF3D_HG_FUSE=0) a
// bottleneck is three (four with a projection) GEMM launches and no elementwise pass; in the default plan
// (DF3D_HG_FUSE=2) one conv_chain launch runs [3x3 -> conv3 (+ residual, + up-sample add) -> next block's
// conv1], and the inter-stack chain five convs, with the intermediates kept in tensor memory.  All
// activations live in the caller's workspace; a small free-list arena reuses buffers so the working set
// stays small.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "conv_chain.cuh"
#include "conv_gemm.cuh"
#include "hg_elementwise.cuh"

namespace df3d {

constexpr float kBnEps = 1e-5f;
constexpr int kDepth = 4;
constexpr int kFeats = 128;        // bottleneck planes inside the hourglass
constexpr int kCh = 2 * kFeats;    // 256 channels on the residual stream
constexpr int kInplanes = 64;
constexpr int kHeatPad = 32;       // fp32 score channels stored per pixel (>= num_classes)

static inline uint16_t f2bf(float f) {  // round-to-nearest-even, like __float2bfloat16_rn
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

// ------------------------------------------------------------------------------ parameter views
struct BNp {
  const float *g = nullptr, *b = nullptr, *m = nullptr, *v = nullptr;
  int n = 0;
};
struct Convp {
  const float *w = nullptr, *bias = nullptr;
  int cout = 0, cin = 0, k = 0;
};
struct Bott {
  BNp bn1, bn2, bn3;
  Convp c1, c2, c3, ds;
  bool has_ds = false;
  int inpl = 0, planes = 0;
};
struct StackP {
  Bott hg[kDepth][4];  // hg[d][0..2], hg[0][3]
  Bott res;
  Convp fc;
  BNp fc_bn;
  Convp score, fc_, score_;
};
struct NetP {
  Convp conv1;
  BNp bn1;
  Bott layer1, layer2, layer3;
  std::vector<StackP> stacks;
};

struct Cursor {
  const float* p;
  size_t n, pos = 0;
  bool ok = true;
  const float* take(size_t k) {
    if (pos + k > n) {
      ok = false;
      return p;  // caller checks ok
    }
    const float* r = p + pos;
    pos += k;
    return r;
  }
};

static BNp read_bn(Cursor& c, int n) {
  BNp b;
  b.n = n;
  b.g = c.take(n);
  b.b = c.take(n);
  b.m = c.take(n);
  b.v = c.take(n);
  return b;
}
static Convp read_conv(Cursor& c, int cout, int cin, int k) {
  Convp v;
  v.cout = cout;
  v.cin = cin;
  v.k = k;
  v.w = c.take((size_t)cout * cin * k * k);
  v.bias = c.take(cout);
  return v;
}
static Bott read_bott(Cursor& c, int inpl, int planes) {
  Bott b;
  b.inpl = inpl;
  b.planes = planes;
  b.bn1 = read_bn(c, inpl);
  b.c1 = read_conv(c, planes, inpl, 1);
  b.bn2 = read_bn(c, planes);
  b.c2 = read_conv(c, planes, planes, 3);
  b.bn3 = read_bn(c, planes);
  b.c3 = read_conv(c, 2 * planes, planes, 1);
  b.has_ds = inpl != 2 * planes;
  if (b.has_ds) b.ds = read_conv(c, 2 * planes, inpl, 1);
  return b;
}
static void read_net(Cursor& c, int num_stacks, int K, NetP* net) {
  net->conv1 = read_conv(c, kInplanes, 3, 7);
  net->bn1 = read_bn(c, kInplanes);
  net->layer1 = read_bott(c, kInplanes, kInplanes);
  net->layer2 = read_bott(c, 2 * kInplanes, kInplanes);
  net->layer3 = read_bott(c, 2 * kInplanes, kFeats);
  net->stacks.resize(num_stacks);
  for (int i = 0; i < num_stacks; ++i) {
    StackP& s = net->stacks[i];
    for (int d = 0; d < kDepth; ++d)
      for (int k = 0; k < (d == 0 ? 4 : 3); ++k) s.hg[d][k] = read_bott(c, kCh, kFeats);
    s.res = read_bott(c, kCh, kFeats);
    s.fc = read_conv(c, kCh, kCh, 1);
    s.fc_bn = read_bn(c, kCh);
    s.score = read_conv(c, K, kCh, 1);
    if (i < num_stacks - 1) {
      s.fc_ = read_conv(c, kCh, kCh, 1);
      s.score_ = read_conv(c, kCh, K, 1);
    }
  }
}

// ------------------------------------------------------------------------------ plan
struct Tensor {
  size_t off = 0;  // byte offset in the workspace
  size_t bytes = 0;
  int H = 0, W = 0, C = 0;
  bool valid = false;
};

struct Arena {
  size_t top = 0;
  std::multimap<size_t, size_t> free_blocks;  // bytes -> offset
  size_t alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    auto it = free_blocks.find(bytes);
    if (it != free_blocks.end()) {
      size_t o = it->second;
      free_blocks.erase(it);
      return o;
    }
    size_t o = top;
    top += bytes;
    return o;
  }
  void release(size_t off, size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);
    free_blocks.emplace(bytes, off);
  }
};

enum OpKind { OP_IM2COL, OP_IM2COL_GRAY, OP_CONV, OP_CHAIN, OP_POOL, OP_ARGMAX };

struct Op {
  OpKind kind;
  int variant = 0;  // stem only: 0 always, 1 three-plane path, 2 gray fast path (uint8 input, one common mean)
  // conv / chain
  ConvParams conv;
  ChainParams chain;
  double chain_bytes_per_image = 0.0;  // algorithmic HBM bytes of a chain (head input + residuals + stored outputs)
  int BN = 0, nb = 1;
  // elementwise / argmax: byte offsets resolved to pointers at plan time
  char *in0 = nullptr, *in1 = nullptr, *out0 = nullptr, *out1 = nullptr;
  const float *scale = nullptr, *shift = nullptr;
  int H = 0, W = 0, C = 0;
  double flops_per_image = 0.0;  // algorithmic 2*MAC of a conv op (real Cin/Cout, no padding)
};

struct Affine {  // index into the float blob
  size_t scale_off, shift_off;
};

}  // namespace df3d

using namespace df3d;

struct df3d_hg {
  df3d_hg_desc desc;
  float mean[3] = {0.5f, 0.5f, 0.5f};
  int chunk = 0;     // images per lane and launch sequence
  int n_lanes = 1;   // independent image groups run concurrently on their own stream and SM share
  int fuse = 2;      // 0: one launch per conv, 1: point-wise chains behind stand-alone 3x3 convs, 2: 3x3-led chains
  int num_sms = 148;
  size_t lane_bytes = 0;
  cudaStream_t lane_stream[4] = {nullptr, nullptr, nullptr, nullptr};  // [0] unused (caller's stream)
  cudaEvent_t fork_ev = nullptr, join_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<float> params;  // host copy of the blob (views in `net` point into it)
  NetP net;
  // host-side packed data (filled during the sizing pass)
  std::vector<uint16_t> wblob;
  std::vector<float> ablob;
  std::vector<std::vector<float>> merged_w, merged_b;  // per-stack merged fc_/score_/score weights
  std::vector<std::vector<float>> fcm_w, fcm_b;        // per-stack merged fc . res.conv3 weights [256][128] and bias
  uint16_t* d_w = nullptr;
  float* d_a = nullptr;
  unsigned long long* d_keys = nullptr;  // [max_batch][kHeatPad] arg-max keys of the fused score head (zero between forwards)
  size_t ws_bytes = 0;
  int ops_per_chunk = 0;
  // plan for the workspace pointer it was built for
  char* plan_base = nullptr;
  std::vector<Op> ops;            // plan under construction / lane 0 (all lanes share its structure)
  std::vector<std::vector<Op>> lane_ops;
  char* heat_ptr = nullptr;  // fp32 score tensor of the last stack inside the workspace
  int score_nb = 0;          // images per M tile of the score head (the fused arg-max needs 1)
  int Hh = 0, Wh = 0;
  // optional per-launch timing (bench / profiling only): events[chunk][op][2]
  bool timing = false;
  std::vector<cudaEvent_t> events;
  std::vector<int> timed_bc;  // images of each timed chunk of the last forward
  int timed_stem_variant = 0;  // stem variant (1 three-plane, 2 gray) the last timed forward ran
};

namespace df3d {

int launch_argmax_keys_decode(unsigned long long* keys, int B, int Cpad, int K, int32_t* idx, float* conf, cudaStream_t s);  // argmax.cu

// One emission pass.  dry == true: pack weights / affines on the host and size the workspace.
// dry == false: same traversal (identical offsets), but produce launchable ops for `base`.
struct Emitter {
  df3d_hg* hg;
  bool dry;
  char* base;
  Arena arena;
  size_t w_cursor = 0, a_cursor = 0;
  int err = DF3D_OK;
  int B;  // chunk
  int n_ops = 0;  // launches per chunk (counted in both passes)

  Tensor talloc(int H, int W, int C, int elem = 2) {
    Tensor t;
    t.H = H;
    t.W = W;
    t.C = C;
    t.bytes = (size_t)B * H * W * C * elem;
    t.off = arena.alloc(t.bytes);
    t.valid = true;
    return t;
  }
  void tfree(Tensor& t) {
    if (t.valid) arena.release(t.off, t.bytes);
    t.valid = false;
  }
  char* ptr(const Tensor& t) const { return t.valid ? base + t.off : nullptr; }

  // reserve n floats in the affine blob; returns offset
  size_t areserve(size_t n) {
    size_t o = a_cursor;
    a_cursor += n;
    if (dry && hg->ablob.size() < a_cursor) hg->ablob.resize(a_cursor, 0.0f);
    return o;
  }
  // folded BN: y = x*scale + shift
  Affine bn_affine(const BNp& bn, int pad) {
    Affine a{areserve(pad), areserve(pad)};
    if (dry) {
      for (int i = 0; i < bn.n; ++i) {
        const float s = bn.g[i] / std::sqrt(bn.v[i] + kBnEps);
        hg->ablob[a.scale_off + i] = s;
        hg->ablob[a.shift_off + i] = bn.b[i] - bn.m[i] * s;
      }
    }
    return a;
  }
  // conv epilogue affine: v = acc*scale + shift where the conv has `bias` and is optionally
  // followed by BN `bn`:  (acc + bias)*s + t
  Affine conv_affine(const Convp& c, const BNp* bn, int pad) {
    Affine a{areserve(pad), areserve(pad)};
    if (dry) {
      for (int i = 0; i < pad; ++i) {
        float s = 1.0f, t = 0.0f, b = i < c.cout ? c.bias[i] : 0.0f;
        if (bn && i < bn->n) {
          s = bn->g[i] / std::sqrt(bn->v[i] + kBnEps);
          t = bn->b[i] - bn->m[i] * s;
        }
        if (i >= c.cout) s = 0.0f;
        hg->ablob[a.scale_off + i] = s;
        hg->ablob[a.shift_off + i] = b * s + t;
      }
    }
    return a;
  }
  // pack OIHW fp32 -> bf16 [CoutPad][taps*CinPad], K index = tap*CinPad + cin
  size_t pack_weights(const Convp& c, int CoutPad, int CinPad) {
    const int taps = c.k * c.k;
    const size_t K = (size_t)taps * CinPad;
    size_t o = w_cursor;
    w_cursor += (size_t)CoutPad * K;
    if (dry) {
      hg->wblob.resize(w_cursor, 0);
      for (int co = 0; co < c.cout; ++co)
        for (int ci = 0; ci < c.cin; ++ci)
          for (int t = 0; t < taps; ++t)
            hg->wblob[o + (size_t)co * K + (size_t)t * CinPad + ci] = f2bf(c.w[((size_t)co * c.cin + ci) * taps + t]);
    }
    return o;
  }
  // K-concatenation of two 1x1 convs with the same outputs: [CoutPad][a.cin + b.cin], a's inputs first
  size_t pack_weights_concat(const Convp& a, const Convp& b, int CoutPad) {
    const size_t K = (size_t)a.cin + b.cin;
    size_t o = w_cursor;
    w_cursor += (size_t)CoutPad * K;
    if (dry) {
      hg->wblob.resize(w_cursor, 0);
      for (int co = 0; co < a.cout; ++co) {
        for (int ci = 0; ci < a.cin; ++ci) hg->wblob[o + (size_t)co * K + ci] = f2bf(a.w[(size_t)co * a.cin + ci]);
        for (int ci = 0; ci < b.cin; ++ci) hg->wblob[o + (size_t)co * K + a.cin + ci] = f2bf(b.w[(size_t)co * b.cin + ci]);
      }
    }
    return o;
  }
  // epilogue affine of that sum: scale 1, shift = both biases
  Affine concat_affine(const Convp& a, const Convp& b, int pad) {
    Affine af{areserve(pad), areserve(pad)};
    if (dry)
      for (int i = 0; i < pad; ++i) {
        hg->ablob[af.scale_off + i] = i < a.cout ? 1.0f : 0.0f;
        hg->ablob[af.shift_off + i] = i < a.cout ? a.bias[i] + b.bias[i] : 0.0f;
      }
    return af;
  }

  // stem: K index = (ky*7+kx)*3 + c  (must match stem_im2col_kernel), padded to 192
  size_t pack_stem(const Convp& c) {
    const size_t K = kStemKPadCols;
    size_t o = w_cursor;
    w_cursor += (size_t)c.cout * K;
    if (dry) {
      hg->wblob.resize(w_cursor, 0);
      for (int co = 0; co < c.cout; ++co)
        for (int ci = 0; ci < 3; ++ci)
          for (int t = 0; t < 49; ++t)
            hg->wblob[o + (size_t)co * K + (size_t)t * 3 + ci] = f2bf(c.w[((size_t)co * 3 + ci) * 49 + t]);
    }
    return o;
  }

  // gray fast path: the three input planes are identical, so the weights are summed over the input
  // channel (fp32) before the bf16 rounding; K index = ky*7 + kx, padded to 64
  size_t pack_stem_gray(const Convp& c) {
    const size_t K = kStemKGray;
    size_t o = w_cursor;
    w_cursor += (size_t)c.cout * K;
    if (dry) {
      hg->wblob.resize(w_cursor, 0);
      for (int co = 0; co < c.cout; ++co)
        for (int t = 0; t < 49; ++t) {
          float sum = 0.0f;
          for (int ci = 0; ci < 3; ++ci) sum += c.w[((size_t)co * 3 + ci) * 49 + t];
          hg->wblob[o + (size_t)co * K + t] = f2bf(sum);
        }
    }
    return o;
  }

  static void tile_geometry(int H, int W, int* tw, int* th, int* nb) {
    int w = W < 16 ? W : 16;
    int h = 128 / w;
    if (h > H) h = H;
    *tw = w;
    *th = h;
    *nb = 128 / (w * h);
  }

  // Generic conv emission.  `in` is the A operand (already activated), weights at w_off with
  // K = taps*CinPad.  Outputs may be invalid tensors (skipped).
  void conv(const Tensor& in, size_t w_off, int taps, int CinPad, int CoutPad, int BN, Affine a1, bool relu1,
            const Tensor* residual, const Tensor* out_raw, const Affine* a2, const Tensor* out_act,
            const Tensor* out_f32, double flop_per_px, const Tensor* res2_half = nullptr, const Tensor* in2 = nullptr,
            const Tensor* pool_raw = nullptr, const Tensor* pool_act = nullptr) {
    ++n_ops;
    if (err) return;
    if (dry) return;
    Op op;
    op.kind = OP_CONV;
    op.BN = BN;
    op.flops_per_image = flop_per_px * in.H * in.W;
    ConvParams& p = op.conv;
    memset(&p, 0, sizeof(p));
    int tw, th, nb;
    tile_geometry(in.H, in.W, &tw, &th, &nb);
    // 3x3 convs with 64 channels in and out (layer1 / layer2 of the front section): halo mode of conv_gemm
    const bool halo9 = taps == 9 && CinPad == 64 && CoutPad == 64 && BN == 64 && in.H % 16 == 0 && in.W % 8 == 0 &&
                       !(residual && residual->valid) && !getenv("DF3D_HG_NO_HALO");
    if (halo9) {
      tw = 8;
      th = 16;
      nb = 1;
      p.halo9 = 1;
      if ((err = make_tmap_box(&p.tmHalo, ptr(in), in.C, in.W, in.H, B, tw + 2, th + 2, 1))) return;
    }
    op.nb = nb;
    if ((err = make_tmap_act(&p.tmA, ptr(in), in.C, in.W, in.H, B, tw, th, nb))) return;
    if ((err = make_tmap_wgt(&p.tmB, hg->d_w + w_off, taps * CinPad, CoutPad, BN))) return;
    p.taps = taps;
    p.kc_per_tap = CinPad / 64;
    if (in2 && in2->valid) {  // K-concatenation: the last in2->C / 64 K blocks come from the second tensor
      p.kb_split = (CinPad - in2->C) / 64;
      if ((err = make_tmap_act(&p.tmA2, ptr(*in2), in2->C, in.W, in.H, B, tw, th, nb))) return;
    }
    p.H = in.H;
    p.W = in.W;
    p.B = B;
    p.tw = tw;
    p.th = th;
    p.nb = nb;
    p.tiles_x = in.W / tw;
    p.tiles_y = in.H / th;
    p.tiles_b = (B + nb - 1) / nb;
    p.n_tiles_n = CoutPad / BN;
    p.scale1 = hg->d_a + a1.scale_off;
    p.shift1 = hg->d_a + a1.shift_off;
    p.relu1 = relu1 ? 1 : 0;
    if (residual && residual->valid) {
      p.residual = reinterpret_cast<const __nv_bfloat16*>(ptr(*residual));
      p.res_ld = residual->C;
      if ((err = make_tmap_act(&p.tmRes, ptr(*residual), residual->C, in.W, in.H, B, tw, th, nb))) return;
    }
    if (res2_half && res2_half->valid) {
      p.has_res2 = 1;
      if ((err = make_tmap_box(&p.tmRes2, ptr(*res2_half), res2_half->C, in.W / 2, in.H / 2, B, tw / 2, th / 2, nb))) return;
    }
    if (out_raw && out_raw->valid) {
      p.out_raw = reinterpret_cast<__nv_bfloat16*>(ptr(*out_raw));
      p.raw_ld = out_raw->C;
      if ((err = make_tmap_act(&p.tmRaw, ptr(*out_raw), out_raw->C, in.W, in.H, B, tw, th, nb))) return;
    }
    if (out_act && out_act->valid && a2) {
      p.out_act = reinterpret_cast<__nv_bfloat16*>(ptr(*out_act));
      p.act_ld = out_act->C;
      p.scale2 = hg->d_a + a2->scale_off;
      p.shift2 = hg->d_a + a2->shift_off;
      if ((err = make_tmap_act(&p.tmAct, ptr(*out_act), out_act->C, in.W, in.H, B, tw, th, nb))) return;
    }
    if (out_f32 && out_f32->valid) {
      p.out_f32 = reinterpret_cast<float*>(ptr(*out_f32));
      p.f32_ld = out_f32->C;
    }
    if (pool_raw && pool_raw->valid && pool_act && pool_act->valid && a2) {  // 2x2 max-pool in the epilogue
      p.pool2 = 1;
      p.out_raw = reinterpret_cast<__nv_bfloat16*>(ptr(*pool_raw));  // marks "bf16 output"; the stores go through tmPool*
      p.raw_ld = pool_raw->C;
      p.scale2 = hg->d_a + a2->scale_off;
      p.shift2 = hg->d_a + a2->shift_off;
      if ((err = make_tmap_box(&p.tmPoolRaw, ptr(*pool_raw), pool_raw->C, in.W / 2, in.H / 2, B, tw / 2, th / 2, nb))) return;
      if ((err = make_tmap_box(&p.tmPoolAct, ptr(*pool_act), pool_act->C, in.W / 2, in.H / 2, B, tw / 2, th / 2, nb))) return;
    }
    hg->ops.push_back(op);
  }

  // merged inter-stack re-injection weights (see run()); cached on the handle so the views stay valid
  Convp merged_skip(const StackP& s, int i) {
    const int K = s.score.cout;
    if ((int)hg->merged_w.size() <= i) {
      hg->merged_w.resize(i + 1);
      hg->merged_b.resize(i + 1);
    }
    std::vector<float>& W = hg->merged_w[i];
    std::vector<float>& b = hg->merged_b[i];
    if (W.empty()) {
      W.assign((size_t)kCh * kCh, 0.0f);
      b.assign(kCh, 0.0f);
      for (int o = 0; o < kCh; ++o) {
        double bb = (double)s.fc_.bias[o] + s.score_.bias[o];
        for (int k = 0; k < K; ++k) bb += (double)s.score_.w[(size_t)o * K + k] * s.score.bias[k];
        b[o] = (float)bb;
        for (int c = 0; c < kCh; ++c) {
          double acc = s.fc_.w[(size_t)o * kCh + c];
          for (int k = 0; k < K; ++k) acc += (double)s.score_.w[(size_t)o * K + k] * s.score.w[(size_t)k * kCh + c];
          W[(size_t)o * kCh + c] = (float)acc;
        }
      }
    }
    Convp m;
    m.w = W.data();
    m.bias = b.data();
    m.cout = kCh;
    m.cin = kCh;
    m.k = 1;
    return m;
  }

  // fc(conv3(t) + b3 + h) + b_fc = (W_fc W_3) t + W_fc h + (W_fc b3 + b_fc): the residual bottleneck's last conv and the
  // fc conv behind it have nothing non-linear between them.  Returns the 256 x 128 product (fp64 accumulate, one
  // rounding to bf16 at packing time) with the combined bias; cached on the handle so the views stay valid.
  Convp merged_fc(const StackP& s, int i) {
    if ((int)hg->fcm_w.size() <= i) {
      hg->fcm_w.resize(i + 1);
      hg->fcm_b.resize(i + 1);
    }
    std::vector<float>& W = hg->fcm_w[i];
    std::vector<float>& b = hg->fcm_b[i];
    if (W.empty()) {
      W.assign((size_t)kCh * kFeats, 0.0f);
      b.assign(kCh, 0.0f);
      for (int o = 0; o < kCh; ++o) {
        double bb = s.fc.bias[o];
        for (int m = 0; m < kCh; ++m) bb += (double)s.fc.w[(size_t)o * kCh + m] * s.res.c3.bias[m];
        b[o] = (float)bb;
        for (int c = 0; c < kFeats; ++c) {
          double acc = 0.0;
          for (int m = 0; m < kCh; ++m) acc += (double)s.fc.w[(size_t)o * kCh + m] * s.res.c3.w[(size_t)m * kFeats + c];
          W[(size_t)o * kFeats + c] = (float)acc;
        }
      }
    }
    Convp m;
    m.w = W.data();
    m.bias = b.data();
    m.cout = kCh;
    m.cin = kFeats;
    m.k = 1;
    return m;
  }

  static int bn_for(int cout_pad) { return cout_pad >= 256 ? 256 : cout_pad; }
  static double fpp(const Convp& c) { return 2.0 * c.cin * c.cout * c.k * c.k; }

  // pre-activation bottleneck: x (raw) / xa = relu(bn1(x)) -> y (raw) [+ ya = relu(next_bn(y))]
  void bottleneck(const Bott& b, const Tensor& x, const Tensor& xa, const BNp* next_bn, Tensor* y, Tensor* ya,
                  const Tensor* up_add = nullptr) {
    const int H = x.H, W = x.W, P = b.planes, O = 2 * b.planes;
    // conv1 (1x1) with bn2+relu folded into its epilogue
    size_t w1 = pack_weights(b.c1, P, b.inpl);
    Affine a1 = conv_affine(b.c1, &b.bn2, P);
    Tensor t1 = talloc(H, W, P);
    conv(xa, w1, 1, b.inpl, P, bn_for(P), a1, true, nullptr, &t1, nullptr, nullptr, nullptr, fpp(b.c1));
    // conv2 (3x3) with bn3+relu folded
    size_t w2 = pack_weights(b.c2, P, P);
    Affine a2 = conv_affine(b.c2, &b.bn3, P);
    Tensor t2 = talloc(H, W, P);
    conv(t1, w2, 9, P, P, bn_for(P), a2, true, nullptr, &t2, nullptr, nullptr, nullptr, fpp(b.c2));
    tfree(t1);
    // projection shortcut on the raw input
    Tensor d;
    const Tensor* res = &x;
    if (b.has_ds) {
      size_t wd = pack_weights(b.ds, O, b.inpl);
      Affine ad = conv_affine(b.ds, nullptr, O);
      d = talloc(H, W, O);
      conv(x, wd, 1, b.inpl, O, bn_for(O), ad, false, nullptr, &d, nullptr, nullptr, nullptr, fpp(b.ds));
      res = &d;
    }
    // conv3 (1x1) + residual, optionally emitting the next block's activated input
    size_t w3 = pack_weights(b.c3, O, P);
    Affine a3 = conv_affine(b.c3, nullptr, O);
    *y = talloc(H, W, O);
    Affine an{};
    if (next_bn) {
      an = bn_affine(*next_bn, O);
      *ya = talloc(H, W, O);
    } else {
      *ya = Tensor();
    }
    conv(t2, w3, 1, P, O, bn_for(O), a3, false, res, y, next_bn ? &an : nullptr, ya, nullptr, fpp(b.c3), up_add);
    tfree(t2);
    tfree(d);
  }

  // ------------------------------------------------------------------ fused plan (conv chains)
  struct StageSpec {
    size_t w_off = 0;
    int K = 0, N = 0;  // packed weights [N][K]
    Affine a1{};
    bool unit_scale = false;  // a1 comes from a conv without BatchNorm: scale == 1
    bool relu1 = false;
    const Tensor* residual = nullptr;
    const Tensor* res2_half = nullptr;
    const Tensor* out_raw = nullptr;
    int x_src = 0;     // 0: last stage, 1: next stage takes bf16(v), 2: relu(bn(bf16(v))) with a2
    Affine a2{};
    double flop_per_px = 0.0;
    // 2x2 max-pool of out_raw in the epilogue: pooled raw / relu(bn(pooled)) tensors at half resolution
    const Tensor* pool_raw = nullptr;
    const Tensor* pool_act = nullptr;
    Affine pool_aff{};
    // the last ss_k of the K weight columns multiply a second activation tensor (read from shared memory) instead of
    // the previous stage's output
    const Tensor* in2 = nullptr;
    int ss_k = 0;
  };

  // the chain kernel pools in its epilogue on 8-wide tiles (halo-mode 8 x 16 tiles, or maps 8 pixels wide)
  bool chain_pools(int H, int W) const {
    if (hg->fuse < 2 || getenv("DF3D_HG_NO_POOL_FUSE")) return false;
    const bool halo = H >= 16 && W >= 8 && H % 16 == 0 && W % 8 == 0 && !getenv("DF3D_HG_NO_HALO");
    return (halo || W == 8) && H % 2 == 0;
  }

  // One chain launch: head conv (taps x CinPad from `in`) + point-wise stages on the same tiles.
  void chain(const Tensor& in, int taps, int CinPad, const StageSpec* sp, int n) {
    ++n_ops;
    if (err || dry) return;
    Op op;
    op.kind = OP_CHAIN;
    ChainParams& p = op.chain;
    memset(&p, 0, sizeof(p));
    int tw, th, nb;
    tile_geometry(in.H, in.W, &tw, &th, &nb);
    // halo mode of the chain kernel: 8 x 16 tiles whose 3x3 head reads row-shifted views of one halo tile
    const bool halo = taps == 9 && CinPad == 128 && sp[0].N == 128 && in.H >= 16 && in.W >= 8 && in.H % 16 == 0 &&
                      in.W % 8 == 0 && !getenv("DF3D_HG_NO_HALO");
    if (halo) {
      tw = 8;
      th = 16;
      nb = 1;
    }
    op.nb = nb;
    if ((err = make_tmap_act(&p.tmA, ptr(in), in.C, in.W, in.H, B, tw, th, nb))) return;
    if (halo) {
      p.halo = 1;
      if ((err = make_tmap_box(&p.tmHalo, ptr(in), in.C, in.W, in.H, B, tw + 2, th + 2, 1))) return;
    }
    p.n_chain = n;
    p.taps = taps;
    p.kc_per_tap = CinPad / 64;
    p.H = in.H;
    p.W = in.W;
    p.B = B;
    p.tw = tw;
    p.th = th;
    p.nb = nb;
    p.tiles_x = in.W / tw;
    p.tiles_y = in.H / th;
    p.tiles_b = (B + nb - 1) / nb;
    double bytes_px = 2.0 * in.C;
    for (int i = 0; i < n; ++i) {
      const StageSpec& s = sp[i];
      ChainStage& st = p.st[i];
      if ((err = make_tmap_wgt(&st.tmB, hg->d_w + s.w_off, s.K, s.N, 64))) return;  // 64 rows: one CTA's half of an N = 128 block
      st.n = s.N;
      st.kblocks = (s.K - s.ss_k) / 64;
      if (s.ss_k) {
        st.ss_kblocks = s.ss_k / 64;
        if ((err = make_tmap_act(&st.tmA2, ptr(*s.in2), s.in2->C, in.W, in.H, B, tw, th, nb))) return;
        bytes_px += 2.0 * s.ss_k;
      }
      st.relu1 = s.relu1 ? 1 : 0;
      st.unit_scale = s.unit_scale ? 1 : 0;
      st.scale1 = hg->d_a + s.a1.scale_off;
      st.shift1 = hg->d_a + s.a1.shift_off;
      st.x_src = s.x_src;
      if (s.x_src == 2) {
        st.scale2 = hg->d_a + s.a2.scale_off;
        st.shift2 = hg->d_a + s.a2.shift_off;
      }
      if (s.residual && s.residual->valid) {
        st.has_res = 1;
        if ((err = make_tmap_act(&st.tmRes, ptr(*s.residual), s.residual->C, in.W, in.H, B, tw, th, nb))) return;
        bytes_px += 2.0 * s.N;
      }
      if (s.res2_half && s.res2_half->valid) {
        st.has_res2 = 1;
        if ((err = make_tmap_box(&st.tmRes2, ptr(*s.res2_half), s.res2_half->C, in.W / 2, in.H / 2, B, tw / 2, th / 2, nb))) return;
        bytes_px += 0.5 * s.N;
      }
      if (s.out_raw && s.out_raw->valid) {
        if (s.out_raw->C != s.N) {
          err = DF3D_EINVAL;
          set_error("chain: stage %d stores %d channels into a %d-channel tensor", i, s.N, s.out_raw->C);
          return;
        }
        st.out_raw = reinterpret_cast<__nv_bfloat16*>(ptr(*s.out_raw));
        if (s.out_raw && (err = make_tmap_quarter(&st.tmOutQ, ptr(*s.out_raw), s.N, in.W, in.H, B, tw, th, nb))) return;
        bytes_px += 2.0 * s.N;
      }
      if (s.pool_raw && s.pool_raw->valid && s.pool_act && s.pool_act->valid) {
        st.pool_raw = reinterpret_cast<__nv_bfloat16*>(ptr(*s.pool_raw));
        st.pool_act = reinterpret_cast<__nv_bfloat16*>(ptr(*s.pool_act));
        st.pool_scale = hg->d_a + s.pool_aff.scale_off;
        st.pool_shift = hg->d_a + s.pool_aff.shift_off;
        bytes_px += 2.0 * s.N * 0.5;  // two tensors at a quarter of the pixels
      }
      op.flops_per_image += s.flop_per_px * in.H * in.W;
    }
    op.chain_bytes_per_image = bytes_px * in.H * in.W;
    hg->ops.push_back(op);
  }

  // stand-alone conv1 of a bottleneck (input already activated): t1 = relu(bn2(conv1(xa)))
  Tensor conv1(const Bott& b, const Tensor& xa) {
    size_t w1 = pack_weights(b.c1, b.planes, b.inpl);
    Affine a1 = conv_affine(b.c1, &b.bn2, b.planes);
    Tensor t1 = talloc(xa.H, xa.W, b.planes);
    conv(xa, w1, 1, b.inpl, b.planes, bn_for(b.planes), a1, true, nullptr, &t1, nullptr, nullptr, nullptr, fpp(b.c1));
    return t1;
  }

  // Everything of bottleneck `b` behind its conv1, plus conv1 of the block that consumes its output:
  //   y = conv3(relu(bn3(conv2(t1)))) + res (+ nearest_x2(up_add));   t1n = relu(next.bn2(next.conv1(relu(next.bn1(y)))))
  // fuse == 2: one chain launch [3x3 -> conv3 -> next.conv1];  fuse == 1: 3x3 alone, then [conv3 -> next.conv1].
  // t1 is consumed (freed).  `res` must have 2*planes channels (the caller applies a projection shortcut).
  // pool_bn != null: y is also needed 2x2 max-pooled (raw in *pool_p, relu(pool_bn(.)) in *pool_pa); the chain's
  // epilogue produces them when it can (chain_pools), otherwise the stand-alone pool kernel runs behind it.
  void tail(const Bott& b, const Tensor& res, Tensor& t1, const Bott* next, Tensor* y, Tensor* t1n,
            const Tensor* up_add = nullptr, const BNp* pool_bn = nullptr, Tensor* pool_p = nullptr, Tensor* pool_pa = nullptr) {
    const int H = res.H, W = res.W, P = b.planes, O = 2 * b.planes;
    StageSpec sp[3];
    int n = 0;
    size_t w2 = pack_weights(b.c2, P, P);
    Affine a2 = conv_affine(b.c2, &b.bn3, P);
    Tensor t2;
    if (hg->fuse >= 2) {
      StageSpec& s = sp[n++];
      s.w_off = w2;
      s.K = 9 * P;
      s.N = P;
      s.a1 = a2;
      s.relu1 = true;
      s.x_src = 1;
      s.flop_per_px = fpp(b.c2);
    } else {
      t2 = talloc(H, W, P);
      conv(t1, w2, 9, P, P, bn_for(P), a2, true, nullptr, &t2, nullptr, nullptr, nullptr, fpp(b.c2));
      tfree(t1);
    }
    size_t w3 = pack_weights(b.c3, O, P);
    Affine a3 = conv_affine(b.c3, nullptr, O);
    *y = talloc(H, W, O);
    Affine an{};
    size_t w1n = 0;
    Affine a1n{};
    if (next) {
      an = bn_affine(next->bn1, O);
      w1n = pack_weights(next->c1, next->planes, next->inpl);
      a1n = conv_affine(next->c1, &next->bn2, next->planes);
      *t1n = talloc(H, W, next->planes);
    }
    if (hg->fuse < 2 && !next) {  // nothing to chain behind conv3: the stand-alone kernel does it
      conv(t2, w3, 1, P, O, bn_for(O), a3, false, &res, y, nullptr, nullptr, nullptr, fpp(b.c3), up_add);
      tfree(t2);
      return;
    }
    {
      StageSpec& s = sp[n++];
      s.w_off = w3;
      s.K = P;
      s.N = O;
      s.a1 = a3;
      s.unit_scale = true;
      s.residual = &res;
      s.res2_half = up_add;
      s.out_raw = y;
      s.x_src = next ? 2 : 0;
      s.a2 = an;
      s.flop_per_px = fpp(b.c3);
      if (pool_bn && chain_pools(H, W)) {
        s.pool_aff = bn_affine(*pool_bn, O);
        *pool_p = talloc(H / 2, W / 2, O);
        *pool_pa = talloc(H / 2, W / 2, O);
        s.pool_raw = pool_p;
        s.pool_act = pool_pa;
      }
    }
    if (next) {
      StageSpec& s = sp[n++];
      s.w_off = w1n;
      s.K = O;
      s.N = next->planes;
      s.a1 = a1n;
      s.relu1 = true;
      s.out_raw = t1n;
      s.flop_per_px = fpp(next->c1);
    }
    if (hg->fuse >= 2) {
      chain(t1, 9, P, sp, n);
      tfree(t1);
    } else {
      chain(t2, 1, P, sp, n);
      tfree(t2);
    }
    if (pool_bn && !pool_p->valid) pool(*y, *pool_bn, pool_p, pool_pa);
  }

  // fused form of hourglass(): x raw, t1_up = conv1 of hg[n-1][0] already applied to x.
  // p / pa: x max-pooled (raw, and activated by hg[n-1][1].bn1), produced by whoever produced x; consumed here
  void hourglass_f(const StackP& s, int n, const Tensor& x, Tensor& t1_up, const Bott* next, Tensor* o, Tensor* t1_o,
                   Tensor& p, Tensor& pa) {
    Tensor t1 = conv1(s.hg[n - 1][1], pa);
    tfree(pa);
    Tensor l1, t1n, lp, lpa;
    tail(s.hg[n - 1][1], p, t1, n > 1 ? &s.hg[n - 2][0] : &s.hg[0][3], &l1, &t1n, nullptr,
         n > 1 ? &s.hg[n - 2][1].bn1 : nullptr, &lp, &lpa);
    tfree(p);
    Tensor l2, t1l3;
    if (n > 1)
      hourglass_f(s, n - 1, l1, t1n, &s.hg[n - 1][2], &l2, &t1l3, lp, lpa);
    else
      tail(s.hg[0][3], l1, t1n, &s.hg[0][2], &l2, &t1l3);
    tfree(l1);
    Tensor l3, none;
    tail(s.hg[n - 1][2], l2, t1l3, nullptr, &l3, &none);
    tfree(l2);
    tail(s.hg[n - 1][0], x, t1_up, next, o, t1_o, &l3);
    tfree(l3);
  }

  // stacks of the fused plan; x0 = output of layer3 (raw), t1 = conv1 of stack 0's hg[3][0] applied to it
  void stacks_fused(Tensor x0, Tensor t1, Tensor xp, Tensor xpa) {
    const df3d_hg_desc& d = hg->desc;
    const NetP& net = hg->net;
    const int H4 = d.in_h / 4, W4 = d.in_w / 4;
    const int S = d.num_stacks;
    for (int i = 0; i < S; ++i) {
      const StackP& s = net.stacks[i];
      Tensor h, t1r;
      hourglass_f(s, kDepth, x0, t1, &s.res, &h, &t1r, xp, xpa);
      const bool last = i == S - 1;
      // [res.conv2 ->] res.conv3 + h -> fc + BN + ReLU [-> merged re-injection + x0 -> next stack's first conv1]
      const Bott& b = s.res;
      StageSpec sp[5];
      int n = 0;
      size_t w2 = pack_weights(b.c2, kFeats, kFeats);
      Affine a2 = conv_affine(b.c2, &b.bn3, kFeats);
      Tensor t2;
      if (hg->fuse >= 2) {
        StageSpec& q = sp[n++];
        q.w_off = w2;
        q.K = 9 * kFeats;
        q.N = kFeats;
        q.a1 = a2;
        q.relu1 = true;
        q.x_src = 1;
        q.flop_per_px = fpp(b.c2);
      } else {
        t2 = talloc(H4, W4, kFeats);
        conv(t1r, w2, 9, kFeats, kFeats, 128, a2, true, nullptr, &t2, nullptr, nullptr, nullptr, fpp(b.c2));
        tfree(t1r);
      }
      const bool fc_merge = hg->fuse >= 2 && !getenv("DF3D_HG_NO_FC_MERGE");
      if (!fc_merge) {
        StageSpec& q = sp[n++];  // r = conv3 + h (feeds fc raw: fc is conv -> BN -> ReLU)
        q.w_off = pack_weights(b.c3, kCh, kFeats);
        q.K = kFeats;
        q.N = kCh;
        q.a1 = conv_affine(b.c3, nullptr, kCh);
        q.unit_scale = true;
        q.residual = &h;
        q.x_src = 1;
        q.flop_per_px = fpp(b.c3);
      }
      Tensor f, nx, t1x;
      {
        StageSpec& q = sp[n++];  // f = relu(bn(fc(r)))
        if (fc_merge) {
          // ... with r never formed: (W_fc W_3) t2 from tensor memory + W_fc h from shared memory (see merged_fc)
          Convp m = merged_fc(s, i);
          q.w_off = pack_weights_concat(m, s.fc, kCh);
          q.K = kFeats + kCh;
          q.ss_k = kCh;
          q.in2 = &h;
          q.a1 = conv_affine(m, &s.fc_bn, kCh);
          q.flop_per_px = fpp(b.c3) + fpp(s.fc);
        } else {
          q.w_off = pack_weights(s.fc, kCh, kCh);
          q.K = kCh;
          q.a1 = conv_affine(s.fc, &s.fc_bn, kCh);
          q.flop_per_px = fpp(s.fc);
        }
        q.N = kCh;
        q.relu1 = true;
        if (last) {
          f = talloc(H4, W4, kCh);
          q.out_raw = &f;
        } else {
          q.x_src = 1;
        }
      }
      if (!last) {
        const Bott& nb0 = net.stacks[i + 1].hg[kDepth - 1][0];
        Convp merged = merged_skip(s, i);
        {
          StageSpec& q = sp[n++];  // x' = x + fc_(f) + score_(score(f)) as one merged conv (see merged_skip)
          q.w_off = pack_weights(merged, kCh, kCh);
          q.K = kCh;
          q.N = kCh;
          q.a1 = conv_affine(merged, nullptr, kCh);
          q.unit_scale = true;
          q.residual = &x0;
          nx = talloc(H4, W4, kCh);
          q.out_raw = &nx;
          q.x_src = 2;
          q.a2 = bn_affine(nb0.bn1, kCh);
          q.flop_per_px = fpp(s.fc_) + fpp(s.score) + fpp(s.score_);
          if (chain_pools(H4, W4)) {  // the next stack pools its input first: done here, in this stage's epilogue
            q.pool_aff = bn_affine(net.stacks[i + 1].hg[kDepth - 1][1].bn1, kCh);
            xp = talloc(H4 / 2, W4 / 2, kCh);
            xpa = talloc(H4 / 2, W4 / 2, kCh);
            q.pool_raw = &xp;
            q.pool_act = &xpa;
          }
        }
        {
          StageSpec& q = sp[n++];  // first conv1 of the next stack
          q.w_off = pack_weights(nb0.c1, nb0.planes, nb0.inpl);
          q.K = kCh;
          q.N = nb0.planes;
          q.a1 = conv_affine(nb0.c1, &nb0.bn2, nb0.planes);
          q.relu1 = true;
          t1x = talloc(H4, W4, nb0.planes);
          q.out_raw = &t1x;
          q.flop_per_px = fpp(nb0.c1);
        }
      }
      if (hg->fuse >= 2) {
        chain(t1r, 9, kFeats, sp, n);
        tfree(t1r);
      } else {
        chain(t2, 1, kFeats, sp, n);
        tfree(t2);
      }
      tfree(h);
      tfree(x0);
      if (last) {
        size_t wsc = pack_weights(s.score, kHeatPad, kCh);
        Affine as = conv_affine(s.score, nullptr, kHeatPad);
        Tensor heat = talloc(H4, W4, kHeatPad, 4);
        conv(f, wsc, 1, kCh, kHeatPad, kHeatPad, as, false, nullptr, nullptr, nullptr, nullptr, &heat, fpp(s.score));
        if (!dry && !err) {  // the score head: arg-max fused into its epilogue (see forward)
          hg->ops.back().variant = -1;
          hg->score_nb = hg->ops.back().nb;
        }
        tfree(f);
        if (!dry && !err) {
          Op op;
          op.kind = OP_ARGMAX;
          op.in0 = ptr(heat);
          op.H = H4;
          op.W = W4;
          op.C = kHeatPad;
          hg->ops.push_back(op);
          hg->heat_ptr = ptr(heat);
        }
        tfree(heat);
      } else {
        x0 = nx;
        t1 = t1x;
        if (!xp.valid) pool(x0, net.stacks[i + 1].hg[kDepth - 1][1].bn1, &xp, &xpa);
      }
    }
  }

  void pool(const Tensor& x, const BNp& bn, Tensor* p, Tensor* pa) {
    Affine a = bn_affine(bn, x.C);
    *p = talloc(x.H / 2, x.W / 2, x.C);
    *pa = talloc(x.H / 2, x.W / 2, x.C);
    ++n_ops;
    if (dry || err) return;
    Op op;
    op.kind = OP_POOL;
    op.in0 = ptr(x);
    op.out0 = ptr(*p);
    op.out1 = ptr(*pa);
    op.scale = hg->d_a + a.scale_off;
    op.shift = hg->d_a + a.shift_off;
    op.H = x.H;
    op.W = x.W;
    op.C = x.C;
    hg->ops.push_back(op);
  }

  // level n of the recursive hourglass of stack s; x/xa stay owned by the caller.
  //   out = up1 + nearest_x2(low3),  up1 = hg[n-1][0](x),  low3 = hg[n-1][2](low2(hg[n-1][1](pool(x))))
  // The low path runs first; the up-sample + add is folded into the epilogue of up1's last conv
  // (second, half-resolution residual), so neither up1 nor a separate add pass touches HBM.
  void hourglass(const StackP& s, int n, const Tensor& x, const Tensor& xa, const BNp& out_bn, Tensor* o, Tensor* oa) {
    Tensor none;
    Tensor p, pa;
    pool(x, s.hg[n - 1][1].bn1, &p, &pa);
    Tensor l1, l1a;
    // low1 feeds (n > 1) the next level -- its pool takes the raw tensor, its up1 bottleneck the
    // activated one -- or (n == 1) hg[0][3]
    const BNp& low1_act_bn = (n > 1) ? s.hg[n - 2][0].bn1 : s.hg[0][3].bn1;
    bottleneck(s.hg[n - 1][1], p, pa, &low1_act_bn, &l1, &l1a);
    tfree(p);
    tfree(pa);
    Tensor l2, l2a;
    if (n > 1)
      hourglass(s, n - 1, l1, l1a, s.hg[n - 1][2].bn1, &l2, &l2a);
    else
      bottleneck(s.hg[0][3], l1, l1a, &s.hg[0][2].bn1, &l2, &l2a);
    tfree(l1);
    tfree(l1a);
    Tensor l3;
    bottleneck(s.hg[n - 1][2], l2, l2a, nullptr, &l3, &none);
    tfree(l2);
    tfree(l2a);
    bottleneck(s.hg[n - 1][0], x, xa, &out_bn, o, oa, &l3);
    tfree(l3);
  }

  void run() {
    const df3d_hg_desc& d = hg->desc;
    const NetP& net = hg->net;
    const int K = d.num_classes;
    const int H2 = d.in_h / 2, W2 = d.in_w / 2, H4 = d.in_h / 4, W4 = d.in_w / 4;
    hg->Hh = H4;
    hg->Wh = W4;

    // ---- stem: im2col + GEMM with bn1+relu folded; also emits layer1's activated input
    Tensor col = talloc(H2, W2, kStemKPadCols);
    n_ops += 2;  // im2col + argmax
    if (!dry && !err) {
      Op op;
      op.kind = OP_IM2COL;
      op.out0 = ptr(col);
      op.variant = 1;
      hg->ops.push_back(op);
    }
    size_t w0 = pack_stem(net.conv1);
    Affine a0 = conv_affine(net.conv1, &net.bn1, kInplanes);
    Affine a0n = bn_affine(net.layer1.bn1, kInplanes);
    Tensor x = talloc(H2, W2, kInplanes), xa = talloc(H2, W2, kInplanes);
    conv(col, w0, 1, kStemKPadCols, kInplanes, 64, a0, true, nullptr, &x, &a0n, &xa, nullptr, fpp(net.conv1));
    if (!dry && !err) hg->ops.back().variant = 1;
    // gray fast path of the same two launches (chosen per call in df3d_hg_forward_argmax)
    Tensor colg = col;
    colg.C = kStemKGray;
    n_ops += 1;
    if (!dry && !err) {
      Op op;
      op.kind = OP_IM2COL_GRAY;
      op.out0 = ptr(colg);
      op.variant = 2;
      hg->ops.push_back(op);
    }
    size_t w0g = pack_stem_gray(net.conv1);
    conv(colg, w0g, 1, kStemKGray, kInplanes, 64, a0, true, nullptr, &x, &a0n, &xa, nullptr, fpp(net.conv1));
    if (!dry && !err) hg->ops.back().variant = 2;
    tfree(col);

    Tensor none;
    Tensor p, pa;
    if (getenv("DF3D_HG_NO_FRONT_FUSE")) {  // profiling knob: the unfused front section (three more launches)
      Tensor y1;
      bottleneck(net.layer1, x, xa, nullptr, &y1, &none);
      tfree(x);
      tfree(xa);
      pool(y1, net.layer2.bn1, &p, &pa);
      tfree(y1);
    } else {
      // layer1 (64 -> 128 channels at half resolution, projection shortcut) + the 2x2 max-pool behind it:
      //   conv1, conv2 as usual; conv3(t2) + downsample(x) is ONE GEMM over K = [t2 | x] (weights side by side, both
      //   biases in the shift), and its epilogue pools the tile and applies layer2's first BatchNorm + ReLU -- the
      //   full-resolution 128-channel tensor, the shortcut tensor and the pool launch never exist in HBM
      const Bott& b = net.layer1;
      const int P = b.planes, O = 2 * b.planes;
      size_t w1 = pack_weights(b.c1, P, b.inpl);
      Affine a1 = conv_affine(b.c1, &b.bn2, P);
      Tensor t1 = talloc(H2, W2, P);
      conv(xa, w1, 1, b.inpl, P, bn_for(P), a1, true, nullptr, &t1, nullptr, nullptr, nullptr, fpp(b.c1));
      tfree(xa);
      size_t w2 = pack_weights(b.c2, P, P);
      Affine a2 = conv_affine(b.c2, &b.bn3, P);
      Tensor t2 = talloc(H2, W2, P);
      conv(t1, w2, 9, P, P, bn_for(P), a2, true, nullptr, &t2, nullptr, nullptr, nullptr, fpp(b.c2));
      tfree(t1);
      size_t w3 = pack_weights_concat(b.c3, b.ds, O);
      Affine a3 = concat_affine(b.c3, b.ds, O);
      Affine an = bn_affine(net.layer2.bn1, O);
      p = talloc(H4, W4, O);
      pa = talloc(H4, W4, O);
      conv(t2, w3, 1, P + b.inpl, O, bn_for(O), a3, false, nullptr, nullptr, &an, nullptr, nullptr, fpp(b.c3) + fpp(b.ds), nullptr, &x,
           &p, &pa);
      tfree(t2);
      tfree(x);
    }
    Tensor y2, y2a;
    bottleneck(net.layer2, p, pa, &net.layer3.bn1, &y2, &y2a);
    tfree(p);
    tfree(pa);
    if (hg->fuse > 0) {
      // layer3 through the chain path: conv1 and the projection shortcut stand alone, the rest is a chain
      // that ends in conv1 of the first hourglass bottleneck
      Tensor t1 = conv1(net.layer3, y2a);
      tfree(y2a);
      size_t wd = pack_weights(net.layer3.ds, kCh, 2 * kInplanes);
      Affine ad = conv_affine(net.layer3.ds, nullptr, kCh);
      Tensor dres = talloc(H4, W4, kCh);
      conv(y2, wd, 1, 2 * kInplanes, kCh, 256, ad, false, nullptr, &dres, nullptr, nullptr, nullptr, fpp(net.layer3.ds));
      tfree(y2);
      Tensor xf, t1x, xp, xpa;
      tail(net.layer3, dres, t1, &net.stacks[0].hg[kDepth - 1][0], &xf, &t1x, nullptr, &net.stacks[0].hg[kDepth - 1][1].bn1, &xp, &xpa);
      tfree(dres);
      stacks_fused(xf, t1x, xp, xpa);
      return;
    }
    Tensor x0, x0a;
    bottleneck(net.layer3, y2, y2a, &net.stacks[0].hg[kDepth - 1][0].bn1, &x0, &x0a);
    tfree(y2);
    tfree(y2a);

    const int S = d.num_stacks;
    for (int i = 0; i < S; ++i) {
      const StackP& s = net.stacks[i];
      Tensor h, ha;
      hourglass(s, kDepth, x0, x0a, s.res.bn1, &h, &ha);
      Tensor r;
      bottleneck(s.res, h, ha, nullptr, &r, &none);
      tfree(h);
      tfree(ha);
      // fc: conv -> BN -> ReLU (post-activation), input is the raw residual stream
      size_t wf = pack_weights(s.fc, kCh, kCh);
      Affine af = conv_affine(s.fc, &s.fc_bn, kCh);
      Tensor f = talloc(H4, W4, kCh);
      conv(r, wf, 1, kCh, kCh, 256, af, true, nullptr, &f, nullptr, nullptr, nullptr, fpp(s.fc));
      tfree(r);
      if (i == S - 1) {
        size_t wsc = pack_weights(s.score, kHeatPad, kCh);
        Affine as = conv_affine(s.score, nullptr, kHeatPad);
        Tensor heat = talloc(H4, W4, kHeatPad, 4);
        conv(f, wsc, 1, kCh, kHeatPad, kHeatPad, as, false, nullptr, nullptr, nullptr, nullptr, &heat, fpp(s.score));
        if (!dry && !err) {  // the score head: arg-max fused into its epilogue (see forward)
          hg->ops.back().variant = -1;
          hg->score_nb = hg->ops.back().nb;
        }
        tfree(f);
        if (!dry && !err) {
          Op op;
          op.kind = OP_ARGMAX;
          op.in0 = ptr(heat);
          op.H = H4;
          op.W = W4;
          op.C = kHeatPad;
          hg->ops.push_back(op);
          hg->heat_ptr = ptr(heat);
        }
        tfree(heat);
        tfree(x0);
        tfree(x0a);
      } else {
        // x' = x + fc_(f) + score_(score(f)).  score is linear in f and nothing non-linear sits
        // between score and score_, so the three convs collapse into ONE 256->256 conv with
        //   W = W_fc_ + W_score_ @ W_score ,  b = b_fc_ + W_score_ @ b_score + b_score_
        // (merged in fp32 on the host; the intermediate stacks' score maps are not an output).
        Convp merged = merged_skip(s, i);
        size_t wm = pack_weights(merged, kCh, kCh);
        Affine am = conv_affine(merged, nullptr, kCh);
        Affine an = bn_affine(net.stacks[i + 1].hg[kDepth - 1][0].bn1, kCh);
        Tensor nx = talloc(H4, W4, kCh), nxa = talloc(H4, W4, kCh);
        conv(f, wm, 1, kCh, kCh, 256, am, false, &x0, &nx, &an, &nxa, nullptr,
             fpp(s.fc_) + fpp(s.score) + fpp(s.score_));
        tfree(f);
        tfree(x0);
        tfree(x0a);
        x0 = nx;
        x0a = nxa;
      }
    }
    (void)K;
  }
};

static size_t param_count(const df3d_hg_desc& d) {
  static const float dummy = 0.0f;  // count by walking the reader; the views are never dereferenced
  Cursor c{&dummy, (size_t)-1};
  NetP net;
  read_net(c, d.num_stacks, d.num_classes, &net);
  return c.pos;
}

static int check_desc(const df3d_hg_desc* d, const char* fn) {
  DF3D_REQUIRE(d, DF3D_EINVAL, "%s: null desc", fn);
  DF3D_REQUIRE(d->num_stacks >= 1 && d->num_stacks <= 16, DF3D_EINVAL, "%s: num_stacks must be in [1,16]", fn);
  DF3D_REQUIRE(d->num_classes >= 1 && d->num_classes <= kHeatPad, DF3D_EINVAL, "%s: num_classes must be in [1,%d]", fn, kHeatPad);
  DF3D_REQUIRE(d->in_h >= 64 && d->in_w >= 64 && d->in_h % 64 == 0 && d->in_w % 64 == 0 && d->in_h <= 4096 && d->in_w <= 4096,
               DF3D_EINVAL, "%s: input size must be a multiple of 64 in [64,4096]", fn);
  DF3D_REQUIRE(d->max_batch >= 1, DF3D_EINVAL, "%s: max_batch must be >= 1", fn);
  return DF3D_OK;
}

}  // namespace df3d

extern "C" size_t df3d_hg_param_count(const df3d_hg_desc* desc) {
  if (check_desc(desc, "df3d_hg_param_count")) return 0;
  return param_count(*desc);
}

static int lanes_for(const df3d_hg_desc& d) {
  // One lane by default.  DF3D_HG_LANES=n splits the batch over n streams with 1/n of the SMs each
  // so that tensor-bound 3x3 convs of one lane could overlap HBM-bound 1x1 convs of another;
  // measured on B200 (profiles/r01_lanes.txt) this is slower (242 / 281 / 330 ms for 1 / 2 / 3
  // lanes at 1792 images), so it stays a tuning knob only.
  (void)d;
  int l = 1;
  if (const char* env = getenv("DF3D_HG_LANES")) {
    const int v = atoi(env);
    if (v >= 1 && v <= 4) l = v;
  }
  return l;
}

static int fuse_for() {
  // 2 (default): 3x3-led conv chains; 1: point-wise chains behind stand-alone 3x3 convs; 0: one launch
  // per conv.  All three produce bit-identical results (tests/test_gpu_hourglass.py); the knob exists for
  // that test and for profiling.
  int f = 2;
  if (const char* env = getenv("DF3D_HG_FUSE")) {
    const int v = atoi(env);
    if (v >= 0 && v <= 2) f = v;
  }
  return f;
}

static int chunk_for(const df3d_hg_desc& d, int lanes) {
  // images per lane and launch sequence.  Measured on B200 (profiles/): activations never fit the
  // 126 MB L2 at any useful chunk, and every launch costs ~10 us of fill/drain, so bigger is
  // better; the cap keeps the whole workspace near 60 GB for 256x256 inputs.
  long long px = (long long)d.in_h * d.in_w;
  int c = (int)((1792ll * 256 * 256) / px) / lanes;
  if (c < 8) c = 8;
  if (const char* env = getenv("DF3D_HG_CHUNK")) {  // tuning / test knob
    const int v = atoi(env);
    if (v >= 1) c = v;
  }
  const int need = (d.max_batch + lanes - 1) / lanes;
  if (c > need) c = need;
  return c;
}

extern "C" size_t df3d_hg_workspace_bytes(const df3d_hg_desc* desc) {
  if (check_desc(desc, "df3d_hg_workspace_bytes")) return 0;
  df3d_hg tmp;
  tmp.desc = *desc;
  tmp.n_lanes = lanes_for(*desc);
  tmp.fuse = fuse_for();
  tmp.chunk = chunk_for(*desc, tmp.n_lanes);
  tmp.params.assign(param_count(*desc), 0.0f);
  Cursor c{tmp.params.data(), tmp.params.size()};
  read_net(c, desc->num_stacks, desc->num_classes, &tmp.net);
  Emitter e{&tmp, true, nullptr};
  e.B = tmp.chunk;
  e.run();
  return (size_t)tmp.n_lanes * ((e.arena.top + 1023) & ~size_t(1023)) + 1024;
}

extern "C" int df3d_hg_create(const df3d_hg_desc* desc, const float* params_host, size_t n_params, df3d_hg** out) {
  if (int e = check_desc(desc, "df3d_hg_create")) return e;
  DF3D_REQUIRE(params_host && out, DF3D_EINVAL, "df3d_hg_create: null pointer");
  const size_t need = param_count(*desc);
  DF3D_REQUIRE(n_params == need, DF3D_EINVAL, "df3d_hg_create: expected %zu parameters, got %zu", need, n_params);
  int dev = 0;
  DF3D_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DF3D_CUDA(cudaGetDeviceProperties(&prop, dev));
  DF3D_REQUIRE(prop.major == 10, DF3D_EUNSUPPORTED, "df3d_hg_create: needs an sm_100 GPU (found sm_%d%d); there is no fallback path",
               prop.major, prop.minor);
  if (int e = tma_init()) return e;
  if (int e = conv_gemm_configure()) return e;
  if (int e = conv_chain_configure()) return e;

  df3d_hg* hg = new df3d_hg();
  hg->desc = *desc;
  hg->fuse = fuse_for();
  hg->num_sms = prop.multiProcessorCount;
  hg->n_lanes = lanes_for(*desc);
  hg->chunk = chunk_for(*desc, hg->n_lanes);
  hg->params.assign(params_host, params_host + n_params);
  Cursor c{hg->params.data(), hg->params.size()};
  read_net(c, desc->num_stacks, desc->num_classes, &hg->net);
  if (!c.ok || c.pos != n_params) {
    delete hg;
    DF3D_REQUIRE(false, DF3D_EINVAL, "df3d_hg_create: parameter blob does not match the architecture");
  }
  Emitter e{hg, true, nullptr};
  e.B = hg->chunk;
  e.run();
  hg->lane_bytes = (e.arena.top + 1023) & ~size_t(1023);
  hg->ws_bytes = (size_t)hg->n_lanes * hg->lane_bytes + 1024;
  hg->ops_per_chunk = e.n_ops - 2;  // of the two stem variants (im2col + GEMM each) one runs
  cudaError_t ce = cudaMalloc(&hg->d_w, hg->wblob.size() * sizeof(uint16_t));
  if (ce == cudaSuccess && hg->n_lanes > 1) ce = cudaEventCreateWithFlags(&hg->fork_ev, cudaEventDisableTiming);
  for (int l = 1; l < hg->n_lanes && ce == cudaSuccess; ++l) {
    ce = cudaStreamCreateWithFlags(&hg->lane_stream[l], cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&hg->join_ev[l], cudaEventDisableTiming);
  }
  if (ce == cudaSuccess) ce = cudaMalloc(&hg->d_a, hg->ablob.size() * sizeof(float));
  if (ce == cudaSuccess) ce = cudaMalloc(&hg->d_keys, (size_t)hg->desc.max_batch * kHeatPad * sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMemset(hg->d_keys, 0, (size_t)hg->desc.max_batch * kHeatPad * sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMemcpy(hg->d_w, hg->wblob.data(), hg->wblob.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemcpy(hg->d_a, hg->ablob.data(), hg->ablob.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    set_error("df3d_hg_create: uploading packed weights failed: %s", cudaGetErrorString(ce));
    (void)cudaGetLastError();
    if (hg->d_w) cudaFree(hg->d_w);
    if (hg->d_a) cudaFree(hg->d_a);
    if (hg->d_keys) cudaFree(hg->d_keys);
    delete hg;
    return DF3D_ECUDA;
  }
  std::vector<uint16_t>().swap(hg->wblob);  // host copies no longer needed
  *out = hg;
  return DF3D_OK;
}

extern "C" void df3d_hg_destroy(df3d_hg* hg) {
  if (!hg) return;
  if (hg->d_w) cudaFree(hg->d_w);
  if (hg->d_a) cudaFree(hg->d_a);
  if (hg->d_keys) cudaFree(hg->d_keys);
  for (cudaEvent_t ev : hg->events) cudaEventDestroy(ev);
  if (hg->fork_ev) cudaEventDestroy(hg->fork_ev);
  for (int l = 1; l < 4; ++l) {
    if (hg->join_ev[l]) cudaEventDestroy(hg->join_ev[l]);
    if (hg->lane_stream[l]) cudaStreamDestroy(hg->lane_stream[l]);
  }
  delete hg;
}

static int build_plan(df3d_hg* hg, char* base) {
  hg->lane_ops.assign(hg->n_lanes, std::vector<Op>());
  for (int l = hg->n_lanes - 1; l >= 0; --l) {  // lane 0 last: hg->ops keeps its plan for the descriptions
    hg->ops.clear();
    hg->heat_ptr = nullptr;
    Emitter e{hg, false, base + (size_t)l * hg->lane_bytes};
    e.B = hg->chunk;
    e.run();
    if (e.err) return e.err;
    hg->lane_ops[l] = hg->ops;
  }
  hg->plan_base = base;
  return DF3D_OK;
}

// how a batch is cut: rounds of up to n_lanes * chunk images, split evenly over the lanes
struct Piece {
  int lane, c0, bc, active_lanes;
};
static std::vector<Piece> cut_batch(const df3d_hg* hg, int B) {
  std::vector<Piece> out;
  int pos = 0;
  while (pos < B) {
    const int round = (B - pos) < hg->n_lanes * hg->chunk ? (B - pos) : hg->n_lanes * hg->chunk;
    const int per = (round + hg->n_lanes - 1) / hg->n_lanes;
    int active = 0;
    for (int l = 0; l < hg->n_lanes; ++l)
      if (l * per < round) ++active;
    for (int l = 0; l < active; ++l) {
      const int c0 = pos + l * per;
      const int bc = (pos + round - c0) < per ? (pos + round - c0) : per;
      out.push_back(Piece{l, c0, bc, active});
    }
    pos += round;
  }
  return out;
}

extern "C" int df3d_hg_launches_per_forward(const df3d_hg* hg, int B) {
  if (!hg || B <= 0) return 0;
  return hg->ops_per_chunk * (int)cut_batch(hg, B).size();
}

extern "C" int df3d_hg_forward_argmax(df3d_hg* hg, const void* images_dev, int dtype, const uint8_t* flip_dev,
                                      int B, int32_t* idx_dev, float* conf_dev, float* heatmap_dev,
                                      void* workspace_dev, size_t workspace_bytes, void* stream) {
  DF3D_REQUIRE(hg && images_dev && idx_dev && conf_dev && workspace_dev, DF3D_EINVAL, "df3d_hg_forward_argmax: null pointer");
  DF3D_REQUIRE(dtype == 0 || dtype == 1, DF3D_EINVAL, "df3d_hg_forward_argmax: dtype must be 0 (uint8 gray) or 1 (float32 NCHW)");
  DF3D_REQUIRE(B >= 0 && B <= hg->desc.max_batch, DF3D_EINVAL, "df3d_hg_forward_argmax: B=%d exceeds max_batch=%d", B, hg->desc.max_batch);
  DF3D_REQUIRE(workspace_bytes >= hg->ws_bytes, DF3D_ENOMEM, "df3d_hg_forward_argmax: workspace too small (%zu < %zu bytes)",
               workspace_bytes, hg->ws_bytes);
  if (B == 0) return DF3D_OK;
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace_dev) + 1023) & ~uintptr_t(1023));
  if (hg->plan_base != base || hg->lane_ops.empty())
    if (int e = build_plan(hg, base)) return e;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const df3d_hg_desc& d = hg->desc;
  const int K = d.num_classes;
  const size_t img_stride = dtype == 0 ? (size_t)d.in_h * d.in_w : (size_t)3 * d.in_h * d.in_w * sizeof(float);
  const size_t heat_elems = (size_t)hg->Hh * hg->Wh * kHeatPad;

  const size_t n_ops = hg->ops.size();
  // gray fast path of the stem: uint8 gray input replicated to three planes with one common mean
  // (rows wider than the fast path's shared-memory window take the generic three-plane variant)
  const bool gray = dtype == 0 && hg->mean[0] == hg->mean[1] && hg->mean[1] == hg->mean[2] && d.in_w <= kStemGrayMaxW &&
                    !getenv("DF3D_HG_NO_GRAY");
  // the fused arg-max accumulates with atomicMax into keys that the decode kernel zeroes again; a forward that
  // failed half-way would leave them dirty, so they are cleared up front (B x 32 x 8 bytes)
  if (hg->score_nb == 1 || hg->lane_ops.empty())
    DF3D_CUDA(cudaMemsetAsync(hg->d_keys, 0, (size_t)B * kHeatPad * sizeof(unsigned long long), s));
  const std::vector<Piece> pieces = cut_batch(hg, B);
  if (hg->timing) {
    while (hg->events.size() < pieces.size() * n_ops * 2) {
      cudaEvent_t ev;
      DF3D_CUDA(cudaEventCreate(&ev));
      hg->events.push_back(ev);
    }
    hg->timed_bc.clear();
    hg->timed_stem_variant = gray ? 2 : 1;
  }
  for (size_t pi = 0; pi < pieces.size(); ++pi) {
    const Piece& pc = pieces[pi];
    const int c0 = pc.c0, bc = pc.bc;
    // lane 0 runs on the caller's stream; the other lanes fork from it and join back
    cudaStream_t ls = pc.lane == 0 ? s : hg->lane_stream[pc.lane];
    if (pc.lane == 0 && pc.active_lanes > 1) DF3D_CUDA(cudaEventRecord(hg->fork_ev, s));
    if (pc.lane > 0) DF3D_CUDA(cudaStreamWaitEvent(ls, hg->fork_ev, 0));
    const int sms = pc.active_lanes > 1 ? hg->num_sms / pc.active_lanes : hg->num_sms;
    const std::vector<Op>& ops = hg->lane_ops[pc.lane];
    const size_t ev_base = hg->timing ? hg->timed_bc.size() * n_ops * 2 : 0;
    if (hg->timing) hg->timed_bc.push_back(bc);
    for (size_t oi = 0; oi < n_ops; ++oi) {
      const Op& op = ops[oi];
      if (hg->timing) DF3D_CUDA(cudaEventRecord(hg->events[ev_base + 2 * oi], ls));
      if (op.variant > 0 && op.variant != (gray ? 2 : 1)) {  // the other stem variant
        if (hg->timing) DF3D_CUDA(cudaEventRecord(hg->events[ev_base + 2 * oi + 1], ls));
        continue;
      }
      switch (op.kind) {
        case OP_IM2COL_GRAY: {
          const uint8_t* img = static_cast<const uint8_t*>(images_dev) + (size_t)c0 * img_stride;
          if (int e = launch_stem_im2col_gray(img, flip_dev ? flip_dev + c0 : nullptr, bc, d.in_h, d.in_w, hg->mean[0],
                                              reinterpret_cast<__nv_bfloat16*>(op.out0), ls))
            return e;
          break;
        }
        case OP_IM2COL: {
          const char* img = static_cast<const char*>(images_dev) + (size_t)c0 * img_stride;
          if (int e = launch_stem_im2col(img, dtype, flip_dev ? flip_dev + c0 : nullptr, bc, d.in_h, d.in_w, hg->mean,
                                         reinterpret_cast<__nv_bfloat16*>(op.out0), ls))
            return e;
          break;
        }
        case OP_CONV: {
          ConvParams p = op.conv;
          p.B = bc;
          p.tiles_b = (bc + op.nb - 1) / op.nb;
          if (op.variant == -1 && op.nb == 1) {
            // score head: the arg-max is taken in the epilogue; the fp32 maps only go to HBM when
            // the caller asked for them
            p.amax_keys = hg->d_keys + (size_t)c0 * kHeatPad;
            p.amax_k = K;
            if (!heatmap_dev) p.out_f32 = nullptr;
          }
          if (int e = launch_conv_gemm(p, op.BN, sms, ls)) return e;
          break;
        }
        case OP_CHAIN: {
          ChainParams p = op.chain;
          p.B = bc;
          p.tiles_b = (bc + op.nb - 1) / op.nb;
          if (int e = launch_conv_chain(p, sms, ls)) return e;
          break;
        }
        case OP_POOL:
          if (int e = launch_maxpool_bn_relu(reinterpret_cast<const __nv_bfloat16*>(op.in0), bc, op.H, op.W, op.C, op.scale,
                                             op.shift, reinterpret_cast<__nv_bfloat16*>(op.out0),
                                             reinterpret_cast<__nv_bfloat16*>(op.out1), ls))
            return e;
          break;
        case OP_ARGMAX:
          if (hg->score_nb == 1) {  // keys written by the score head's epilogue
            if (int e = launch_argmax_keys_decode(hg->d_keys + (size_t)c0 * kHeatPad, bc, kHeatPad, K, idx_dev + (size_t)c0 * K,
                                                  conf_dev + (size_t)c0 * K, ls))
              return e;
          } else if (int e = df3d_heatmap_argmax_nhwc(reinterpret_cast<const float*>(op.in0), bc, op.H, op.W, op.C, K,
                                                      idx_dev + (size_t)c0 * K, conf_dev + (size_t)c0 * K, ls)) {
            return e;
          }
          if (heatmap_dev)
            DF3D_CUDA(cudaMemcpyAsync(heatmap_dev + (size_t)c0 * heat_elems, op.in0, (size_t)bc * heat_elems * sizeof(float),
                                      cudaMemcpyDeviceToDevice, ls));
          break;
      }
      if (hg->timing) DF3D_CUDA(cudaEventRecord(hg->events[ev_base + 2 * oi + 1], ls));
    }
    if (pc.lane > 0) {
      DF3D_CUDA(cudaEventRecord(hg->join_ev[pc.lane], ls));
      DF3D_CUDA(cudaStreamWaitEvent(s, hg->join_ev[pc.lane], 0));
    }
  }
  return DF3D_OK;
}

extern "C" int df3d_hg_set_timing(df3d_hg* hg, int enable) {
  DF3D_REQUIRE(hg, DF3D_EINVAL, "df3d_hg_set_timing: null handle");
  hg->timing = enable != 0;
  if (!hg->timing) {
    for (cudaEvent_t ev : hg->events) cudaEventDestroy(ev);
    hg->events.clear();
    hg->timed_bc.clear();
  }
  return DF3D_OK;
}

extern "C" int df3d_hg_read_timing(df3d_hg* hg, double* out8) {
  DF3D_REQUIRE(hg && out8, DF3D_EINVAL, "df3d_hg_read_timing: null pointer");
  DF3D_REQUIRE(hg->timing && !hg->timed_bc.empty(), DF3D_EINVAL, "df3d_hg_read_timing: no timed forward recorded");
  const size_t n_ops = hg->ops.size();
  double conv_ms = 0, conv_flops = 0, other_ms = 0, conv3_ms = 0, conv3_flops = 0;
  int conv_n = 0, other_n = 0, conv3_n = 0;
  for (size_t ci = 0; ci < hg->timed_bc.size(); ++ci) {
    for (size_t oi = 0; oi < n_ops; ++oi) {
      float ms = 0.f;
      DF3D_CUDA(cudaEventSynchronize(hg->events[(ci * n_ops + oi) * 2 + 1]));
      DF3D_CUDA(cudaEventElapsedTime(&ms, hg->events[(ci * n_ops + oi) * 2], hg->events[(ci * n_ops + oi) * 2 + 1]));
      const Op& op = hg->ops[oi];
      if (op.variant > 0 && op.variant != hg->timed_stem_variant) continue;  // the stem variant that did not run
      if (op.kind == OP_CONV || op.kind == OP_CHAIN) {
        conv_ms += ms;
        conv_flops += op.flops_per_image * hg->timed_bc[ci];
        ++conv_n;
        if ((op.kind == OP_CONV ? op.conv.taps : op.chain.taps) == 9) {
          conv3_ms += ms;
          conv3_flops += op.flops_per_image * hg->timed_bc[ci];
          ++conv3_n;
        }
      } else {
        other_ms += ms;
        ++other_n;
      }
    }
  }
  out8[0] = conv_ms;
  out8[1] = conv_flops;
  out8[2] = conv_n;
  out8[3] = other_ms;
  out8[4] = other_n;
  out8[5] = conv3_ms;
  out8[6] = conv3_flops;
  out8[7] = conv3_n;
  return DF3D_OK;
}

extern "C" int df3d_hg_num_ops(const df3d_hg* hg) { return hg ? (int)hg->ops.size() : 0; }

// per-op device time (ms, summed over the chunks of the last timed forward) + a one-line description
extern "C" int df3d_hg_op_timing(df3d_hg* hg, int op_index, double* ms_out, double* flops_out, double* bytes_out,
                                 char* desc, int desc_len) {
  DF3D_REQUIRE(hg && ms_out && flops_out && bytes_out && desc, DF3D_EINVAL, "df3d_hg_op_timing: null pointer");
  DF3D_REQUIRE(hg->timing && !hg->timed_bc.empty(), DF3D_EINVAL, "df3d_hg_op_timing: no timed forward recorded");
  const size_t n_ops = hg->ops.size();
  DF3D_REQUIRE(op_index >= 0 && (size_t)op_index < n_ops, DF3D_EINVAL, "df3d_hg_op_timing: bad op index");
  const Op& op = hg->ops[op_index];
  double ms = 0, images = 0;
  for (size_t ci = 0; ci < hg->timed_bc.size(); ++ci) {
    float t = 0.f;
    DF3D_CUDA(cudaEventSynchronize(hg->events[(ci * n_ops + op_index) * 2 + 1]));
    DF3D_CUDA(cudaEventElapsedTime(&t, hg->events[(ci * n_ops + op_index) * 2], hg->events[(ci * n_ops + op_index) * 2 + 1]));
    ms += t;
    images += hg->timed_bc[ci];
  }
  *ms_out = ms;
  *flops_out = op.flops_per_image * images;
  double bpi = 0;  // algorithmic HBM bytes per image of this op (activations in + out)
  if (op.kind == OP_CONV) {
    const ConvParams& p = op.conv;
    const double px = (double)p.H * p.W;
    bpi = px * 2.0 * (p.kc_per_tap * 64 + (p.residual ? p.res_ld : 0) + (p.has_res2 ? p.res_ld / 4.0 : 0) +
                      (p.out_raw ? p.raw_ld : 0) + (p.out_act ? p.act_ld : 0)) +
          px * 4.0 * (p.out_f32 ? p.f32_ld : 0);
    snprintf(desc, desc_len, "conv%dx%d %4dx%-4d cin=%3d BN=%3d%s%s%s", p.taps == 9 ? 3 : 1, p.taps == 9 ? 3 : 1, p.H, p.W,
             p.kc_per_tap * 64, op.BN, p.residual ? (p.has_res2 ? " +res+up" : " +res") : "", p.pool2 ? " +pool" : (p.out_act ? " +act" : ""),
             p.out_f32 ? " f32" : "");
  } else if (op.kind == OP_CHAIN) {
    const ChainParams& p = op.chain;
    bpi = op.chain_bytes_per_image;
    char shape[64];
    int o = snprintf(shape, sizeof(shape), "%d", p.kc_per_tap * 64);
    bool res = false, up = false;
    for (int i = 0; i < p.n_chain; ++i) {
      o += snprintf(shape + o, sizeof(shape) - o, ">%d", p.st[i].n);
      res |= p.st[i].has_res != 0;
      up |= p.st[i].has_res2 != 0;
    }
    snprintf(desc, desc_len, "chain%dx%d %4dx%-4d %s%s", p.taps == 9 ? 3 : 1, p.taps == 9 ? 3 : 1, p.H, p.W, shape,
             up ? " +res+up" : (res ? " +res" : ""));
  } else if (op.kind == OP_POOL) {
    bpi = (double)op.H * op.W * op.C * 2.0 * 1.5;
    snprintf(desc, desc_len, "maxpool+bn  %4dx%-4d c=%3d", op.H, op.W, op.C);
  } else if (op.kind == OP_IM2COL_GRAY) {
    bpi = (double)hg->desc.in_h * hg->desc.in_w + (double)hg->desc.in_h * hg->desc.in_w / 4 * kStemKGray * 2.0;
    snprintf(desc, desc_len, "stem im2col (gray)");
  } else if (op.kind == OP_IM2COL) {
    bpi = (double)hg->desc.in_h * hg->desc.in_w + (double)hg->desc.in_h * hg->desc.in_w / 4 * kStemKPadCols * 2.0;
    snprintf(desc, desc_len, "stem im2col");
  } else {
    bpi = (double)op.H * op.W * op.C * 4.0;
    snprintf(desc, desc_len, "argmax      %4dx%-4d", op.H, op.W);
  }
  *bytes_out = bpi * images;
  return DF3D_OK;
}

extern "C" int df3d_hg_set_mean(df3d_hg* hg, float m0, float m1, float m2) {
  DF3D_REQUIRE(hg, DF3D_EINVAL, "df3d_hg_set_mean: null handle");
  hg->mean[0] = m0;
  hg->mean[1] = m1;
  hg->mean[2] = m2;
  return DF3D_OK;
}
