/*
 * df3d_b200.h -- C ABI of libdf3d_b200.so, the B200-native (sm_100a) replacement for the
 * 2D->3D pose hot path of NeLy-EPFL/DeepFly3D.
 *
 * The reference has no native/FFI boundary: the seam is the Python surface of the two
 * un-vendored packages it imports (df3d/core.py:11-12).  Each entry point below names the
 * reference call site it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; every `*_dev` pointer is DEVICE memory owned by the caller
 *     (PyTorch allocator, cudaMalloc, ...).  The library never allocates or frees on the hot
 *     path; opaque handles own only their packed weights / launch plans / TMA descriptors.
 *   - every call is asynchronous w.r.t. the host: work is enqueued on `stream` (a cudaStream_t
 *     passed as void*), no implicit synchronisation.
 *   - return value: DF3D_OK (0) or a negative DF3D_E* code; the message of the last failure on
 *     the calling thread is returned by df3d_last_error().  No exceptions, no exit().
 *   - one handle must not be used from two streams concurrently; distinct handles are
 *     independent.  All functions are re-entrant across handles.
 */
#ifndef DF3D_B200_H_
#define DF3D_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DF3D_ABI_VERSION 2

enum {
  DF3D_OK = 0,
  DF3D_EINVAL = -1,     /* bad argument (null pointer, size out of range, misaligned) */
  DF3D_ECUDA = -2,      /* a CUDA runtime / driver call or kernel launch failed         */
  DF3D_ENOMEM = -3,     /* caller-provided workspace too small                          */
  DF3D_EUNSUPPORTED = -4/* shape not supported by the sm_100a kernels                   */
};

#define DF3D_MAX_CAMS 8   /* reference uses 7 (df3d/config.py:17) */

int df3d_abi_version(void);
const char* df3d_last_error(void);

/* --------------------------------------------------------------------------------------------
 * Heat-map decode.  Replaces the per-channel read-out inside df2d.inference.inference_folder
 * (call site df3d/core.py:177-185; rule documented at README.md:404: argmax_{h,w} H and H[h,w]
 * as confidence).  First occurrence wins on ties (row-major flat index), like numpy/torch argmax.
 *
 *   hm_dev   : (B,K,H,W) contiguous, NCHW.  dtype 0 = float32, 1 = bfloat16
 *   idx_dev  : (B,K) int32   flat index row*W + col
 *   conf_dev : (B,K) float32 peak value
 * ------------------------------------------------------------------------------------------ */
int df3d_heatmap_argmax(const void* hm_dev, int dtype, int B, int K, int H, int W,
                        int32_t* idx_dev, float* conf_dev, void* stream);

/* Same decode for the channels-last score tensor the hourglass kernels write:
 *   hm_dev (B,H,W,Cpad) float32, only channels [0,K) are decoded. */
int df3d_heatmap_argmax_nhwc(const float* hm_dev, int B, int H, int W, int Cpad, int K,
                             int32_t* idx_dev, float* conf_dev, void* stream);

/* --------------------------------------------------------------------------------------------
 * Image ingest: bilinear resize of uint8 gray images to the network input on the device.
 * Replaces the cv2.resize(img, (Wd, Hd), interpolation=INTER_LINEAR) of the image loader in front of
 * df2d.inference.inference_folder (call site df3d/core.py:177-185); bit-identical to OpenCV's
 * fixed-point algorithm for 8-bit images.
 *
 *   src_dev : (B, Hs, Ws) uint8, images contiguous
 *   dst_dev : (B, Hd, Wd) uint8; Wd a multiple of 4, 4-byte aligned
 * ------------------------------------------------------------------------------------------ */
int df3d_resize_gray_u8(const uint8_t* src_dev, int B, int Hs, int Ws, uint8_t* dst_dev, int Hd, int Wd,
                        void* stream);

/* --------------------------------------------------------------------------------------------
 * JPEG decode on the device (luminance), bound to nvJPEG at run time (dlopen; the library itself does not
 * depend on it).  Opt-in alternative to reading camera_C_img_I.jpg on the host (the reference's path behind
 * df2d.inference.inference_folder, df3d/core.py:177-185; frames come from core.py:446-459).  nvJPEG's inverse
 * DCT differs from libjpeg's by a few grey levels, so results are NOT bit-identical to the host read.
 *
 *   data, lens : HOST arrays of n compressed streams        dst_dev : (n, H, W) uint8 on the device
 *
 * df3d_jpeg_create probes the GPU's hardware JPEG engines first (nvJPEG backend 3, the frame block decoded with
 * nvjpegDecodeBatched) and settles for nvJPEG's default backend (0: one nvjpegDecode per image, Huffman stage on
 * the calling host thread) where the device or driver exposes none; df3d_jpeg_create_backend asks for one of
 * those or for backend 2 (GPU-hybrid: batched, Huffman stage on the SMs; DF3D_EUNSUPPORTED if it is not there); df3d_jpeg_backend reports the backend in use.
 * ------------------------------------------------------------------------------------------ */
typedef struct df3d_jpeg df3d_jpeg;
int df3d_jpeg_create(df3d_jpeg** out);
int df3d_jpeg_create_backend(df3d_jpeg** out, int backend);
int df3d_jpeg_backend(const df3d_jpeg* j);
void df3d_jpeg_destroy(df3d_jpeg* j);
int df3d_jpeg_info(df3d_jpeg* j, const uint8_t* data, size_t len, int* width, int* height);
int df3d_jpeg_decode_gray(df3d_jpeg* j, const uint8_t* const* data, const size_t* lens, int n, uint8_t* dst_dev,
                          int H, int W, void* stream);

/* --------------------------------------------------------------------------------------------
 * 19 -> 38 joint packing.  Replaces df3d/core.py:187-203 (Core.pose2d_estimation after
 * inference_folder) and the pixel scaling of core.py:247 (`points2d * image_shape[::-1]`).
 *
 *   idx_dev        : (C*T, K) int32 flat arg-max indices, image b = c*T + t   (camera-major)
 *   camera_ordering: HOST array of C ints (core.py:65-71)
 *   points2d_dev   : (C,T,2K,2) float64 normalised (row/Hh, col/Wh) with the reference's
 *                    blanking and un-flip quirks ((0,1) for unseen joints of cameras 4..6)
 *   pts_xy_dev     : (C,T,2K,2) float64 pixel (x, y) = (col*img_w, row*img_h); may be NULL
 * Requires C == 7 (the reference's slicing [:3], [4:], [2], [4] is hard-wired to 7 cameras).
 * ------------------------------------------------------------------------------------------ */
int df3d_pack_points2d(const int32_t* idx_dev, int C, int T, int K, int Hh, int Wh,
                       const int* camera_ordering, int img_w, int img_h,
                       double* points2d_dev, double* pts_xy_dev, void* stream);

/* --------------------------------------------------------------------------------------------
 * Multi-view DLT triangulation.  Replaces pyba CameraNetwork.triangulate()
 * (call site df3d/core.py:355; also used to initialise bundle adjustment).
 *
 *   P_dev      : (C,3,4) float64 projection matrices intr @ [R|t]
 *   pts_xy_dev : (C,T,J,2) float64 pixel (x,y); an observation is used iff x != 0 and y != 0
 *   pts3d_dev  : (T,J,3) float64; joints seen by < 2 cameras are written as 0
 * ------------------------------------------------------------------------------------------ */
int df3d_triangulate_dlt(const double* P_dev, const double* pts_xy_dev, int C, int T, int J,
                         double* pts3d_dev, void* stream);

/* P = intr @ [R(rvec) | tvec] for C cameras (device helper used between BA and DLT).
 *   cam_rt_dev (C,6) rvec,tvec ; intr_dev (C,4) fx,fy,cx,cy ; P_dev (C,3,4) ; R_dev (C,3,3) or NULL */
int df3d_projection_matrices(const double* cam_rt_dev, const double* intr_dev, int C,
                             double* P_dev, double* R_dev, void* stream);

/* --------------------------------------------------------------------------------------------
 * Bundle adjustment of the C camera extrinsics + all 3-D points.  Replaces pyba
 * CameraNetwork.bundle_adjust(update_intrinsic=False, update_distort=False)
 * (call site df3d/core.py:249), i.e. scipy.optimize.least_squares(method='trf', x_scale='jac',
 * ftol=1e-4) on the re-projection residuals.  The gauge is free, so the result depends on the
 * step rule: the kernels follow SciPy's trf_no_bounds (tr_solver='lsmr') iteration -- Jacobi
 * column scaling with a running maximum, Cauchy-step regularisation, 2-D subspace trust-region
 * step, ratio test / radius update / ftol-xtol-gtol termination -- with the regularised
 * Gauss-Newton step solved exactly through the Schur complement on the 6C camera unknowns
 * (analytic Jacobian, fp64).  Cameras without observations are returned unchanged.
 * ------------------------------------------------------------------------------------------ */
typedef struct df3d_ba_opts {
  int max_iters;      /* candidate steps evaluated (SciPy: nfev - 1); the reference needs 3-4 */
  double ftol;        /* stop when dF < ftol * F and ratio > 0.25  (reference: 1e-4)          */
  double xtol;        /* stop when |dx| < xtol * (xtol + |x|)      (SciPy default 1e-8)       */
  double gtol;        /* stop when |g|_inf < gtol                  (SciPy default 1e-8)       */
  int solver;         /* regularised Gauss-Newton step: 1 = LSMR with SciPy's tolerances and stopping rules
                         (reproduces the truncated iterate SciPy stops at; default), 0 = exact Schur-complement
                         solve (what LSMR converges to; 1-5e-5 mm away on the 3-D joints)        */
} df3d_ba_opts;

typedef struct df3d_ba_report {   /* written to DEVICE memory (no host sync) */
  double cost0;       /* 0.5 * sum r^2 at entry */
  double cost;        /* at exit                */
  double reg;         /* last regularisation of the Gauss-Newton step (Jacobi-scaled space) */
  int32_t iters;      /* candidate steps evaluated */
  int32_t accepted;   /* accepted steps            */
  int32_t n_obs;      /* observations used         */
  int32_t status;     /* SciPy's codes: 1 gtol, 2 ftol, 3 xtol, 4 ftol and xtol, 0 max_iters */
  int32_t lsmr_itn;   /* solver 1: iterations and stop code of the last LSMR solve              */
  int32_t lsmr_istop;
} df3d_ba_report;

size_t df3d_bundle_adjust_workspace_bytes(int C, int T, int J);

/*   cam_rt_dev : (C,6) rvec(3),tvec(3) in/out
 *   intr_dev   : (C,4) fx,fy,cx,cy
 *   pts_xy_dev : (C,T,J,2) pixel (x,y), visibility rule as in df3d_triangulate_dlt
 *   pts3d_dev  : (T,J,3) in: initial points (DLT with the initial cameras), out: BA points
 *   report_dev : df3d_ba_report in device memory (may be NULL)
 * Fixed launch sequence (2 + 6 * max_iters kernels; solver 1: 2 + 5 * max_iters, one of them a persistent
 * cooperative kernel; no host synchronisation); every reduction runs in a
 * fixed order, so the result is bit-reproducible -- frame-sharded multi-GPU runs all-gather the 2-D points and
 * run this solver replicated, every rank ends with identical cameras. */
int df3d_bundle_adjust(double* cam_rt_dev, const double* intr_dev, const double* pts_xy_dev,
                       int C, int T, int J, const df3d_ba_opts* opts, double* pts3d_dev,
                       df3d_ba_report* report_dev, void* workspace_dev, size_t workspace_bytes,
                       void* stream);
/* kernels one df3d_bundle_adjust call launches (for bench.py's gpu_launches) */
int df3d_bundle_adjust_launches(const df3d_ba_opts* opts);

/* Frame-sharded form of df3d_bundle_adjust for solver 0 (one process per GPU, the bundle adjustment is
 * ONE problem over the frames of all ranks).  Every rank holds the same cameras, 2-D points and workspace layout; the
 * per-point work of each of the four passes of an iteration (0 gradient, 1 Schur system, 2 back-substitution, 3 candidate
 * step) is split by blocks of points: rank r runs blocks [r n_blocks / world, (r+1) n_blocks / world) and writes their
 * partial sums; the CALLER all-gathers the partials (torch.distributed / NCCL -- the library holds no communicator):
 * region [partials_offset, + n_blocks * pass_doubles[pass] * 8) of the workspace, rank r's slice being its blocks;
 * df3d_ba_sharded_finish then sums ALL blocks in the fixed order the single-GPU solver uses.  Same blocks, same order:
 * the cameras are bit-identical to df3d_bundle_adjust on one GPU, on every rank.  pts3d_dev: only the points of the
 * rank's own blocks are updated.  Sequence: begin; per iteration 1..max_iters: for pass 0..3 { pass; all-gather;
 * finish(pass, iteration) }; end.  df3d_ba_sharded_plan fails with DF3D_EUNSUPPORTED when the blocks do not split
 * evenly over `world` (run df3d_bundle_adjust replicated then). */
int df3d_ba_sharded_plan(int C, int T, int J, int world, int* n_blocks, size_t* partials_offset, int* pass_doubles /* [4] */);
int df3d_ba_sharded_begin(const double* cam_rt_dev, int C, int T, int J, const df3d_ba_opts* opts, void* workspace_dev,
                          size_t workspace_bytes, void* stream);
int df3d_ba_sharded_pass(int pass, int rank, int world, const double* intr_dev, const double* pts_xy_dev,
                         const double* pts3d_dev, int C, int T, int J, void* workspace_dev, size_t workspace_bytes, void* stream);
int df3d_ba_sharded_finish(int pass, int iter, int rank, int world, double* pts3d_dev, int C, int T, int J, void* workspace_dev,
                           size_t workspace_bytes, void* stream);
int df3d_ba_sharded_end(double* cam_rt_dev, int C, int T, int J, df3d_ba_report* report_dev, void* workspace_dev,
                        size_t workspace_bytes, void* stream);

/* Mean L2 reprojection error in pixels (pyba CameraNetwork.reprojection_error(), printed at
 * df3d/core.py:250).  out_dev: 2 float64 = {sum of distances, number of observations}. */
int df3d_reprojection_error(const double* cam_rt_dev, const double* intr_dev,
                            const double* pts_xy_dev, const double* pts3d_dev, int C, int T, int J,
                            double* out_dev, void* stream);

/* --------------------------------------------------------------------------------------------
 * Procrustes registration of the triangulated skeleton to the template pose, on the device.  Replaces
 * df3d.procrustes.procrustes_seperate (df3d/procrustes.py:51-263, df3d/plot_util.py:85-91), run by Core.save
 * (df3d/core.py:358) and Core.get_points3d (core.py:339).  Joints 0-18 and 19-37 are registered separately:
 * scale = median over 12 bones of (template median length / median length), subtract the median of all points,
 * orthogonal fit (reflection allowed, no scaling) of the median BODY_COXA / COXA_FEMUR joints.  The medians over
 * all frames are radix selects on the device (numpy's even-count rule).
 *
 *   pts3d_dev            : (T,38,3) float64          out_dev : (T,38,3) float64 (may alias pts3d_dev: no)
 *   template_medians_dev : 2 x 30 float64, per half: 12 median bone lengths of the template, then the median
 *                          coordinates (6 joints x 3) of its alignment joints (a constant of the template)
 * ------------------------------------------------------------------------------------------ */
size_t df3d_procrustes_workspace_bytes(int T);
int df3d_procrustes(const double* pts3d_dev, int T, int J, const double* template_medians_dev, double* out_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------------------------------
 * Temporal filters of the pose tracks.  Replace df3d.signal_util.filter_batch / filter_batch_2d (One-Euro filter,
 * df3d/signal_util.py:31-132; used by Core.get_points3d, df3d/core.py:342) and smooth_pose2d
 * (signal_util.py:135-160; Core.smooth_points2d, core.py:286-296).
 *
 *   pts_dev / out_dev : (T, n_tracks) float64, frame-major (n_tracks = joints x coordinates)
 *   t_first           : 1 = time stamps (i+1)*0.1 (filter_batch), 0 = i*0.1 (filter_batch_2d).  As in the
 *                       reference, the sampling frequency is re-derived from consecutive time stamps.
 * The One-Euro recurrence is evaluated with the reference's operation order and no fused multiply-add:
 * bit-identical to the Python implementation.
 * ------------------------------------------------------------------------------------------ */
int df3d_one_euro_filter(const double* pts_dev, int T, int n_tracks, double freq, double mincutoff, double beta,
                         double dcutoff, int t_first, double* out_dev, void* stream);
int df3d_smooth_pose2d(const double* pts_dev, int T, int n_tracks, int window_size, double std_thr, double* out_dev,
                       void* stream);

/* --------------------------------------------------------------------------------------------
 * Stacked-hourglass forward + decode.  Replaces the network forward inside
 * df2d.inference.inference_folder (call site df3d/core.py:177-185; hyper-parameters hinted at
 * df3d/config.py:18,33-36).  Convolutions run as tcgen05 implicit-GEMM tiles fed by TMA.
 * ------------------------------------------------------------------------------------------ */
typedef struct df3d_hg_desc {
  int num_stacks;     /* 2 (reference config) or 8 (sh8 weights)          */
  int num_classes;    /* 19                                               */
  int in_h, in_w;     /* network input size, multiple of 64 (256x512 ref) */
  int max_batch;      /* largest B passed to df3d_hg_forward_*            */
} df3d_hg_desc;

typedef struct df3d_hg df3d_hg;

/* Number of float32 values in the flat parameter blob expected by df3d_hg_create; the layout is
 * the module order of the oracle model (see deepfly3d_b200/hourglass.py: flatten_params). */
size_t df3d_hg_param_count(const df3d_hg_desc* desc);
size_t df3d_hg_workspace_bytes(const df3d_hg_desc* desc);

/* params_host: HOST float32 blob (conv weights OIHW, biases, BN gamma/beta/mean/var).
 * BN is folded, weights are packed to bf16 K-major tiles and uploaded once. */
int df3d_hg_create(const df3d_hg_desc* desc, const float* params_host, size_t n_params,
                   df3d_hg** out);
void df3d_hg_destroy(df3d_hg* hg);

/*   images_dev : dtype 0: (B,H,W) uint8 gray, replicated to 3 channels, x/255 - mean
 *                dtype 1: (B,3,H,W) float32 already normalised (what the torch model takes)
 *   flip_dev   : (B) uint8, 1 = mirror the image left-right before the network
 *                (camera_ids_to_flip, df3d/core.py:179); may be NULL
 *   idx_dev    : (B,K) int32, conf_dev : (B,K) float32 (decode of the LAST stack)
 *   heatmap_dev: optional (B,Hh,Wh,32) float32 copy of the last stack's scores, or NULL      */
int df3d_hg_forward_argmax(df3d_hg* hg, const void* images_dev, int dtype, const uint8_t* flip_dev,
                           int B, int32_t* idx_dev, float* conf_dev, float* heatmap_dev,
                           void* workspace_dev, size_t workspace_bytes, void* stream);

/* per-channel mean subtracted from x/255 for dtype 0 input (default 0.5, 0.5, 0.5) */
int df3d_hg_set_mean(df3d_hg* hg, float m0, float m1, float m2);

/* Single convolution layer through the same tcgen05 kernel (operator-level entry point; used by
 * the parity tests to check every layer shape against a float32 reference).  Stride 1, "same"
 * padding.  Packs the weights on every call (cudaMalloc + synchronous copy): NOT a hot-path call.
 *   in_dev (B,H,W,Cin) bf16 NHWC, Cin % 64 == 0;  w_host (Cout,Cin,k,k) float32, k in {1,3};
 *   Cout in {64,128,256};  v = conv*scale1[c] + shift1[c] (+ residual) ; relu1 ;
 *   out_dev = bf16(v) ; out_act_dev = bf16(relu(bf16(v)*scale2[c] + shift2[c])) (optional). */
int df3d_conv2d_nhwc_bf16(const void* in_dev, int B, int H, int W, int Cin, const float* w_host, int Cout,
                          int ksize, const float* scale1_host, const float* shift1_host, int relu1,
                          const void* residual_dev, void* out_dev, const float* scale2_host,
                          const float* shift2_host, void* out_act_dev, void* stream);

/* Optional per-launch timing for bench.py / profiling (NOT for production runs: it creates CUDA
 * events).  When enabled, every kernel of df3d_hg_forward_argmax is bracketed by events on the
 * launching stream; df3d_hg_read_timing synchronises them and returns, for the last forward,
 *   out8 = { conv_gemm ms, conv_gemm algorithmic FLOP, conv_gemm launches,
 *            other kernels ms, other launches, 3x3-conv ms, 3x3-conv FLOP, 3x3-conv launches }. */
int df3d_hg_set_timing(df3d_hg* hg, int enable);
int df3d_hg_read_timing(df3d_hg* hg, double* out8);
/* per-op view of the same timing: device ms, algorithmic FLOP and activation bytes of plan entry
 * `op_index` (0 <= op_index < df3d_hg_num_ops) summed over the last timed forward, plus a label */
int df3d_hg_num_ops(const df3d_hg* hg);
int df3d_hg_op_timing(df3d_hg* hg, int op_index, double* ms_out, double* flops_out, double* bytes_out,
                      char* desc, int desc_len);

/* number of kernels one df3d_hg_forward_argmax call launches (for bench.py's gpu_launches) */
int df3d_hg_launches_per_forward(const df3d_hg* hg, int B);

#ifdef __cplusplus
}
#endif
#endif /* DF3D_B200_H_ */
