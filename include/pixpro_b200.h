/*
 * pixpro_b200.h — C ABI of the B200-native pixel-pretext hot path of PixPro-with-OpticalFlow.
 *
 * One shared library (libpixpro_b200.so, built from pixpro-with-opticalflow_b200/csrc/ for
 * sm_100a).  Plain pointers and sizes only: no torch types cross this boundary.  All
 * pointers are DEVICE pointers unless the function name ends in `_host`.  Every launch is
 * asynchronous on `stream` (a cudaStream_t passed as void*; 0 = legacy default stream) and
 * never synchronises.  Tensors are dense row-major fp32 unless stated; masks are one byte
 * per element (0/1), bit-compatible with torch.bool.
 *
 * Each entry point replaces one function of the reference (paths relative to the
 * reference repository root); the citation is on the declaration.  Return value: 0 on
 * success, non-zero PP_ERR_* otherwise (pp_last_error() gives the message for the calling
 * thread).  There is no CPU fallback: without a CUDA device every compute entry fails.
 *
 * div_mode selects how `tensor / python_scalar` sites of the reference round:
 *   PP_DIV_IEEE (0)  true IEEE division  — what the reference computes on CPU (pinned oracle)
 *   PP_DIV_RCP  (1)  x * fl32(1/s)       — what torch's CUDA true-divide kernel computes
 */
#ifndef PIXPRO_B200_H
#define PIXPRO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* exported symbol (the library is built with -fvisibility=hidden) */
#if defined(__GNUC__)
#define PP_API __attribute__((visibility("default")))
#else
#define PP_API
#endif

#define PP_OK 0
#define PP_ERR_INVALID 1   /* bad argument (shape / null pointer / unsupported size) */
#define PP_ERR_CUDA 2      /* CUDA runtime error at launch */

#define PP_DIV_IEEE 0
#define PP_DIV_RCP 1

#define PP_NORM_COORD 0
#define PP_NORM_FLOW 1
#define PP_DENORM_FLOW 2

/* library identification / diagnostics */
PP_API int pp_abi_version(void);
PP_API const char* pp_last_error(void);
/* number of kernels launched by this library in the calling process since load */
PP_API int64_t pp_launch_count(void);
/* Diagnostics of the TMA-staged FB-mask kernel: number of pixels (since load / last reset) whose
 * gather footprint fell outside the staged box and were recomputed from global memory.
 * Synchronises the device.  -1 on error. */
PP_API int64_t pp_fb_redo_count(int reset);
/* Diagnostics of the fused x8-up-sampling chain kernel (pp_flow_stage, n > 1, flow_up): number of pixel-links
 * (since load / last reset) whose tap footprint fell outside the staged region and were evaluated from the
 * low-res link directly (same values, slow path).  Synchronises the device.  -1 on error. */
PP_API int64_t pp_chain_slow_count(int reset);
/* Per-kernel device timing (tracing aid).  pp_profile_enable(1) clears the records and makes
 * every launch bracket its kernel with a cudaEvent pair on the launch stream;
 * pp_profile_num_kernels() synchronises those events and aggregates by kernel name;
 * pp_profile_get(i, ...) returns name, launch count and total device milliseconds. */
PP_API int pp_profile_enable(int on);
PP_API int pp_profile_num_kernels(void);
PP_API int pp_profile_get(int idx, char* name, int name_cap, int64_t* launches, double* total_ms);

/* ---- a1: upflow8 — contrast/flow/utils/utils.py:87-89 -----------------------------------
 * in [planes,h,w] -> out [planes,8h,8w] = 8 * bilinear(align_corners=True).               */
PP_API int pp_upflow8(const float* in, int64_t planes, int h, int w, float* out, void* stream);

/* ---- a2: normalize_coord / normalize_flow / denormalize_flow — contrast/util.py:334-357 -
 * x,out [B,2,H,W]; kind = PP_NORM_COORD | PP_NORM_FLOW | PP_DENORM_FLOW.                   */
PP_API int pp_normalize(const float* x, int64_t B, int H, int W, int kind, int div_mode, float* out, void* stream);

/* ---- a3: concat_flow — contrast/util.py:301-330 -----------------------------------------
 * Chains n dense flow links.  Link i of sample b is the [2,H,W] block at
 * flows + i*stride_n + b*stride_b (elements), so [n,B,2,H,W] and [B,n,2,H,W] both work.
 * out [B,2,H,W].  n==1 copies (normalises if is_norm), as the reference does.             */
PP_API int pp_concat_flow(const float* flows, int n, int64_t B, int H, int W, int64_t stride_n, int64_t stride_b,
                   int is_norm, int div_mode, float* out, void* stream);

/* ---- a5: forward_backward_consistency — contrast/util.py:253-297 ------------------------
 * fwd,bwd [B,2,H,W] -> mask u8 [B,H,W]; optional cycle [B,2,H,W] and coords1_norm
 * [B,2,H,W] (NULL to skip).  is_norm: inputs are already normalised (util.py:258-262).    */
PP_API int pp_fb_consistency(const float* fwd, const float* bwd, int64_t B, int H, int W, double alpha_1, double alpha_2,
                      int is_norm, int div_mode, uint8_t* mask, float* cycle, float* coords1_norm, void* stream);

/* ---- a5 x2: both FB masks of apply_optical_flow — contrast/util.py:211-213 ----------------
 * (forward_backward_consistency(fwd, bwd) and (bwd, fwd)) in one launch: fwd,bwd [B,2,H,W] ->
 * mask_fwd, mask_bwd u8 [B,H,W].  The mask half of pp_flow_stage as its own entry, for callers
 * that schedule it on another stream than the chain (sample chunks).                          */
PP_API int pp_fb_masks(const float* fwd, const float* bwd, int64_t B, int H, int W, double alpha_1, double alpha_2,
                int is_norm, int div_mode, uint8_t* mask_fwd, uint8_t* mask_bwd, void* stream);

/* ---- a6 (a1+a3+a5 fused): flow stage of apply_optical_flow — contrast/util.py:175-248 ---
 * (use_flow_file, not use_flow_frames).  lo_fwd/lo_bwd in the loader layout [B,n,2,h,w]
 * (contrast/data/dataset.py:485-495).  flow_up: links are x8-upsampled on the fly, fused into the
 * chain kernels for every n (csrc/pp_chainup.cuh for n > 1): no device scratch is needed any more —
 * pp_flow_stage_workspace() returns 0 and `workspace` may be NULL (both kept for ABI stability).
 * Outputs flow_fwd/flow_bwd [B,2,H,W]
 * (H,W = 8h,8w if flow_up), and, when use_mask, mask_fwd/mask_bwd u8 [B,H,W].  is_norm restates
 * --flow_cat_norm.                                                                          */
PP_API int64_t pp_flow_stage_workspace(int64_t B, int n, int h, int w, int flow_up);
PP_API int pp_flow_stage(const float* lo_fwd, const float* lo_bwd, int64_t B, int n, int h, int w, int flow_up, int use_mask,
                         double alpha_1, double alpha_2, int is_norm, int div_mode, float* flow_fwd, float* flow_bwd,
                         uint8_t* mask_fwd, uint8_t* mask_bwd, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- a11: calc_mask_ratio — contrast/util.py:361-366 ------------------------------------
 * mask u8 [B,H,W] -> ratio [B] = fraction of zero entries.                                */
PP_API int pp_calc_mask_ratio(const uint8_t* mask, int64_t B, int H, int W, float* ratio, void* stream);

/* ---- a7: add_optical_flow — contrast/models/PixPro.py:46-89 -----------------------------
 * flow [B,2,Hin,Win]; x_grid,y_grid [B,P] (pixels of the H_orig x W_orig frame); mask u8
 * [B,Hin,Win] or NULL -> out_x,out_y [B,P]; mask_grid u8 [B,P] (NULL to skip).            */
PP_API int pp_add_optical_flow(const float* flow, int64_t B, int Hin, int Win, const float* x_grid, const float* y_grid, int P,
                        int H_orig, int W_orig, const uint8_t* mask, int div_mode, float* out_x, float* out_y,
                        uint8_t* mask_grid, void* stream);

/* ---- a8: regression_loss forward (+ gradient) — contrast/models/PixPro.py:92-247 --------
 * q,k [B,C,P] with P=G*G; coord_q,coord_k [B,10]; flow [B,2,Hin,Win] or NULL (no-flow path,
 * PixPro.py:167-175); mask u8 [B,Hin,Win] or NULL.
 * Outputs: loss[1]; pos_num[B]; pos_mean[B]; dq [B,C,P] = d loss/d q (the backward of the
 * reference's autograd graph for upstream gradient 1; NULL to skip); optional debug/parity
 * outputs pos_mask u8 [B,P,P] and centres [4,B,P] = (warped q x, warped q y, k x, k y).
 * workspace: pp_regression_loss_workspace(B,C,G) bytes of device scratch.                 */
PP_API int64_t pp_regression_loss_workspace(int64_t B, int C, int G);
PP_API int pp_regression_loss(const float* q, const float* k, int64_t B, int C, int G, const float* coord_q, const float* coord_k,
                       const float* flow, int Hin, int Win, const uint8_t* mask, int H_orig, int W_orig, double pos_ratio,
                       int div_mode, float* loss, float* pos_num, float* pos_mean, float* dq, uint8_t* pos_mask,
                       float* centres, void* workspace, void* stream);

/* Both loss directions of PixPro.forward (contrast/models/PixPro.py:429-430) in ONE launch:
 * every argument that differs between the two regression_loss calls is a table of two
 * pointers (entries of flow[] / mask[] may be NULL).  Same results as two pp_regression_loss
 * calls; workspace[i] each pp_regression_loss_workspace(B,C,G) bytes.                       */
PP_API int pp_regression_loss_pair(const float* const* q, const float* const* k, int64_t B, int C, int G,
                                   const float* const* coord_q, const float* const* coord_k, const float* const* flow,
                                   int Hin, int Win, const uint8_t* const* mask, int H_orig, int W_orig, double pos_ratio,
                                   int div_mode, float* const* loss, float* const* pos_num, float* const* pos_mean,
                                   float* const* dq, void* const* workspace, void* stream);

/* ---- a6+a7 fused, sparse: flow-guided correspondence evaluated only at the grid centres ------
 * What regression_loss consumes of apply_optical_flow's dense outputs (contrast/util.py:175-248) is
 * add_optical_flow's P = G*G samples per image (contrast/models/PixPro.py:46-89).  This entry
 * evaluates exactly those: the composite flow (x8 up-sampling + chain, util.py:185-191,301-330) at
 * the 4 integer neighbours of every grid centre (centres from the crop descriptors,
 * PixPro.py:140-143,192-199) and the forward-backward test (util.py:253-297) at its nearest pixel,
 * with the arithmetic of the dense kernels — bit-identical to pp_flow_stage + pp_add_optical_flow,
 * without materialising [B,2,H,W] / [B,H,W].  lo_fwd/lo_bwd: loader layout [B,n,2,h,w] (full-res
 * links if !flow_up).  coord_fwd [B,10]: descriptors of the view warped by the FORWARD composite
 * (coord1 in PixPro.forward), coord_bwd: by the backward one; either pair (coord, warped) may be
 * NULL.  warped_* [3,B,P] = (out_x, out_y, mask_grid as 0/1; all ones when !use_mask).
 * Not supported: --flow_cat_norm (use pp_flow_stage).                                          */
PP_API int pp_sparse_corr(const float* lo_fwd, const float* lo_bwd, int64_t B, int n, int h, int w, int flow_up, int use_mask,
                          double alpha_1, double alpha_2, const float* coord_fwd, const float* coord_bwd, int G, int H_orig,
                          int W_orig, int div_mode, float* warped_fwd, float* warped_bwd, void* stream);

/* regression_loss on centres already warped by pp_sparse_corr (`warped` [3,B,P]) instead of a dense
 * flow / mask; everything else as pp_regression_loss / pp_regression_loss_pair.               */
PP_API int pp_regression_loss_warped(const float* q, const float* k, int64_t B, int C, int G, const float* coord_q,
                                     const float* coord_k, const float* warped, int H_orig, int W_orig, double pos_ratio,
                                     int div_mode, float* loss, float* pos_num, float* pos_mean, float* dq, uint8_t* pos_mask,
                                     float* centres, void* workspace, void* stream);
PP_API int pp_regression_loss_pair_warped(const float* const* q, const float* const* k, int64_t B, int C, int G,
                                          const float* const* coord_q, const float* const* coord_k,
                                          const float* const* warped, int H_orig, int W_orig, double pos_ratio, int div_mode,
                                          float* const* loss, float* const* pos_num, float* const* pos_mean,
                                          float* const* dq, void* const* workspace, void* stream);

/* ---- a9: PixPro.featprop (+ the caller's F.normalize) — contrast/models/PixPro.py:339-363,380
 * feat,val [B,C,P] (val = value_transform(feat), computed by the caller) -> out [B,C,P].
 * final_norm: also apply the L2 normalisation of PixPro.py:380.  saved: device scratch of
 * pp_ppm_saved_bytes(B,C,P) bytes that pp_ppm_bwd reads (kept by the autograd node).       */
PP_API int64_t pp_ppm_saved_bytes(int64_t B, int C, int P);
PP_API int pp_ppm_fwd(const float* feat, const float* val, int64_t B, int C, int P, double gamma, double clamp_value,
               int final_norm, float* out, void* saved, void* stream);
/* out = what pp_ppm_fwd returned; g [B,C,P] = dL/d out -> d_feat_sim (gradient reaching feat
 * through the similarity), d_val [B,C,P] (gradient w.r.t. val; the caller back-propagates it
 * through value_transform).  workspace: pp_ppm_bwd_workspace(B,C,P) bytes of device scratch. */
PP_API int64_t pp_ppm_bwd_workspace(int64_t B, int C, int P);
PP_API int pp_ppm_bwd(const float* feat, const float* val, const float* out, const float* g, const void* saved, int64_t B, int C,
               int P, double gamma, double clamp_value, int final_norm, float* d_feat_sim, float* d_val, void* workspace,
               void* stream);

/* ---- value transform of the PPM: 1x1 convolution (contrast/models/PixPro.py:21-23,300,343) ---
 * x [B,Cin,P], w [Cout,Cin], bias [Cout] or NULL -> y [B,Cout,P]; backward: dy -> dx (NULL to
 * skip), dw, db (NULL to skip).  One tcgen05 3xTF32 GEMM each: per sample through the TMA-fed kernel
 * (csrc/pp_tc2.cuh) over hi / lo planes written into the workspace when P >= 128 and Cin, Cout, P are multiples of 4,
 * else over the joint (sample,pixel) index.  workspace: pp_conv1x1_fwd_workspace() bytes for the forward (may be 0 -> NULL),
 * pp_conv1x1_bwd_workspace() bytes for the backward (operand planes, per-sample / split-K partials of dw).        */
PP_API int64_t pp_conv1x1_fwd_workspace(int64_t B, int Cin, int Cout, int P);
PP_API int pp_conv1x1_fwd(const float* x, const float* w, const float* bias, int64_t B, int Cin, int Cout, int P, float* y,
                          void* workspace, void* stream);
PP_API int64_t pp_conv1x1_bwd_workspace(int64_t B, int Cin, int Cout, int P);
PP_API int pp_conv1x1_bwd(const float* x, const float* w, const float* dy, int64_t B, int Cin, int Cout, int P, float* dx,
                          float* dw, float* db, void* workspace, void* stream);

/* ---- multi-tensor optimizer-side kernels ------------------------------
 * A parameter set = a DEVICE table of PpMtTensor entries + a DEVICE chunk map of int pairs
 * {tensor index, chunk index within the tensor}, pp_mt_chunk_elems() elements per chunk, the chunks
 * of one tensor contiguous and in order; first_chunk[t] .. first_chunk[t+1] are tensor t's chunks. */
typedef struct PpMtTensor {
    void* a;         /* EMA: online parameter q (read)        LARS/SGD: parameter p (read/write) */
    void* b;         /* EMA: momentum parameter k (read/write) LARS/SGD: gradient g (read)       */
    void* c;         /* LARS/SGD: momentum buffer (read/write), may be null when momentum == 0   */
    int64_t numel;
    float s0, s1, s2, s3; /* LARS/SGD: weight_decay, lr, momentum, dampening */
    int flags;       /* PP_MT_* */
    int pad_;
} PpMtTensor;
#define PP_MT_LARS 1        /* apply the adaptive rate (param group 'ignore' is False, lars.py:123) */
#define PP_MT_FIRST_STEP 2  /* momentum buffer not initialised yet: buf = grad (torch.optim.SGD) */
PP_API int pp_mt_chunk_elems(void);
/* EMA of the key branch — contrast/models/PixPro.py:322-337: k = k*m + q*(1-m), all tensors, one launch.
 * `one_minus_momentum` is passed separately because the reference forms 1-m in double precision. */
PP_API int pp_ema_update(const PpMtTensor* table_dev, const int* chunk_map_dev, int nchunks, double momentum,
                         double one_minus_momentum, void* stream);
/* LARS.step() over all parameters — contrast/lars.py:109-152 wrapping torch.optim.SGD (no nesterov):
 * weight decay folded into the gradient, per-tensor norms, adaptive rate, momentum, parameter update.
 * Three launches, no host synchronisation.  workspace: pp_lars_workspace(ntensors, nchunks) bytes. */
PP_API int64_t pp_lars_workspace(int ntensors, int nchunks);
PP_API int pp_lars_sgd_step(const PpMtTensor* table_dev, int ntensors, const int* chunk_map_dev, const int* first_chunk_dev,
                            int nchunks, double trust_coef, double eps, void* workspace, void* stream);

/* ---- tensor-core building block (tcgen05, 3xTF32: fp32-accurate) ----------------------------
 * C[b] = A[b] * B[b]^T;  A [batch,M,K], B [batch,N,K], C [batch,M,N], all row-major fp32.
 * The PPM / loss contractions at large grids run on the same kernel with fused loaders; this
 * entry exposes the bare GEMM for numerics tests and microbenchmarks.                        */
PP_API int pp_tc_gemm_nt(const float* A, const float* B, float* C, int64_t batch, int M, int N, int K, void* stream);
/* The same product through the TMA-fed, warp-specialised kernel (csrc/pp_tc2.cuh): operands are split once into hi / lo
 * planes in `workspace` (pp_tc_gemm_nt_workspace bytes, 16-byte aligned) and streamed by TMA in the 128-byte-swizzled UMMA
 * layout; 128 x 256 CTA tiles.  K % 4 != 0 or PIXPRO_B200_TC2=0 falls back to the kernel above.                  */
PP_API int64_t pp_tc_gemm_nt_workspace(int64_t batch, int M, int N, int K);
PP_API int pp_tc_gemm_nt_ws(const float* A, const float* B, float* C, int64_t batch, int M, int N, int K, void* workspace, void* stream);
/* General operand layouts through the same kernel: a_mn / b_mn != 0 = that operand is stored MN-major (A as [batch][K][M], B as
 * [batch][K][N]) and read in place by TMA boxes + MN-major UMMA descriptors; fp32 operands are split into hi / lo by the kernel's
 * converter warps (no workspace).  PP_ERR_INVALID when a shape is not streamable (no fallback).                     */
PP_API int pp_tc_gemm_ws(const float* A, const float* B, float* C, int64_t batch, int M, int N, int K, int a_mn, int b_mn, void* stream);
/* The same with padded operands and sums over batch entries.  a_pitch / b_pitch: allocated length in floats (a multiple of 4;
 * 0 = dense) of the operand's contiguous dimension (K when K-major, M resp. N when MN-major) — a [C, 49] map of the 7x7 grid
 * (contrast/models/PixPro.py:339-363 at the published crop size) copied to a pitch of 52 floats is streamable, the dense one is not
 * (TMA strides are multiples of 16 bytes).  kb > 1: C[g] = sum over the kb batch entries of group g of A[b] B[b]^T, C is
 * [ceil(batch / kb)][M][N] — how the value transform's weight gradient sums its per-sample products.              */
PP_API int pp_tc_gemm_ex(const float* A, const float* B, float* C, int64_t batch, int M, int N, int K, int a_mn, int b_mn, int a_pitch,
                  int b_pitch, int kb, void* stream);

/* ---- synchronised batch normalisation in three launches per direction (csrc/pp_syncbn.cu) ----------------------------
 * The reference converts every BatchNorm of encoder / projector to torch.nn.SyncBatchNorm (contrast/models/PixPro.py:289-292,
 * 315-317) and trains under DDP (main_pretrain.py:78); torch issues ~10 launches + one all_gather per layer and direction.
 * x, y, dy, dx: logical [N, C, HW] in `layout` (PP_LAYOUT_NCHW contiguous, or PP_LAYOUT_NHWC = torch.channels_last memory),
 * element type `dtype` (PP_DTYPE_F32 / PP_DTYPE_BF16); every statistic is fp32.  workspace: pp_bn_workspace(C) bytes, ZEROED
 * once by the caller (holds a completion ticket that every launch leaves at 0), one per stream in flight.
 *   pp_bn_stats     -> stats[0..C) = local mean, [C..2C) = local M2 = sum (x - mean)^2, [2C] = local count; the caller
 *                      all-gathers this ONE vector across ranks (nranks rows of 2C+1)
 *   pp_bn_apply     global mean / invstd from the gathered rows (parallel-variance combination); running = (1-momentum)
 *                   running + momentum stat (unbiased variance); y = (x - mean) invstd weight + bias; save_mean [C] and
 *                   save_invstd [C+1] (its last element = the total count) for the backward
 *   pp_bn_bwd_stats -> sums[0..C) = sum dy, [C..2C) = sum dy (x - mean) (the caller all-reduces); grad_weight / grad_bias [C]
 *                      from the LOCAL sums (DDP reduces parameter gradients itself), either may be NULL
 *   pp_bn_bwd_apply dx = (dy - mean(dy) - (x - mean) invstd^2 mean(dy (x - mean))) invstd weight; total_count: DEVICE pointer
 *                   to the total element count (= save_invstd[C] of the forward)                                           */
#define PP_LAYOUT_NCHW 0
#define PP_LAYOUT_NHWC 1
#define PP_DTYPE_F32 0
#define PP_DTYPE_BF16 1
PP_API int64_t pp_bn_workspace(int C);
PP_API int pp_bn_stats(const void* x, int64_t N, int C, int HW, int layout, int dtype, void* workspace, float* stats, void* stream);
PP_API int pp_bn_apply(const void* x, void* y, int64_t N, int C, int HW, int layout, int dtype, const float* stats, int nranks, const float* weight,
                       const float* bias, float* running_mean, float* running_var, double eps, double momentum, float* save_mean,
                       float* save_invstd, void* stream);
PP_API int pp_bn_bwd_stats(const void* dy, const void* x, int64_t N, int C, int HW, int layout, int dtype, const float* save_mean,
                           const float* save_invstd, void* workspace, float* sums, float* grad_weight, float* grad_bias, void* stream);
PP_API int pp_bn_bwd_apply(const void* dy, const void* x, void* dx, int64_t N, int C, int HW, int layout, int dtype, const float* save_mean,
                           const float* save_invstd, const float* weight, const float* sums, const float* total_count, void* stream);

/* ---- RAFT correlation volume, pyramid and lookup ----------------------------
 * The reference's torch CorrBlock (contrast/flow/corr.py:12-60; its CUDA twin `alt_cuda_corr` is not shipped).
 * pp_corr_volume : CorrBlock.corr, corr.py:52-60.  fmap1, fmap2 [B, D, h, w] -> corr [B, h*w, h*w]
 *                  (= the reference's [B,h,w,1,h,w]); <f1[:,i], f2[:,j]> / sqrt(D), tcgen05 3xTF32 for h*w >= 128.
 * pp_corr_pool   : one pyramid step, corr.py:26-28: avg_pool2d(2, stride 2) of `planes` planes [h,w] -> [h/2,w/2].
 * pp_corr_lookup : CorrBlock.__call__, corr.py:30-50 + bilinear_sampler (flow/utils/utils.py:64-78).
 *                  levels: HOST array of num_levels device pointers, level l = [B*h*w, h>>l, w>>l]; coords [B,2,h,w]
 *                  (x, y in level-0 pixels) -> out [B, num_levels*(2r+1)^2, h, w], channel = l*(2r+1)^2 + i*(2r+1) + j with
 *                  the reference's window convention (index i shifts x, j shifts y).  div_mode: PP_DIV_* of the
 *                  tensor / python-scalar divisions in bilinear_sampler.                                        */
PP_API int pp_corr_volume(const float* fmap1, const float* fmap2, int64_t B, int D, int h, int w, float* corr, void* stream);
PP_API int pp_corr_pool(const float* in, int64_t planes, int h, int w, float* out, void* stream);
PP_API int pp_corr_lookup(const float* const* levels, int num_levels, const float* coords, int64_t B, int h, int w, int radius,
                          int div_mode, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIXPRO_B200_H */
