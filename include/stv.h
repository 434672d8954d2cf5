/* stv.h — C ABI of libstv (slowtv_monodepth_b200/csrc), the B200 (sm_100a) kernels behind the SlowTV-monodepth
 * training hot path.
 *
 * The reference (jspenmar/slowtv_monodepth) has no native layer: its "FFI" for this path is a set of Python
 * nn.Module.forward / handler signatures resolved through src/registry.py. Each entry point below replaces the ATen op
 * chain launched by one of those call sites (cited per function, paths under /root/reference); the host-side Python
 * mirror in slowtv_monodepth_b200/ binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (unless typed otherwise), NCHW contiguous, owned by the caller;
 *   - the library never allocates, frees or synchronises; all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - `const float* const* x` arguments are HOST arrays of device pointers (one per scale);
 *   - return value 0 = success, otherwise an STV_E_* code; stv_last_error() returns a human-readable message
 *     (thread-local) for the last failing call;
 *   - scratch memory is passed in as `ws` with at least the number of bytes reported by the matching *_workspace_bytes().
 */
#ifndef STV_H_
#define STV_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STV_OK 0
#define STV_E_ARG 1       /* invalid argument (shape, null pointer, unsupported configuration) */
#define STV_E_WORKSPACE 2 /* workspace too small */
#define STV_E_CUDA 3      /* CUDA runtime / launch error */

#define STV_MAX_SCALES 8
#define STV_MAX_SUPPORT 8
#define STV_SEL_STATIC 255 /* `sel` value: pixel auto-masked (identity error won) */
#define STV_SEL_MEAN 254   /* `sel` value: use_min=0, every support frame contributes 1/n */

int stv_version(void);
const char* stv_last_error(void);
/* Number of kernels launched by this library since load (all threads); used by bench.py's `gpu_launches`. */
unsigned long long stv_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Photometric view-synthesis loss: replaces handlers.image_recon (src/core/handlers.py:14-67) =
 *   ViewSynth.forward (src/tools/geometry.py:366-391: BackprojectDepth :304-316, ProjectPoints :329-350, grid_sample :364)
 *   + ReconstructionLoss.forward (src/losses/reconstruction.py:98-126: compute_photo :79-96, apply_automask :59-77)
 *   + PhotoError / SSIMError / DenseL1Error (src/losses/photometric.py:11-88).
 * Index order of every (S, b, ...) tensor is s*b + i, of (n, b, ...) k*b + i, as in handlers.py:48-60.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int b, n, S, H, W;
    float w_ssim, w_l1;       /* 0.85/0.15 for loss_name='ssim', 0/1 for 'l1' (reconstruction.py:37-41) */
    int use_min;              /* reconstruction.py:43-44 */
    int use_automask;         /* reconstruction.py:59-77 */
    uint64_t noise_seed;      /* automask tie-break noise when `noise == NULL`: 0 = none, else in-kernel Philox normal */
    int64_t depth_stride_s;   /* unused when per-scale pointers are given; reserved */
} stv_photo_cfg;

size_t stv_photo_workspace_bytes(const stv_photo_cfg* cfg);

/* loss (device scalar) = mean over (S,b,H,W) of the reduced, auto-masked photometric error.
 * depth: S pointers to (b,1,H,W) upsampled depth maps; tgt (b,3,H,W); supp (n,b,3,H,W); T (n,b,4,4); K, Kinv (b,4,4);
 * noise: NULL or (S,b,H,W) standard-normal samples replacing randn_like (reconstruction.py:72);
 * noise_step: NULL or a DEVICE counter: when the noise is drawn in-kernel (noise == NULL, noise_seed != 0) the effective seed is
 *   noise_seed + *noise_step and the call advances *noise_step by one, stream-ordered after its last reader — a captured CUDA
 *   graph therefore draws fresh noise on every replay, as torch.randn_like does every step;
 * sel (S,b,H,W) u8: per-pixel decision (support index | STV_SEL_STATIC | STV_SEL_MEAN), consumed by the backward;
 * warp0: NULL or (n,b,3,H,W) warped support frames at scale 0 (handlers.py:66).
 * Two-pass formulation, any reduction (min / mean) and any n <= STV_MAX_SUPPORT; the default configuration (min-reprojection,
 * n <= STV_FUSED_MAX_SUPPORT) is served by the single-pass stv_photo_fused_fwd below. */
int stv_photo_fwd(const stv_photo_cfg* cfg, const float* const* depth, const float* tgt, const float* supp,
                  const float* T, const float* K, const float* Kinv, const float* noise, unsigned long long* noise_step,
                  float* loss, uint8_t* sel, float* warp0, void* ws, size_t ws_bytes, void* stream);

/* Backward of stv_photo_fwd w.r.t. depth, T, K and Kinv. grad_loss: device scalar dL/dloss.
 * g_depth: S pointers to (b,1,H,W) (overwritten); gT (n,b,4,4) (overwritten; row 3 = 0);
 * gK, gKinv (b,4,4) nullable (overwritten; only the 3x3 block is non-zero).
 * Self-contained: re-warps a halo-2 tile per support frame and rebuilds the SSIM window sums. */
int stv_photo_bwd(const stv_photo_cfg* cfg, const float* const* depth, const float* tgt, const float* supp,
                  const float* T, const float* K, const float* Kinv, const uint8_t* sel, const float* grad_loss,
                  float* const* g_depth, float* gT, float* gK, float* gKinv, void* ws, size_t ws_bytes, void* stream);

/* compute_photo (reconstruction.py:79-96) on its own: pred (n,b,3,H,W) vs target (b,3,H,W) -> err (b,1,H,W).
 * Forward only (used for the identity/static error and by `depth_regr`, src/core/trainer.py:430). */
int stv_photo_error(const stv_photo_cfg* cfg /* b,n,H,W,w_ssim,w_l1,use_min */, const float* pred, const float* tgt,
                    float* err, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Single-pass fused photometric loss (min-reprojection, <= STV_FUSED_MAX_SUPPORT support frames): same inputs and semantics as
 * stv_photo_fwd, but the loss AND its gradients come out of ONE sweep (the loss is a mean, so d loss/d(.) is known up to the
 * scalar dL/dloss as soon as a pixel's decision is taken) — no SSIM coefficient planes or re-warp between a forward and a
 * backward kernel. The kernel can also consume the network's low-resolution sigmoid disparities directly, fusing
 * ops.interpolate_like + to_scaled (src/core/trainer.py:320-321) into the sweep.
 *   src->mode STV_PHOTO_SRC_DEPTH: maps[s] = (b,1,H,W) up-sampled depth (what handlers.image_recon receives, handlers.py:48);
 *   src->mode STV_PHOTO_SRC_DISP:  maps[s] = (b,1,h[s],w[s]) sigmoid disparity; depth = to_inv(to_scaled(upsample(disp))) with
 *                                  min_depth / max_depth (<= 0: unset -> to_inv only).
 * supp_tex: 0, or a handle from stv_tex_create(supp, n*b*3*H, W) (2x2 gathers through the texture unit; same results).
 * g_unit: NULL (forward only), or S pointers to (b,1,H,W) receiving d loss/d(up-sampled map) for dL/dloss = 1 (mode DEPTH: d/d depth;
 *   mode DISP: d/d up-sampled disparity); gpart: stv_photo_fused_partial_bytes() of device memory receiving per-strip moment sums
 *   of the pose / intrinsics gradients. Both are handed unchanged to stv_photo_fused_bwd.
 * Everything else as stv_photo_fwd (loss, sel, warp0, noise, noise_step, ws >= stv_photo_fused_workspace_bytes()).
 * ------------------------------------------------------------------------------------------------------------------ */
#define STV_FUSED_MAX_SUPPORT 4
#define STV_PHOTO_SRC_DEPTH 0
#define STV_PHOTO_SRC_DISP 1
typedef struct {
    int mode;
    int h[STV_MAX_SCALES], w[STV_MAX_SCALES];
    float min_depth, max_depth;
} stv_photo_src;

size_t stv_photo_fused_workspace_bytes(const stv_photo_cfg* cfg);
size_t stv_photo_fused_partial_bytes(const stv_photo_cfg* cfg);
int stv_photo_fused_fwd(const stv_photo_cfg* cfg, const stv_photo_src* src, const float* const* maps, const float* tgt,
                        const float* supp, unsigned long long supp_tex, const float* T, const float* K, const float* Kinv,
                        const float* noise, unsigned long long* noise_step, float* loss, uint8_t* sel, float* warp0,
                        float* const* g_unit, float* gpart, void* ws, size_t ws_bytes, void* stream);
/* Backward of stv_photo_fused_fwd: nothing is recomputed. g_maps[s] (nullable array; overwritten) = grad_loss * g_unit[s] pulled back
 * to the layout of maps[s] (mode DISP: through the adjoint of the bilinear up-sampling, deterministic gathers; mode DEPTH: as is;
 * g_maps[s] may alias g_unit[s] in mode DEPTH). gT (n,b,4,4) (nullable), gK, gKinv (b,4,4) (nullable) from gpart, scaled by
 * grad_loss (device scalar). ws >= stv_photo_fused_bwd_workspace_bytes(). */
size_t stv_photo_fused_bwd_workspace_bytes(const stv_photo_cfg* cfg, const stv_photo_src* src);
int stv_photo_fused_bwd(const stv_photo_cfg* cfg, const stv_photo_src* src, const float* grad_loss, const float* const* g_unit,
                        const float* gpart, const float* T, const float* Kinv, float* const* g_maps, float* gT, float* gK,
                        float* gKinv, void* ws, size_t ws_bytes, void* stream);
/* Texture view of a (rows, W) fp32 plane stack for the 2x2 gathers of the photometric kernels. *handle = 0 when the buffer cannot
 * be bound (alignment / size limits): the kernels then use plain loads. The CALLER owns the handle (create it outside CUDA-graph
 * capture, destroy it when the buffer goes away); the library keeps no cache. */
int stv_tex_create(const float* ptr, long long rows, int W, unsigned long long* handle);
int stv_tex_destroy(unsigned long long handle);

/* ReconstructionLoss.forward on ALREADY WARPED frames (src/losses/reconstruction.py:98-126; what a caller of the registered
 * `img_recon` class gets when it warps elsewhere, e.g. the virtual-stereo branch src/core/trainer.py:394-399), and its gradient
 * w.r.t. those frames. cfg: b, n, H, W, w_ssim, w_l1, use_min, use_automask, noise_seed (S is ignored).
 * pred (n,b,3,H,W); tgt (b,3,H,W); source (n,b,3,H,W) un-warped frames (required when use_automask, reconstruction.py:121);
 * noise NULL | (b,1,H,W); noise_step as in stv_photo_fwd; loss: device scalar = mean over (b,H,W); sel (b,H,W) u8 decisions;
 * err: NULL | (b,1,H,W) the reduced, auto-masked error map. Cold path: plain one-thread-per-pixel kernels. */
size_t stv_recon_workspace_bytes(const stv_photo_cfg* cfg);
int stv_recon_fwd(const stv_photo_cfg* cfg, const float* pred, const float* tgt, const float* source, const float* noise,
                  unsigned long long* noise_step, float* loss, uint8_t* sel, float* err, void* ws, size_t ws_bytes, void* stream);
/* g_pred (n,b,3,H,W) (overwritten) = grad_loss * d loss / d pred given the decisions `sel` of the forward call. */
int stv_recon_bwd(const stv_photo_cfg* cfg, const float* pred, const float* tgt, const uint8_t* sel, const float* grad_loss,
                  float* g_pred, void* stream);

/* The whole registered `img_recon` / `feat_recon` / `autoenc_recon` class on ALREADY WARPED frames (SURVEY 8f rank 4;
 * src/losses/reconstruction.py:13-126): any channel count C (feat_recon hands encoder features, src/core/handlers.py:70-119),
 * loss_name ssim | l1 | l2 (src/losses/photometric.py:12-23, 54-88), explainability / uncertainty weighting masks
 * (reconstruction.py:46-57), min / mean reduction, automask. pred, source (n,b,C,H,W); tgt (b,C,H,W); mask (b,n,H,W).
 * sel (b,H,W) u8: k = warped frame k carries the pixel, 0x40 = mean of the frames; bit 7 set = the static error won (low bits:
 * the static frame that was the minimum). err (b,H,W) nullable: the reduced error map (compute_photo / apply_automask).
 * stv_recon_ex_bwd: g_pred (n,b,C,H,W) and / or g_mask (b,n,H,W) (either may be NULL). */
#define STV_RECON_SSIM 0
#define STV_RECON_L1 1
#define STV_RECON_L2 2
#define STV_RECON_MASK_NONE 0
#define STV_RECON_MASK_EXPLAIN 1
#define STV_RECON_MASK_UNCERT 2
typedef struct {
    int b, n, C, H, W;
    int loss;                      /* STV_RECON_* */
    int use_min, use_automask;
    int mask_mode;                 /* STV_RECON_MASK_* */
    unsigned long long noise_seed; /* as stv_photo_cfg.noise_seed */
} stv_recon_cfg;
size_t stv_recon_ex_workspace_bytes(const stv_recon_cfg* cfg);
int stv_recon_ex_fwd(const stv_recon_cfg* cfg, const float* pred, const float* tgt, const float* source, const float* mask,
                     const float* noise, unsigned long long* noise_step, float* loss, uint8_t* sel, float* err, void* ws,
                     size_t ws_bytes, void* stream);
int stv_recon_ex_bwd(const stv_recon_cfg* cfg, const float* pred, const float* tgt, const float* source, const float* mask,
                     const uint8_t* sel, const float* grad_loss, float* g_pred, float* g_mask, void* stream);

/* ViewSynth.forward (src/tools/geometry.py:366-391) on its own: input (B,C,H,W), depth (B,1,H,W), T,K,Kinv (B,4,4) ->
 * warp (B,C,H,W), depth_warp (B,1,H,W), mask_valid (B,1,H,W) u8. Any C; used by the stand-alone ViewSynth module. */
int stv_view_synth_fwd(int B, int C, int H, int W, const float* input, const float* depth, const float* T,
                       const float* K, const float* Kinv, float* warp, float* depth_warp, uint8_t* mask_valid,
                       void* stream);
/* Backward w.r.t. depth/T/K/Kinv (and input when g_input != NULL, accumulated with atomics into a zeroed buffer).
 * g_warp (B,C,H,W), g_depth_warp (B,1,H,W) nullable. partial: workspace of stv_view_synth_workspace_bytes(). */
size_t stv_view_synth_workspace_bytes(int B, int C, int H, int W);
int stv_view_synth_bwd(int B, int C, int H, int W, const float* input, const float* depth, const float* T,
                       const float* K, const float* Kinv, const float* g_warp, const float* g_depth_warp,
                       float* g_depth, float* gT, float* gK, float* gKinv, float* g_input,
                       void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Disparity post-processing: replaces ops.interpolate_like (src/tools/ops.py:311-314, trainer.py:320) fused with
 * to_scaled / to_inv (src/tools/geometry.py:62-76, 86-90, trainer.py:49,321).
 * disp (b,1,h,w) -> disp_up (b,1,H,W) [nullable], depth_up (b,1,H,W). min_depth/max_depth <= 0 mean "unset".
 * ------------------------------------------------------------------------------------------------------------------ */
int stv_disp_to_depth_fwd(int b, int h, int w, int H, int W, float min_depth, float max_depth, const float* disp,
                          float* disp_up, float* depth_up, void* stream);
/* g_disp (b,1,h,w) = d(depth_up)/d(disp)^T g_depth_up [+ d(disp_up)/d(disp)^T g_disp_up]; either gradient may be NULL.
 * Two separable deterministic gathers (rows, then columns); ws >= stv_disp_to_depth_bwd_workspace_bytes(). */
size_t stv_disp_to_depth_bwd_workspace_bytes(int b, int h, int w, int H, int W);
int stv_disp_to_depth_bwd(int b, int h, int w, int H, int W, float min_depth, float max_depth, const float* disp,
                          const float* g_depth_up, const float* g_disp_up, float* g_disp, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Edge-aware smoothness: replaces handlers.disp_smooth (src/core/handlers.py:262-281) = per scale
 *   interpolate_like(imgs, disp) (:278) + SmoothReg.forward (src/regularizers/smooth.py:71-97, compute_grad :12-30,
 *   ops.mean_normalize src/tools/ops.py:279-286); loss = mean_s(loss_s / scale_div[s]).
 * disp: S pointers to (b,1,h_s,w_s); img (b,3,H,W).
 * disp_grad / image_grad: NULL or (b,1,h_0,w_0) logging maps of the FIRST scale (handlers.py:280).
 * stats: workspace-resident per (s,i) {mean, loss_sum} consumed by the backward (kept inside `ws`).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int b, S, H, W;
    int h[STV_MAX_SCALES], w[STV_MAX_SCALES];
    float scale_div[STV_MAX_SCALES]; /* 2**s of the reference's dict key (handlers.py:279) */
    int use_edges;                   /* smooth.py:91-94 */
} stv_smooth_cfg;

size_t stv_smooth_workspace_bytes(const stv_smooth_cfg* cfg);
int stv_smooth_fwd(const stv_smooth_cfg* cfg, const float* const* disp, const float* img, float* loss,
                   float* disp_grad, float* image_grad, void* ws, size_t ws_bytes, void* stream);
/* `ws` must be the workspace filled by the matching stv_smooth_fwd call. g_disp: S pointers (overwritten). */
int stv_smooth_bwd(const stv_smooth_cfg* cfg, const float* const* disp, const float* img, const float* grad_loss,
                   float* const* g_disp, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * ConvNeXt block, memory-bound half (channels-last (N,H,W,C) fp32): replaces the ATen/cuDNN ops launched by the timm
 * ConvNeXt encoder the reference builds at src/networks/depth.py:97 (timm==0.6.12, third-party): `conv_dw` (depthwise 7x7,
 * padding 3) and `norm` (LayerNorm over C, eps 1e-6) of every ConvNeXtBlock, plus their backward.
 * w is the depthwise filter (C,1,7,7) contiguous = (C,49).
 * ------------------------------------------------------------------------------------------------------------------ */
/* y = dwconv7x7(x; w) [+ bias] [+ res]. flip != 0 applies the 180-degree rotated filter (= data gradient of the conv). */
int stv_dwconv7_fwd(int N, int H, int W, int C, const float* x, const float* w, const float* bias, const float* res,
                    float* y, int flip, void* stream);
size_t stv_dwconv7_wgrad_workspace_bytes(int N, int H, int W, int C);
/* gw (C,49) = d/dw, gb (C) = d/dbias (nullable), from the layer input x and the output gradient gy; accumulate != 0: += . */
int stv_dwconv7_wgrad(int N, int H, int W, int C, const float* x, const float* gy, float* gw, float* gb, int accumulate, void* ws,
                      size_t ws_bytes, void* stream);
/* Row-wise LayerNorm of a (P, C) matrix; mean / rstd (P) are saved for the backward. */
int stv_layernorm_fwd(long long P, int C, const float* x, const float* gamma, const float* beta, float eps, float* y,
                      float* mean, float* rstd, void* stream);
size_t stv_layernorm_bwd_workspace_bytes(long long P, int C);
int stv_layernorm_bwd(long long P, int C, const float* dy, const float* x, const float* mean, const float* rstd,
                      const float* gamma, float* dx, float* dgamma, float* dbeta, int accumulate /* dgamma, dbeta += */, void* ws,
                      size_t ws_bytes, void* stream);

/* RegressionLoss (src/losses/regression.py:11-75), the class registered as `stereo_const` and `depth_regr`
 * (src/core/handlers.py:151-259): loss = sum(mask e)/sum(mask), e = |p - t| (STV_REGR_L1), log(1 + |p - t|) (LOG_L1) or berHu with
 * the dynamic threshold 0.2 max|p - t| (BERHU); invert: p, t = to_inv(pred), to_inv(target). pred, target, mask (nullable: ones),
 * err (nullable: the masked error map `err_regr`), g_pred, g_target (either nullable): n contiguous floats. Deterministic
 * (per-block partials, fixed order). The backward needs the workspace the forward filled (stv_regr_workspace_bytes()). */
#define STV_REGR_L1 0
#define STV_REGR_LOG_L1 1
#define STV_REGR_BERHU 2
size_t stv_regr_workspace_bytes(void);
int stv_regr_fwd(long long n, int loss, int invert, const float* pred, const float* target, const float* mask, float* loss_out,
                 float* err, void* ws, size_t ws_bytes, void* stream);
int stv_regr_bwd(long long n, int loss, int invert, const float* pred, const float* target, const float* mask,
                 const float* grad_loss, float* g_pred, float* g_target, void* ws, size_t ws_bytes, void* stream);

/* The two pointwise regularisers: OccReg (`disp_occ`, src/regularizers/occlusion.py:9-40): sign * mean(x); MaskReg (`disp_mask`,
 * src/regularizers/mask.py:11-30): binary cross-entropy of x against 1. ws >= stv_regr_workspace_bytes(). */
#define STV_PWREG_MEAN 0
#define STV_PWREG_BCE_ONE 1
int stv_pwreg_fwd(long long n, int kind, float sign, const float* x, float* loss, void* ws, size_t ws_bytes, void* stream);
int stv_pwreg_bwd(long long n, int kind, float sign, const float* x, const float* grad_loss, float* g, void* stream);

/* Feature regularisers (`feat_peaky` = FeatPeakReg, order 1; `feat_smooth` = FeatSmoothReg, order 2; src/regularizers/smooth.py:100-176):
 * first- / second-order absolute differences of C-channel feature maps, optionally weighted by exp(-image differences) (use_edges).
 * feat (b,C,H,W), img (b,Ci,H,W); feat_grad (b,C,H,W) nullable logging map. The backward needs the workspace the forward filled. */
size_t stv_feat_reg_workspace_bytes(int b, int C, int Ci, int H, int W);
int stv_feat_reg_fwd(int b, int C, int Ci, int H, int W, int order, int use_edges, const float* feat, const float* img, float* loss,
                     float* feat_grad, void* ws, size_t ws_bytes, void* stream);
int stv_feat_reg_bwd(int b, int C, int Ci, int H, int W, int order, int use_edges, const float* feat, const float* grad_loss,
                     float* g_feat, void* ws, size_t ws_bytes, void* stream);

/* SmoothReg.forward with every constructor flag (src/regularizers/smooth.py:12-97): use_laplacian = second-order
 * absolute gradients (compute_laplacian, :33-48), use_blur = 3x3 sigma-1 Gaussian pre-blur of every differentiated map
 * (kornia.filters.gaussian_blur2d, reflect border), use_edges = exp(-|image gradient|) weights. Single scale: disp (b,1,H,W),
 * img (b,C,H,W) at the same resolution. loss (); disp_grad / image_grad (b,1,H,W) nullable logging maps (:88-92).
 * The backward needs the workspace the forward filled (>= stv_smooth_ex_workspace_bytes). */
size_t stv_smooth_ex_workspace_bytes(int b, int C, int H, int W);
int stv_smooth_ex_fwd(int b, int C, int H, int W, int use_edges, int use_laplacian, int use_blur, const float* disp,
                      const float* img, float* loss, float* disp_grad, float* image_grad, void* ws, size_t ws_bytes, void* stream);
int stv_smooth_ex_bwd(int b, int C, int H, int W, int use_edges, int use_laplacian, int use_blur, const float* disp,
                      const float* grad_loss, float* g_disp, void* ws, size_t ws_bytes, void* stream);

/* Reproducible mode. With STV_DETERMINISTIC=1 in the environment (read once per process) every floating-point accumulation of the
 * library has one contributor per address per launch: stv_gemm_tf32 / stv_conv_wgrad ignore split_k (and stv_gemm_tf32 refuses the
 * fused `colsum`), the column reductions behind stv_colsum / stv_act_bwd / stv_bn_* run on one row block, stv_head3x3_bwd accumulates
 * its weight gradient one x-tile per launch. Two runs of the same step then give bit-identical gradients (tests/test_determinism_gpu.py);
 * the step is several times slower. Without it the weight / bias gradients and the BatchNorm sums are accumulated with atomics whose
 * order varies from run to run (differences at rounding level). */

/* ------------------------------------------------------------------------------------------------------------------
 * Tensor-core products of the network layers (tcgen05.mma kind::tf32 + TMA + TMEM; fp32 storage, TF32 multiply, fp32
 * accumulate = the reference's `torch.set_float32_matmul_precision('high')`, src/core/trainer.py:30).
 * Replaces the cuBLAS / cuDNN calls behind nn.Linear and 1x1 nn.Conv2d (forward, data gradient, weight gradient) of the
 * timm encoders (src/networks/depth.py:97, pose.py:40) and the pose heads (src/networks/pose.py:46,75-106).
 *
 *   C[M,N] (+)= epilogue( sum_k A[m,k] * B[n,k] )
 * a_mn = 0: A is stored row-major [M][lda] (k contiguous);  a_mn = 1: A is stored [K][lda] (m contiguous).
 * b_mn = 0: B is stored row-major [N][ldb] (k contiguous);  b_mn = 1: B is stored [K][ldb] (n contiguous).
 * Epilogue, in this order:  v = acc + bias[n];  aux[m,n] = v;  v = act(v);  v *= gamma[n];  v += res[m,n];
 *                           v *= act'(dact_src[m,n]);  C[m,n] = v  (accumulate = 0)  or  C[m,n] += v  (atomic);  colsum[n] += v.
 * `dact_src` holds the pre-activation for GELU and the activation OUTPUT for ReLU / ELU / sigmoid.
 * aux / res / dact_src share C's leading dimension ldc. Alignment: every pointer 16 bytes; N, lda, ldb, ldc multiples of 4.
 * split_k > 1 partitions the reduction over gridDim.z and requires accumulate = 1 into a pre-initialised C.
 * ------------------------------------------------------------------------------------------------------------------ */
#define STV_ACT_NONE 0
#define STV_ACT_RELU 1
#define STV_ACT_GELU 2    /* exact (erf) GELU, torch.nn.functional.gelu default */
#define STV_ACT_ELU 3     /* alpha = 1 */
#define STV_ACT_SIGMOID 4

typedef struct {
    const float* bias;     /* [N] or NULL */
    float* aux;            /* [M][ldc] or NULL */
    const float* gamma;    /* [N] or NULL */
    const float* res;      /* [M][ldc] or NULL */
    const float* dact_src; /* [M][ldc] or NULL */
    float* colsum;         /* [N] or NULL: colsum[n] += sum_m (final value written for C[m,n]) — bias gradients, atomically accumulated */
    int act;               /* STV_ACT_* applied forward */
    int dact;              /* STV_ACT_* whose derivative multiplies the result when dact_src != NULL */
    int accumulate;
} stv_gemm_epi;

int stv_gemm_tf32(int M, int N, int K, const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn,
                  float* C, long long ldc, const stv_gemm_epi* epi, int split_k, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Spatial convolutions as implicit GEMMs on the tcgen05 tensor cores (TF32 multiply, fp32 accumulate; activations fp32
 * channels-last (N,H,W,C); filters (Cout,R,S,Cin) contiguous, i.e. a PyTorch channels-last (Cout,Cin,R,S) weight).
 * Replaces the cuDNN calls behind every nn.Conv2d of the Monodepth decoder (src/networks/decoders/monodepth.py:51-89,
 * utils.py:44-54: reflect-padded 3x3 + ELU / sigmoid, with the `F.interpolate(scale_factor=2, 'nearest')` and skip
 * `torch.cat` of :76-79 fused into the operand gather), of the pose network (src/networks/pose.py:40,46,75-106) and of the
 * timm encoder stems / down-sampling layers (src/networks/depth.py:97).
 *
 * The convolution input is the VIRTUAL tensor V = cat(src1', src2) along channels (C = C1 + C2), where src1' is src1 itself
 * (up1 = 0) or its nearest x2 upsampling (up1 = 1: src1 is stored as (N, H/2, W/2, C1)); V is padded by `pad` with zeros or by
 * reflection. It is never materialised. Output size P = (H + 2 pad - R)/stride + 1 (same for Q).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int N, H, W;   /* batch, height, width of the virtual input V */
    int C1, C2;    /* channels of src1 / src2 (C2 = 0: single source); multiples of 4 */
    int up1;       /* 1: src1 is stored at half resolution and nearest-upsampled x2 on the fly */
    int Cout, R, S, stride, pad;
    int reflect;   /* 0: zero padding; 1: reflection padding (nn.Conv2d(padding_mode='reflect')) */
} stv_conv_geom;

/* y (N,P,Q,Cout) = epilogue(conv(V, w)); the epilogue is the one of stv_gemm_tf32 (bias, activation, aux, ...; no accumulate). */
int stv_conv_fprop(const stv_conv_geom* g, const float* src1, const float* src2, const float* w, float* y,
                   const stv_gemm_epi* epi, void* stream);
/* dv = d loss / d V given dy (N,P,Q,Cout); Cout % 4 == 0. With zero padding dv is (N,H,W,C); with reflection padding dv is
 * produced on the PADDED grid (N, H+2pad, W+2pad, C) and stv_grad_pull folds the border back. */
int stv_conv_dgrad(const stv_conv_geom* g, const float* dy, const float* w, float* dv, const stv_gemm_epi* epi, void* stream);
/* dw (Cout,R,S,C) += dy^T im2col(V) (atomic accumulation: dw must be initialised); Cout % 4 == 0. split_k <= 0: automatic. */
int stv_conv_wgrad(const stv_conv_geom* g, const float* src1, const float* src2, const float* dy, float* dw, int split_k,
                   void* stream);
/* out (N, H+2p, W+2p, Cp) = the virtual input V materialised: cat(up2(src1) | src1, src2), reflection-padded by p = g.pad when
 * g.reflect (p = 0 otherwise), channels [C1+C2, Cp) zero. Lets a reflect / upsample / concat / narrow-channel convolution run
 * through the TMA im2col path (pad 0 on `out`, filters zero-padded to Cp input channels). */
int stv_vpad(const stv_conv_geom* g, const float* src1, const float* src2, int Cp, float* out, void* stream);
/* dst (N,H,W,C) (+)= slice/fold/pool of src (N, H*pool + 2 pad, W*pool + 2 pad, Cs): channels [c_off, c_off + C), the
 * reflection-padding border folded back onto the interior, and pool x pool (1 or 2) sum-pooling (adjoint of nearest x2). */
int stv_grad_pull(int N, int H, int W, int C, const float* src, int Cs, int c_off, int pad, int pool, float* dst, int accumulate,
                  void* stream);
/* dz[m,c] = da[m,c] * act'(y[m,c]) (y = activation OUTPUT; pre-activation for GELU); dbias[c] += sum_m dz[m,c] (nullable). */
int stv_act_bwd(long long M, int C, const float* da, const float* y, int act, float* dz, float* dbias, void* stream);
/* ---------------------------------------------------------------------------------------------------------------------
 * Aspect-ratio augmentation (src/core/aspect_ratio.py:36-186, called from training_step, src/core/trainer.py:106): bilinear
 * resampling of (P, H, W) fp32 planes to (P, oh, ow) with separable sample positions.
 *   STV_RESAMPLE_GRID   ix = ax*j + bx, iy = ay*i + by (pixel units), zero padding: kornia.center_crop(mode='bilinear',
 *                       align_corners=False) = warp_affine -> F.affine_grid + F.grid_sample (crop_aug, aspect_ratio.py:82)
 *   STV_RESAMPLE_INTERP F.interpolate(size, mode='bilinear', align_corners=False) with ax = W/ow, ay = H/oh (resize_aug, :139)
 * ------------------------------------------------------------------------------------------------------------------ */
#define STV_RESAMPLE_GRID 0
#define STV_RESAMPLE_INTERP 1
int stv_resample_bilinear(long long P, int H, int W, int oh, int ow, float ax, float bx, float ay, float by, int mode,
                          const float* src, float* dst, void* stream);

/* Logging statistics (src/core/trainer.py:486-503 `summarize_depth`: `.mean().item()` / `.std().item()` per up-sampled
 * disparity / depth map, 16 host syncs per logging step): mean and unbiased standard deviation of k <= STV_STATS_MAX device
 * tensors in one launch pair. out (k,2) fp32 = {mean, std}; ws >= stv_mean_std_workspace_bytes(k). No host synchronisation. */
#define STV_STATS_MAX 16
size_t stv_mean_std_workspace_bytes(int k);
int stv_mean_std(int k, const float* const* tensors, const long long* counts, float* out, void* ws, size_t ws_bytes, void* stream);

/* out[c] += sum_m x[m*ld + c] */
int stv_colsum(long long M, int C, long long ld, const float* x, float* out, void* stream);
/* Parameter gradients behind the ConvNeXt layer-scale (timm ConvNeXtBlock: x + gamma * fc2(...), encoder built at
 * src/networks/depth.py:97). G = g^T h (C, Hd), gs = column sums of g (C), w2 (C, Hd):
 *   dw2 += gamma[:,None]*G;  db2 += gamma*gs;  dgamma += sum_j w2[c,j]*G[c,j] + b2*gs.   (all accumulated in place) */
int stv_ls_tail(int C, int Hd, const float* G, const float* w2, const float* b2, const float* gamma, const float* gs,
                float* dw2, float* db2, float* dgamma, void* stream);
/* out[c,j] = w[c,j] * gamma[c]  (layer-scale folded into fc2 for the data gradient) */
int stv_rowscale(int C, int Hd, const float* w, const float* gamma, float* out, void* stream);

/* Train-mode BatchNorm over the M = N*H*W rows of a channels-last (M, C) matrix, fused with the residual add and ReLU that
 * follow it in a ResNet BasicBlock: y = [relu]((x - mean_c) * rstd_c * gamma_c + beta_c [+ res]). Replaces the cuDNN batch-norm
 * + ATen add / relu kernels behind nn.BatchNorm2d / F.relu of the timm ResNet encoder (src/networks/pose.py:40, depth.py:97).
 * Batch statistics per GPU (biased variance for normalisation); run_mean / run_var (nullable) get nn.BatchNorm2d's momentum
 * update with the unbiased variance. mean / rstd (C) are saved for the backward. C % 4 == 0. */
size_t stv_bn_workspace_bytes(int C);
int stv_bn_fwd(long long M, int C, const float* x, const float* gamma, const float* beta, const float* res, int relu, float eps,
               float momentum, float* y, float* mean, float* rstd, float* run_mean, float* run_var, void* ws, size_t ws_bytes,
               void* stream);
/* dz = dy * (y > 0) when relu; dres = dz (nullable); dx = gamma rstd (dz - mean_m(dz) - xhat mean_m(dz xhat)); dgamma, dbeta (C). */
int stv_bn_bwd(long long M, int C, const float* dy, const float* y, const float* x, const float* mean, const float* rstd,
               const float* gamma, int relu, float* dx, float* dres, float* dgamma, float* dbeta, int accumulate /* dgamma, dbeta += */,
               void* ws, size_t ws_bytes, void* stream);

/* 3x3 stride-2 max-pool, padding 1, channels-last (N,H,W,C) -> (N,(H-1)/2+1,(W-1)/2+1,C): the `maxpool` of the timm ResNet stem
 * (src/networks/pose.py:40, depth.py:97). idx (same shape as y, u8) = winning tap 0..8 (first maximum in row-major order). */
int stv_maxpool3x3s2_fwd(int N, int H, int W, int C, const float* x, float* y, uint8_t* idx, void* stream);
int stv_maxpool3x3s2_bwd(int N, int H, int W, int C, const float* dy, const uint8_t* idx, float* dx, void* stream);

/* Disparity heads: y (N,H,W) = act(bias + reflect-padded 3x3 convolution of x (N,H,W,C) with w (3,3,C)) — `outconv_i` + sigmoid of
 * the Monodepth decoder (src/networks/decoders/monodepth.py:66-69,86-87). One output channel = a dot product per pixel:
 * memory-bound, CUDA cores. C a power of two in [4,128]. Backward: dz = da*act'(y); dx (N,H,W,C) (nullable, overwritten);
 * dw (3,3,C) and db (1) (nullable, atomically accumulated); dz_ws: N*H*W floats of scratch (receives dz). */
int stv_head3x3_fwd(int N, int H, int W, int C, const float* x, const float* w, const float* bias, int act, float* y, void* stream);
int stv_head3x3_bwd(int N, int H, int W, int C, const float* x, const float* w, const float* da, const float* y, int act, float* dx,
                    float* dw, float* db, float* dz_ws, void* stream);

/* Batched 4x4 inverse B = A^-1 (n matrices, row-major) and its backward gA = -B^T gB B^T: `K.inverse()` of ViewSynth.forward
 * (src/tools/geometry.py:383) and `T.inverse()` of the backward poses (src/core/trainer.py:253). Stream-ordered, no host sync
 * (ATen's batched LU synchronises the host and cannot be captured in a CUDA graph). */
int stv_inv4x4(int n, const float* A, float* B, void* stream);
int stv_inv4x4_bwd(int n, const float* B, const float* gB, float* gA, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Optimiser: replaces torch.optim.AdamW(foreach) built by timm create_optimizer_v2 (src/tools/parsers.py:205-243)
 * on one flat fp32 parameter/gradient buffer. `wd` is a per-element weight-decay mask value selector: elements in
 * [0, n_decay) use `weight_decay`, elements in [n_decay, n) use 0 (timm excludes biases / 1-D params).
 * grad_scale multiplies the gradient first (1/world_size after a sum all-reduce). step is 1-based.
 * ------------------------------------------------------------------------------------------------------------------ */
int stv_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, size_t n_decay,
                   float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, int step,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STV_H_ */
