// ReconstructionLoss.forward on ALREADY WARPED frames (src/losses/reconstruction.py:98-126) and its gradient w.r.t. them.
//
// This is the cold side of the reference's `img_recon` class: the training hot path never materialises the warped frames (the
// warp is fused into stv_photo_fwd), but anything that calls the registered class directly — the virtual-stereo branch
// (src/core/trainer.py:394-399), ablations that warp elsewhere — hands it (*n,b,3,H,W) predictions. Straightforward
// one-thread-per-pixel kernels: every thread rebuilds the 3x3 window sums it needs from global memory (L1/L2 resident), no tiling.
#include "stv_common.cuh"

namespace stv {

struct ReconParams {
    int b, n, H, W;
    float w_ssim, w_l1;
    int use_min, use_automask;
    uint64_t seed;
    const unsigned long long* step;
    const float *pred, *tgt, *src, *noise, *grad_loss;
    const uint8_t* sel_in;
    float *err, *partial, *g_pred;
    uint8_t* sel;
};

// Window sums of one channel plane around (y, x) with reflection padding 1.
__device__ __forceinline__ void win5(const float* __restrict__ w, const float* __restrict__ t, int y, int x, int H, int W, float& S1,
                                     float& S2, float& S3, float& T1, float& T2) {
    S1 = S2 = S3 = T1 = T2 = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int ya = min(max(reflect_idx(y + dy, H), 0), H - 1);
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int xa = min(max(reflect_idx(x + dx, W), 0), W - 1);
            const float a = __ldg(w + ya*W + xa), c = __ldg(t + ya*W + xa);
            S1 += a; S2 = fmaf(a, a, S2); S3 = fmaf(a, c, S3); T1 += c; T2 = fmaf(c, c, T2);
        }
    }
}

// PhotoError (photometric.py:54-88) of a 3-channel frame against the target at one pixel.
__device__ __forceinline__ float photo_at(const ReconParams& p, const float* __restrict__ w, const float* __restrict__ t, int y, int x) {
    const int HW = p.H*p.W;
    float e = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (p.w_ssim > 0.f) {
            float S1, S2, S3, T1, T2;
            win5(w + c*HW, t + c*HW, y, x, p.H, p.W, S1, S2, S3, T1, T2);
            e = fmaf(p.w_ssim*(1.f/3.f), ssim_err(S1, S2, S3, T1, T2), e);
        }
        if (p.w_l1 > 0.f) e = fmaf(p.w_l1*(1.f/3.f), fabsf(__ldg(w + c*HW + y*p.W + x) - __ldg(t + c*HW + y*p.W + x)), e);
    }
    return e;
}

__global__ void __launch_bounds__(256) recon_fwd_kernel(ReconParams p) {
    __shared__ float red[32];
    const int HW = p.H*p.W;
    const int i = blockIdx.y;
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    float e = 0.f;
    if (q < HW) {
        const int y = q/p.W, x = q - y*p.W;
        const float* t = p.tgt + (size_t)i*3*HW;
        float er = p.use_min ? INFINITY : 0.f;
        int sel = p.use_min ? 0 : STV_SEL_MEAN;
        for (int k = 0; k < p.n; ++k) {
            const float ek = photo_at(p, p.pred + ((size_t)k*p.b + i)*3*HW, t, y, x);
            if (p.use_min) { if (ek < er) { er = ek; sel = k; } }   // first index wins ties (torch.min)
            else er += ek;
        }
        if (!p.use_min) er /= (float)p.n;
        if (p.use_automask) {
            float es = p.use_min ? INFINITY : 0.f;
            for (int k = 0; k < p.n; ++k) {
                const float ek = photo_at(p, p.src + ((size_t)k*p.b + i)*3*HW, t, y, x);
                es = p.use_min ? fminf(es, ek) : es + ek;
            }
            if (!p.use_min) es /= (float)p.n;
            const size_t idx = (size_t)i*HW + q;
            if (p.noise) es = fmaf(STV_EPS32, __ldg(p.noise + idx), es);
            else if (p.seed) es = fmaf(STV_EPS32, hash_normal(p.seed + (p.step ? *p.step : 0ull), idx), es);
            if (!(er <= es)) { er = es; sel = STV_SEL_STATIC; }
        }
        p.sel[(size_t)i*HW + q] = (uint8_t)sel;
        if (p.err) p.err[(size_t)i*HW + q] = er;
        e = er;
    }
    e = block_sum(e, red);
    if (threadIdx.x == 0) p.partial[(size_t)blockIdx.y*gridDim.x + blockIdx.x] = e;
}

__global__ void recon_reduce_kernel(const float* __restrict__ partial, int n, double inv_count, float* __restrict__ out,
                                    unsigned long long* step) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) a += (double)partial[q];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x/2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) { *out = (float)(sh[0]*inv_count); if (step) *step += 1ull; }
}

// g_pred[k,i,c,y,x] = g * [ sum_{q in N(p), on(q,k)} m(p,q) w_ssim/3 (A_c(q) + 2 w B_c(q) + t C_c(q)) + on(p,k) w_l1/3 sign(w - t) ]
__global__ void __launch_bounds__(256) recon_bwd_kernel(ReconParams p) {
    const int HW = p.H*p.W, H = p.H, W = p.W;
    const int i = blockIdx.y, k = blockIdx.z;
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= HW) return;
    const int y = q/W, x = q - y*W;
    float g = __ldg(p.grad_loss)/((float)p.b*(float)HW);
    if (!p.use_min) g /= (float)p.n;
    const float* wp = p.pred + ((size_t)k*p.b + i)*3*HW;
    const float* tp = p.tgt + (size_t)i*3*HW;
    const uint8_t* sel = p.sel_in + (size_t)i*HW;
    float gw[3] = {0.f, 0.f, 0.f};
    if (p.w_ssim > 0.f) {
        for (int dy = -1; dy <= 1; ++dy) {
            const int yc = y + dy;
            if (yc < 0 || yc >= H) continue;
            for (int dx = -1; dx <= 1; ++dx) {
                const int xc = x + dx;
                if (xc < 0 || xc >= W) continue;
                const uint8_t sv = sel[yc*W + xc];
                if (!(sv == k || sv == STV_SEL_MEAN)) continue;
                // multiplicity of p inside the reflect-padded 3x3 window centred on (yc, xc)
                float m = 1.f;
                if (dy != 0 && ((yc == 0 && y == 1) || (yc == H - 1 && y == H - 2))) m *= 2.f;
                if (dx != 0 && ((xc == 0 && x == 1) || (xc == W - 1 && x == W - 2))) m *= 2.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float S1, S2, S3, T1, T2, a, bq, cq;
                    win5(wp + c*HW, tp + c*HW, yc, xc, H, W, S1, S2, S3, T1, T2);
                    ssim_err_grad(S1, S2, S3, T1, T2, a, bq, cq);
                    const float wv = __ldg(wp + c*HW + q), tv = __ldg(tp + c*HW + q);
                    gw[c] = fmaf(m, fmaf(2.f*wv, bq, fmaf(tv, cq, a)), gw[c]);
                }
            }
        }
    }
    const uint8_t sp = sel[q];
    const bool own = sp == k || sp == STV_SEL_MEAN;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = gw[c]*p.w_ssim*(1.f/3.f);
        if (own && p.w_l1 > 0.f) {
            const float df = __ldg(wp + c*HW + q) - __ldg(tp + c*HW + q);
            v += p.w_l1*(1.f/3.f)*(df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
        }
        p.g_pred[((size_t)k*p.b + i)*3*HW + c*HW + q] = g*v;
    }
}

}  // namespace stv

using namespace stv;

static int recon_check(const stv_photo_cfg* c) {
    STV_REQUIRE(c != nullptr, "stv_recon: cfg is NULL");
    STV_REQUIRE(c->b > 0 && c->n > 0 && c->n <= STV_MAX_SUPPORT, "stv_recon: bad b/n (b=%d n=%d)", c->b, c->n);
    STV_REQUIRE(c->H >= 3 && c->W >= 3, "stv_recon: H, W must be >= 3 for reflection padding (H=%d W=%d)", c->H, c->W);
    STV_REQUIRE(c->b <= 65535, "stv_recon: batch too large for one launch");
    return STV_OK;
}

static void recon_fill(ReconParams& p, const stv_photo_cfg* c) {
    p.b = c->b; p.n = c->n; p.H = c->H; p.W = c->W; p.w_ssim = c->w_ssim; p.w_l1 = c->w_l1;
    p.use_min = c->use_min; p.use_automask = c->use_automask; p.seed = c->noise_seed;
}

extern "C" size_t stv_recon_workspace_bytes(const stv_photo_cfg* c) {
    if (recon_check(c) != STV_OK) return 0;
    return (size_t)((c->H*c->W + 255)/256)*c->b*sizeof(float);
}

extern "C" int stv_recon_fwd(const stv_photo_cfg* c, const float* pred, const float* tgt, const float* source, const float* noise,
                             unsigned long long* noise_step, float* loss, uint8_t* sel, float* err, void* ws, size_t ws_bytes,
                             void* stream) {
    if (int rc = recon_check(c)) return rc;
    STV_REQUIRE(pred && tgt && loss && sel, "stv_recon_fwd: NULL pointer");
    STV_REQUIRE(!c->use_automask || source, "stv_recon_fwd: automasking needs the original `source` frames (reconstruction.py:121)");
    if (ws == nullptr || ws_bytes < stv_recon_workspace_bytes(c)) { set_error("stv_recon_fwd: workspace too small"); return STV_E_WORKSPACE; }
    ReconParams p{};
    recon_fill(p, c);
    p.pred = pred; p.tgt = tgt; p.src = source; p.noise = noise; p.step = noise_step; p.sel = sel; p.err = err; p.partial = (float*)ws;
    const int blocks = (c->H*c->W + 255)/256;
    recon_fwd_kernel<<<dim3(blocks, c->b), 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    if (int rc = check_launch("recon_fwd_kernel")) return rc;
    recon_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p.partial, blocks*c->b, 1.0/((double)c->b*c->H*c->W), loss,
                                                             (c->use_automask && !noise && c->noise_seed) ? noise_step : nullptr);
    count_launch();
    return check_launch("recon_reduce_kernel");
}

extern "C" int stv_recon_bwd(const stv_photo_cfg* c, const float* pred, const float* tgt, const uint8_t* sel, const float* grad_loss,
                             float* g_pred, void* stream) {
    if (int rc = recon_check(c)) return rc;
    STV_REQUIRE(pred && tgt && sel && grad_loss && g_pred, "stv_recon_bwd: NULL pointer");
    ReconParams p{};
    recon_fill(p, c);
    p.pred = pred; p.tgt = tgt; p.sel_in = sel; p.grad_loss = grad_loss; p.g_pred = g_pred;
    recon_bwd_kernel<<<dim3((c->H*c->W + 255)/256, c->b, c->n), 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_launch("recon_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Extended variant: the whole registered class — any channel count (feat_recon hands C-channel feature
// maps, src/core/handlers.py:70-119), loss_name in {ssim, l1, l2} (src/losses/photometric.py:12-23, 54-88) and the
// explainability / uncertainty weighting masks (src/losses/reconstruction.py:46-57), differentiable in `pred` AND `mask`.
//   e_k   = photo(pred_k, target)                                   per support frame k
//   e'_k  = e_k m_k (explainability) | e_k exp(-m_k) + m_k (uncertainty) | e_k
//   err   = min_k e'_k (first index on ties) | mean_k e'_k;   static likewise from the un-warped frames, + eps * noise
//   loss  = mean over (b, H, W) of min(err, static)  (err wins ties)
// `sel` byte: k = warped frame k carries the pixel; 0x40 = mean over the frames; bit 7 = the static error won (automasked), the
// low bits then name the static frame whose masked error was the minimum (it still feeds the mask gradient).
// Same one-thread-per-pixel structure as above: a cold path, L1/L2-resident re-reads instead of tiles.
// ---------------------------------------------------------------------------------------------------------------------
namespace stv {

struct ReconExParams {
    int b, n, C, H, W, loss, use_min, use_automask, mask_mode;
    uint64_t seed;
    const unsigned long long* step;
    const float *pred, *tgt, *src, *mask, *noise, *grad_loss;
    const uint8_t* sel_in;
    float *err, *partial, *g_pred, *g_mask;
    uint8_t* sel;
};

__device__ __forceinline__ float rex_photo(const ReconExParams& p, const float* __restrict__ w, const float* __restrict__ t, int y, int x) {
    const int HW = p.H*p.W, q = y*p.W + x;
    if (p.loss == STV_RECON_L2) {
        float a = 0.f;
        for (int c = 0; c < p.C; ++c) { const float d = __ldg(w + (size_t)c*HW + q) - __ldg(t + (size_t)c*HW + q); a = fmaf(d, d, a); }
        return sqrtf(fmaxf(a, STV_EPS32));
    }
    const float ws = p.loss == STV_RECON_SSIM ? 0.85f : 0.f, wl = p.loss == STV_RECON_SSIM ? 0.15f : 1.f;
    float es = 0.f, el = 0.f;
    for (int c = 0; c < p.C; ++c) {
        if (ws > 0.f) {
            float S1, S2, S3, T1, T2;
            win5(w + (size_t)c*HW, t + (size_t)c*HW, y, x, p.H, p.W, S1, S2, S3, T1, T2);
            es += ssim_err(S1, S2, S3, T1, T2);
        }
        el += fabsf(__ldg(w + (size_t)c*HW + q) - __ldg(t + (size_t)c*HW + q));
    }
    return (ws*es + wl*el)/(float)p.C;
}

__device__ __forceinline__ float rex_masked(int mode, float e, float m) {
    return mode == STV_RECON_MASK_EXPLAIN ? e*m : (mode == STV_RECON_MASK_UNCERT ? fmaf(e, __expf(-m), m) : e);
}
// d e'/d e and d e'/d m
__device__ __forceinline__ float rex_dfe(int mode, float m) { return mode == STV_RECON_MASK_EXPLAIN ? m : (mode == STV_RECON_MASK_UNCERT ? __expf(-m) : 1.f); }
__device__ __forceinline__ float rex_dfm(int mode, float e, float m) { return mode == STV_RECON_MASK_EXPLAIN ? e : (mode == STV_RECON_MASK_UNCERT ? 1.f - e*__expf(-m) : 0.f); }

__device__ __forceinline__ float rex_reduce(const ReconExParams& p, const float* __restrict__ frames, const float* __restrict__ t, int i,
                                            int y, int x, int& sel) {
    const int HW = p.H*p.W, q = y*p.W + x;
    float er = p.use_min ? INFINITY : 0.f;
    sel = p.use_min ? 0 : 0x40;
    for (int k = 0; k < p.n; ++k) {
        float ek = rex_photo(p, frames + ((size_t)k*p.b + i)*p.C*HW, t, y, x);
        if (p.mask_mode) ek = rex_masked(p.mask_mode, ek, __ldg(p.mask + ((size_t)i*p.n + k)*HW + q));
        if (p.use_min) { if (ek < er) { er = ek; sel = k; } }   // first index wins ties (torch.min)
        else er += ek;
    }
    return p.use_min ? er : er/(float)p.n;
}

__global__ void __launch_bounds__(256) recon_ex_fwd_kernel(ReconExParams p) {
    __shared__ float red[32];
    const int HW = p.H*p.W;
    const int i = blockIdx.y;
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    float e = 0.f;
    if (q < HW) {
        const int y = q/p.W, x = q - y*p.W;
        const float* t = p.tgt + (size_t)i*p.C*HW;
        int sel;
        float er = rex_reduce(p, p.pred, t, i, y, x, sel);
        if (p.use_automask) {
            int ssel;
            float es = rex_reduce(p, p.src, t, i, y, x, ssel);
            const size_t idx = (size_t)i*HW + q;
            if (p.noise) es = fmaf(STV_EPS32, __ldg(p.noise + idx), es);
            else if (p.seed) es = fmaf(STV_EPS32, hash_normal(p.seed + (p.step ? *p.step : 0ull), idx), es);
            if (!(er <= es)) { er = es; sel = 0x80 | ssel; }
        }
        p.sel[(size_t)i*HW + q] = (uint8_t)sel;
        if (p.err) p.err[(size_t)i*HW + q] = er;
        e = er;
    }
    e = block_sum(e, red);
    if (threadIdx.x == 0) p.partial[(size_t)blockIdx.y*gridDim.x + blockIdx.x] = e;
}

// Weight of frame k in the reduced error of centre pixel (sel byte sv): warped frames only (static pixels give `pred` nothing).
__device__ __forceinline__ float rex_share(uint8_t sv, int k, int n) {
    if (sv & 0x80) return 0.f;
    return sv == 0x40 ? 1.f/(float)n : (sv == k ? 1.f : 0.f);
}

__global__ void __launch_bounds__(256) recon_ex_bwd_kernel(ReconExParams p) {
    const int HW = p.H*p.W, H = p.H, W = p.W;
    const int i = blockIdx.y, k = blockIdx.z;
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= HW) return;
    const int y = q/W, x = q - y*W;
    const float g = __ldg(p.grad_loss)/((float)p.b*(float)HW);
    const float* wp = p.pred + ((size_t)k*p.b + i)*p.C*HW;
    const float* tp = p.tgt + (size_t)i*p.C*HW;
    const uint8_t* sel = p.sel_in + (size_t)i*HW;
    const float* mk = p.mask_mode ? p.mask + ((size_t)i*p.n + k)*HW : nullptr;
    float* gp = p.g_pred + ((size_t)k*p.b + i)*p.C*HW;
    // pointwise factor of this pixel's own error, window factors of the up-to-nine centres whose 3x3 window holds it
    const float own = rex_share(sel[q], k, p.n)*(mk ? rex_dfe(p.mask_mode, __ldg(mk + q)) : 1.f);
    if (p.loss == STV_RECON_L2) {
        float a = 0.f;
        for (int c = 0; c < p.C; ++c) { const float d = __ldg(wp + (size_t)c*HW + q) - __ldg(tp + (size_t)c*HW + q); a = fmaf(d, d, a); }
        const float inv = a >= STV_EPS32 ? own*g*rsqrtf(a) : 0.f;   // clamp(min=eps) passes the gradient at and above the bound
        for (int c = 0; c < p.C; ++c) gp[(size_t)c*HW + q] = inv*(__ldg(wp + (size_t)c*HW + q) - __ldg(tp + (size_t)c*HW + q));
        return;
    }
    const float ws = p.loss == STV_RECON_SSIM ? 0.85f : 0.f, wl = p.loss == STV_RECON_SSIM ? 0.15f : 1.f;
    float fac[9], mul[9];
    bool any = false;
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const int dy = j/3 - 1, dx = j % 3 - 1, yc = y + dy, xc = x + dx;
        fac[j] = 0.f; mul[j] = 1.f;
        if (ws == 0.f || yc < 0 || yc >= H || xc < 0 || xc >= W) continue;
        const int qc = yc*W + xc;
        fac[j] = rex_share(sel[qc], k, p.n)*(mk ? rex_dfe(p.mask_mode, __ldg(mk + qc)) : 1.f);
        // multiplicity of this pixel inside the reflect-padded 3x3 window centred on (yc, xc)
        if (dy != 0 && ((yc == 0 && y == 1) || (yc == H - 1 && y == H - 2))) mul[j] *= 2.f;
        if (dx != 0 && ((xc == 0 && x == 1) || (xc == W - 1 && x == W - 2))) mul[j] *= 2.f;
        any = any || fac[j] != 0.f;
    }
    for (int c = 0; c < p.C; ++c) {
        const float* wc = wp + (size_t)c*HW; const float* tc = tp + (size_t)c*HW;
        const float wv = __ldg(wc + q), tv = __ldg(tc + q);
        float gs = 0.f;
        if (any) {
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                if (fac[j] == 0.f) continue;
                float S1, S2, S3, T1, T2, a, bq, cq;
                win5(wc, tc, y + j/3 - 1, x + j % 3 - 1, H, W, S1, S2, S3, T1, T2);
                ssim_err_grad(S1, S2, S3, T1, T2, a, bq, cq);
                gs = fmaf(fac[j]*mul[j], fmaf(2.f*wv, bq, fmaf(tv, cq, a)), gs);
            }
        }
        const float df = wv - tv;
        const float gl = own*(df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
        gp[(size_t)c*HW + q] = g*(ws*gs + wl*gl)/(float)p.C;
    }
}

// g_mask[i,k,q] = g * share * d e'_k/d m_k, the error being the warped one where a warped frame carries the pixel and the
// static one (un-warped frame k) where the automask chose the static branch.
__global__ void __launch_bounds__(256) recon_ex_mask_bwd_kernel(ReconExParams p) {
    const int HW = p.H*p.W;
    const int i = blockIdx.y, k = blockIdx.z;
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= HW) return;
    const int y = q/p.W, x = q - y*p.W;
    const uint8_t sv = p.sel_in[(size_t)i*HW + q];
    const bool stat = (sv & 0x80) != 0;
    const uint8_t lo = sv & 0x7f;
    const float share = lo == 0x40 ? 1.f/(float)p.n : (lo == k ? 1.f : 0.f);
    float out = 0.f;
    if (share != 0.f) {
        const float* fr = (stat ? p.src : p.pred) + ((size_t)k*p.b + i)*p.C*HW;
        const float e = rex_photo(p, fr, p.tgt + (size_t)i*p.C*HW, y, x);
        const float m = __ldg(p.mask + ((size_t)i*p.n + k)*HW + q);
        out = __ldg(p.grad_loss)/((float)p.b*(float)HW)*share*rex_dfm(p.mask_mode, e, m);
    }
    p.g_mask[((size_t)i*p.n + k)*HW + q] = out;
}

}  // namespace stv

static int recon_ex_check(const stv_recon_cfg* c) {
    STV_REQUIRE(c != nullptr, "stv_recon_ex: cfg is NULL");
    STV_REQUIRE(c->b > 0 && c->n > 0 && c->n <= 0x3f && c->C > 0, "stv_recon_ex: bad b/n/C (b=%d n=%d C=%d)", c->b, c->n, c->C);
    STV_REQUIRE(c->H >= 3 && c->W >= 3, "stv_recon_ex: H, W must be >= 3 for reflection padding (H=%d W=%d)", c->H, c->W);
    STV_REQUIRE(c->b <= 65535 && c->n <= 65535, "stv_recon_ex: batch too large for one launch");
    STV_REQUIRE(c->loss == STV_RECON_SSIM || c->loss == STV_RECON_L1 || c->loss == STV_RECON_L2, "stv_recon_ex: bad loss %d", c->loss);
    STV_REQUIRE(c->mask_mode >= STV_RECON_MASK_NONE && c->mask_mode <= STV_RECON_MASK_UNCERT, "Invalid mask type: %d", c->mask_mode);
    STV_REQUIRE((long long)c->C*c->H*c->W < (1ll << 31), "stv_recon_ex: frame too large");
    return STV_OK;
}

static void recon_ex_fill(ReconExParams& p, const stv_recon_cfg* c) {
    p.b = c->b; p.n = c->n; p.C = c->C; p.H = c->H; p.W = c->W; p.loss = c->loss; p.use_min = c->use_min;
    p.use_automask = c->use_automask; p.mask_mode = c->mask_mode; p.seed = c->noise_seed;
}

extern "C" size_t stv_recon_ex_workspace_bytes(const stv_recon_cfg* c) {
    if (recon_ex_check(c) != STV_OK) return 0;
    return (size_t)((c->H*c->W + 255)/256)*c->b*sizeof(float);
}

extern "C" int stv_recon_ex_fwd(const stv_recon_cfg* c, const float* pred, const float* tgt, const float* source, const float* mask,
                                const float* noise, unsigned long long* noise_step, float* loss, uint8_t* sel, float* err, void* ws,
                                size_t ws_bytes, void* stream) {
    if (int rc = recon_ex_check(c)) return rc;
    STV_REQUIRE(pred && tgt && loss && sel, "stv_recon_ex_fwd: NULL pointer");
    STV_REQUIRE(!c->use_automask || source, "stv_recon_ex_fwd: automasking needs the original `source` frames (reconstruction.py:121)");
    STV_REQUIRE(!c->mask_mode || mask, "Must provide a 'mask' when masking... (reconstruction.py:53)");
    if (ws == nullptr || ws_bytes < stv_recon_ex_workspace_bytes(c)) { set_error("stv_recon_ex_fwd: workspace too small"); return STV_E_WORKSPACE; }
    ReconExParams p{};
    recon_ex_fill(p, c);
    p.pred = pred; p.tgt = tgt; p.src = source; p.mask = mask; p.noise = noise; p.step = noise_step; p.sel = sel; p.err = err;
    p.partial = (float*)ws;
    const int blocks = (c->H*c->W + 255)/256;
    recon_ex_fwd_kernel<<<dim3(blocks, c->b), 256, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    if (int rc = check_launch("recon_ex_fwd_kernel")) return rc;
    recon_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p.partial, blocks*c->b, 1.0/((double)c->b*c->H*c->W), loss,
                                                             (c->use_automask && !noise && c->noise_seed) ? noise_step : nullptr);
    count_launch();
    return check_launch("recon_reduce_kernel");
}

extern "C" int stv_recon_ex_bwd(const stv_recon_cfg* c, const float* pred, const float* tgt, const float* source, const float* mask,
                                const uint8_t* sel, const float* grad_loss, float* g_pred, float* g_mask, void* stream) {
    if (int rc = recon_ex_check(c)) return rc;
    STV_REQUIRE(pred && tgt && sel && grad_loss && (g_pred || g_mask), "stv_recon_ex_bwd: NULL pointer");
    STV_REQUIRE(!c->mask_mode || mask, "stv_recon_ex_bwd: mask_mode set without a mask");
    STV_REQUIRE(!g_mask || (c->mask_mode && (!c->use_automask || source)), "stv_recon_ex_bwd: g_mask needs mask_mode (and `source` when automasking)");
    ReconExParams p{};
    recon_ex_fill(p, c);
    p.pred = pred; p.tgt = tgt; p.src = source; p.mask = mask; p.sel_in = sel; p.grad_loss = grad_loss; p.g_pred = g_pred; p.g_mask = g_mask;
    const dim3 grid((c->H*c->W + 255)/256, c->b, c->n);
    if (g_pred) {
        recon_ex_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
        count_launch();
        if (int rc = check_launch("recon_ex_bwd_kernel")) return rc;
    }
    if (g_mask) {
        recon_ex_mask_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
        count_launch();
        if (int rc = check_launch("recon_ex_mask_bwd_kernel")) return rc;
    }
    return STV_OK;
}
