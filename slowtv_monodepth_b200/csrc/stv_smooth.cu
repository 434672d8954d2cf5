// Edge-aware disparity smoothness (rows 15-16 of SURVEY 8a), all scales of a batch in three launches per direction.
//
//   pass A  per-(scale, image) sums of the disparity, NCHUNK partial sums each (fixed order -> deterministic mean)
//   pass B  per pixel: mean-normalised disparity gradients, bilinear-downsampled image gradients, exp(-|dI|) weights,
//           per-block partial sums of the weighted gradients, optional logging maps of the first scale
//   pass C  tiny finalize: loss = mean_s(loss_s / 2**s) and the per-(scale, image) stats the backward needs
// The backward uses Euler's identity for the 1-homogeneous loss, sum_q dL/dd^(q) d^(q) = L, so the mean-normalisation
// term needs no second reduction: g_disp(p) = dL/dd^(p)/m - [mean > eps] * g_s * L_i / (m * h*w).
#include "stv_common.cuh"

namespace stv {

constexpr int NCHUNK = 32;
constexpr int SM_NT = 256;

struct SmoothParams {
    int b, S, H, W, use_edges;
    int h[STV_MAX_SCALES], w[STV_MAX_SCALES];
    float scale_div[STV_MAX_SCALES];
    const float* disp[STV_MAX_SCALES];
    float* g_disp[STV_MAX_SCALES];
    int blk_off[STV_MAX_SCALES + 1];  // pass-B blocks per image for each scale, prefix sums
    const float* img;
    const float* grad_loss;
    float *sum_part;   // [S][b][NCHUNK]
    float *loss_part;  // [S][b][max blocks per image]
    float *stats;      // [S][b][2] = {mean, loss_sum}
    int max_blk;
    float *disp_grad, *image_grad;
};

__global__ void __launch_bounds__(SM_NT) smooth_sum_kernel(SmoothParams p) {
    __shared__ float red[32];
    const int chunk = blockIdx.x, i = blockIdx.y, s = blockIdx.z;
    const int hw = p.h[s]*p.w[s];
    const int per = (hw + NCHUNK - 1)/NCHUNK;
    const int lo = chunk*per, hi = min(hw, lo + per);
    const float* d = p.disp[s] + (size_t)i*hw;
    float a = 0.f;
    for (int q = lo + threadIdx.x; q < hi; q += SM_NT) a += __ldg(d + q);
    a = block_sum(a, red);
    if (threadIdx.x == 0) p.sum_part[((size_t)s*p.b + i)*NCHUNK + chunk] = a;
}

__device__ __forceinline__ float mean_from_parts(const float* part, int hw) {
    double a = 0.0;
#pragma unroll 8
    for (int q = 0; q < NCHUNK; ++q) a += (double)part[q];
    return (float)(a/(double)hw);
}

// Bilinear (align_corners=False) sample of the 3-channel image at output pixel (y, x) of an (h, w) grid: the target image
// resized to the disparity's resolution (handlers.py:278). Identity when (h, w) == (H, W).
__device__ __forceinline__ void img_at(const float* __restrict__ img, int H, int W, int h, int w, int y, int x, float* rgb) {
    const int HW = H*W;
    if (h == H && w == W) {
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] = __ldg(img + c*HW + y*W + x);
        return;
    }
    const float sy = fmaxf(((float)H/(float)h)*((float)y + 0.5f) - 0.5f, 0.f);
    const float sxx = fmaxf(((float)W/(float)w)*((float)x + 0.5f) - 0.5f, 0.f);
    const int y0 = min((int)sy, H - 1), x0 = min((int)sxx, W - 1);
    const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sxx - (float)x0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* q = img + c*HW;
        rgb[c] = (1.f - ly)*((1.f - lx)*__ldg(q + y0*W + x0) + lx*__ldg(q + y0*W + x1)) +
                 ly*((1.f - lx)*__ldg(q + y1*W + x0) + lx*__ldg(q + y1*W + x1));
    }
}

// Edge weights of pixel (y, x): exp(-mean_c |I(y,x) - I(y,x+1)|), exp(-mean_c |I(y,x) - I(y+1,x)|); also raw means.
__device__ __forceinline__ void edge_terms(const SmoothParams& p, const float* img, int s, int y, int x, float& ix, float& iy) {
    const int h = p.h[s], w = p.w[s];
    float c0[3], cx[3], cy[3];
    img_at(img, p.H, p.W, h, w, y, x, c0);
    ix = iy = 0.f;
    if (x + 1 < w) {
        img_at(img, p.H, p.W, h, w, y, x + 1, cx);
        ix = (fabsf(c0[0] - cx[0]) + fabsf(c0[1] - cx[1]) + fabsf(c0[2] - cx[2]))*(1.f/3.f);
    }
    if (y + 1 < h) {
        img_at(img, p.H, p.W, h, w, y + 1, x, cy);
        iy = (fabsf(c0[0] - cy[0]) + fabsf(c0[1] - cy[1]) + fabsf(c0[2] - cy[2]))*(1.f/3.f);
    }
}

// grid = (total pass-B blocks over scales, b). Each block covers SM_NT consecutive pixels of one (scale, image).
__global__ void __launch_bounds__(SM_NT) smooth_fwd_kernel(SmoothParams p) {
    __shared__ float red[32];
    int s = 0;
    while (s + 1 < p.S && (int)blockIdx.x >= p.blk_off[s + 1]) ++s;
    const int blk = blockIdx.x - p.blk_off[s], i = blockIdx.y;
    const int h = p.h[s], w = p.w[s], hw = h*w;
    __shared__ float mean_sh;   // one thread adds the 32 partial sums (in double) for the block instead of all 256
    if (threadIdx.x == 0) mean_sh = mean_from_parts(p.sum_part + ((size_t)s*p.b + i)*NCHUNK, hw);
    __syncthreads();
    const float mean = mean_sh;
    const float inv_m = 1.0f/fmaxf(mean, STV_EPS32);
    const int q = blk*SM_NT + threadIdx.x;
    float v = 0.f;
    if (q < hw) {
        const int y = q/w, x = q - y*w;
        const float* d = p.disp[s] + (size_t)i*hw;
        // __fmul_rn: every normalised value is rounded once, identically in every thread (an FMA-contracted a*m - b*m would
        // turn exact ties into rounding noise of random sign).
        const float d0 = __fmul_rn(__ldg(d + q), inv_m);
        const float dx = x + 1 < w ? fabsf(d0 - __fmul_rn(__ldg(d + q + 1), inv_m)) : 0.f;
        const float dy = y + 1 < h ? fabsf(d0 - __fmul_rn(__ldg(d + q + w), inv_m)) : 0.f;
        float ix, iy;
        edge_terms(p, p.img + (size_t)i*3*p.H*p.W, s, y, x, ix, iy);
        v = p.use_edges ? dx*expf(-ix) + dy*expf(-iy) : dx + dy;
        if (s == 0 && p.disp_grad) p.disp_grad[(size_t)i*hw + q] = sqrtf(fmaxf(dx*dx + dy*dy, STV_EPS32));
        if (s == 0 && p.image_grad) p.image_grad[(size_t)i*hw + q] = sqrtf(fmaxf(ix*ix + iy*iy, STV_EPS32));
    }
    v = block_sum(v, red);
    if (threadIdx.x == 0) p.loss_part[((size_t)s*p.b + i)*p.max_blk + blk] = v;
}

__global__ void smooth_finalize_kernel(SmoothParams p, float* __restrict__ loss) {
    // One WARP per (s, i) sums that image's block partials (lanes stride over them, fixed-order shuffle tree in double: the result
    // does not depend on the launch) and computes the stats; thread 0 then combines (tiny: S*b <= a few hundred).
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = wid; e < p.S*p.b; e += nw) {
        const int s = e/p.b;
        const int nb = p.blk_off[s + 1] - p.blk_off[s];
        double a = 0.0;
        for (int q = lane; q < nb; q += 32) a += (double)p.loss_part[(size_t)e*p.max_blk + q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            p.stats[e*2 + 0] = mean_from_parts(p.sum_part + (size_t)e*NCHUNK, p.h[s]*p.w[s]);
            p.stats[e*2 + 1] = (float)a;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int s = 0; s < p.S; ++s) {
            double a = 0.0;
            for (int i = 0; i < p.b; ++i) a += (double)p.stats[(s*p.b + i)*2 + 1];
            tot += a/((double)p.b*p.h[s]*p.w[s])/(double)p.scale_div[s];
        }
        *loss = (float)(tot/(double)p.S);
    }
}

__global__ void __launch_bounds__(SM_NT) smooth_bwd_kernel(SmoothParams p) {
    int s = 0;
    while (s + 1 < p.S && (int)blockIdx.x >= p.blk_off[s + 1]) ++s;
    const int blk = blockIdx.x - p.blk_off[s], i = blockIdx.y;
    const int h = p.h[s], w = p.w[s], hw = h*w;
    const int q = blk*SM_NT + threadIdx.x;
    if (q >= hw) return;
    const float mean = p.stats[(s*p.b + i)*2 + 0], Li = p.stats[(s*p.b + i)*2 + 1];
    const float m = fmaxf(mean, STV_EPS32), inv_m = 1.0f/m;
    const float gs = __ldg(p.grad_loss)/((float)p.S*p.scale_div[s]*(float)p.b*(float)hw);
    const int y = q/w, x = q - y*w;
    const float* d = p.disp[s] + (size_t)i*hw;
    const float* img = p.img + (size_t)i*3*p.H*p.W;
    const float d0 = __fmul_rn(__ldg(d + q), inv_m);
    auto sgn = [](float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); };
    float g = 0.f;
    // the four edges of this pixel: weights exp(-mean_c |I(a) - I(b)|) from FIVE image samples (centre and its four neighbours; the
    // same expression, in the same operand order, as edge_terms evaluates for the forward)
    float c0[3] = {0.f, 0.f, 0.f}, cn[3];
    if (p.use_edges) img_at(img, p.H, p.W, h, w, y, x, c0);
    auto weight = [&](int yy, int xx, bool centre_first) -> float {
        if (!p.use_edges) return 1.f;
        img_at(img, p.H, p.W, h, w, yy, xx, cn);
        const float a = centre_first ? (fabsf(c0[0] - cn[0]) + fabsf(c0[1] - cn[1]) + fabsf(c0[2] - cn[2]))
                                     : (fabsf(cn[0] - c0[0]) + fabsf(cn[1] - c0[1]) + fabsf(cn[2] - c0[2]));
        return expf(-a*(1.f/3.f));
    };
    if (x + 1 < w) g += sgn(d0 - __fmul_rn(__ldg(d + q + 1), inv_m))*weight(y, x + 1, true);
    if (y + 1 < h) g += sgn(d0 - __fmul_rn(__ldg(d + q + w), inv_m))*weight(y + 1, x, true);
    if (x > 0) g -= sgn(__fmul_rn(__ldg(d + q - 1), inv_m) - d0)*weight(y, x - 1, false);
    if (y > 0) g -= sgn(__fmul_rn(__ldg(d + q - w), inv_m) - d0)*weight(y - 1, x, false);
    float out = gs*g*inv_m;
    if (mean >= STV_EPS32) out -= gs*Li*inv_m/(float)hw;  // d/d mean through the normalisation (clamp passes at equality)
    p.g_disp[s][(size_t)i*hw + q] = out;
}

// Fused AdamW over one flat buffer (torch.optim.AdamW semantics, decoupled weight decay).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                    float* __restrict__ m, float* __restrict__ v, size_t n, size_t n_decay,
                                                    float lr, float b1, float b2, float eps, float wd, float gscale,
                                                    float inv_bc1, float inv_sqrt_bc2) {
    const size_t stride = (size_t)gridDim.x*blockDim.x;
    // 128-bit streams over the 16-byte aligned part (every tensor of the flat buffers starts on a 16-byte boundary; n_decay is a
    // multiple of 4 there), scalar tail otherwise.
    const bool vec = ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0) && (n_decay & 3) == 0;
    const size_t n4 = vec ? n/4 : 0;
    for (size_t q = (size_t)blockIdx.x*blockDim.x + threadIdx.x; q < n4; q += stride) {
        float4 g4 = reinterpret_cast<const float4*>(grad)[q], p4 = reinterpret_cast<float4*>(param)[q];
        float4 m4 = reinterpret_cast<float4*>(m)[q], v4 = reinterpret_cast<float4*>(v)[q];
        const float dec = (q*4 < n_decay) ? 1.f - lr*wd : 1.f;
        float* gp = &g4.x; float* pp = &p4.x; float* mp = &m4.x; float* vp = &v4.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float g = gp[e]*gscale;
            const float pv = pp[e]*dec;
            const float mq = fmaf(b1, mp[e], (1.f - b1)*g);
            const float vq = fmaf(b2, vp[e], (1.f - b2)*g*g);
            mp[e] = mq; vp[e] = vq;
            pp[e] = pv - lr*inv_bc1*(mq/(sqrtf(vq)*inv_sqrt_bc2 + eps));
        }
        reinterpret_cast<float4*>(param)[q] = p4; reinterpret_cast<float4*>(m)[q] = m4; reinterpret_cast<float4*>(v)[q] = v4;
    }
    for (size_t q = n4*4 + (size_t)blockIdx.x*blockDim.x + threadIdx.x; q < n; q += stride) {
        const float g = grad[q]*gscale;
        float pv = param[q];
        if (q < n_decay) pv *= 1.f - lr*wd;
        const float mq = fmaf(b1, m[q], (1.f - b1)*g);
        const float vq = fmaf(b2, v[q], (1.f - b2)*g*g);
        m[q] = mq; v[q] = vq;
        const float denom = sqrtf(vq)*inv_sqrt_bc2 + eps;
        param[q] = pv - lr*inv_bc1*(mq/denom);
    }
}

}  // namespace stv

using namespace stv;

static int smooth_setup(const stv_smooth_cfg* c, SmoothParams& p, void* ws, size_t ws_bytes, const char* who) {
    STV_REQUIRE(c != nullptr, "%s: cfg is NULL", who);
    STV_REQUIRE(c->b > 0 && c->S > 0 && c->S <= STV_MAX_SCALES && c->H > 0 && c->W > 0, "%s: bad shape", who);
    STV_REQUIRE(c->b <= 65535, "%s: batch too large", who);
    p.b = c->b; p.S = c->S; p.H = c->H; p.W = c->W; p.use_edges = c->use_edges;
    p.blk_off[0] = 0;
    p.max_blk = 0;
    for (int s = 0; s < c->S; ++s) {
        STV_REQUIRE(c->h[s] > 0 && c->w[s] > 0 && c->scale_div[s] > 0.f, "%s: bad scale %d (h=%d w=%d div=%g)", who, s, c->h[s], c->w[s], c->scale_div[s]);
        p.h[s] = c->h[s]; p.w[s] = c->w[s]; p.scale_div[s] = c->scale_div[s];
        const int nb = (c->h[s]*c->w[s] + SM_NT - 1)/SM_NT;
        p.blk_off[s + 1] = p.blk_off[s] + nb;
        p.max_blk = nb > p.max_blk ? nb : p.max_blk;
    }
    const size_t need = stv_smooth_workspace_bytes(c);
    if (!ws || ws_bytes < need) {
        set_error("%s: workspace too small (%zu < %zu bytes)", who, ws_bytes, need);
        return STV_E_WORKSPACE;
    }
    float* f = (float*)ws;
    p.stats = f; f += (size_t)c->S*c->b*2;
    p.sum_part = f; f += (size_t)c->S*c->b*NCHUNK;
    p.loss_part = f;
    return STV_OK;
}

extern "C" size_t stv_smooth_workspace_bytes(const stv_smooth_cfg* c) {
    if (!c || c->S <= 0 || c->S > STV_MAX_SCALES || c->b <= 0) return 0;
    int max_blk = 0;
    for (int s = 0; s < c->S; ++s) {
        const int nb = (c->h[s]*c->w[s] + SM_NT - 1)/SM_NT;
        max_blk = nb > max_blk ? nb : max_blk;
    }
    return ((size_t)c->S*c->b*(2 + NCHUNK + max_blk))*sizeof(float);
}

extern "C" int stv_smooth_fwd(const stv_smooth_cfg* c, const float* const* disp, const float* img, float* loss,
                              float* disp_grad, float* image_grad, void* ws, size_t ws_bytes, void* stream) {
    SmoothParams p{};
    if (int rc = smooth_setup(c, p, ws, ws_bytes, "stv_smooth_fwd")) return rc;
    STV_REQUIRE(disp && img && loss, "stv_smooth_fwd: NULL pointer");
    for (int s = 0; s < c->S; ++s) { STV_REQUIRE(disp[s], "stv_smooth_fwd: disp[%d] is NULL", s); p.disp[s] = disp[s]; }
    p.img = img; p.disp_grad = disp_grad; p.image_grad = image_grad;
    cudaStream_t st = (cudaStream_t)stream;
    smooth_sum_kernel<<<dim3(NCHUNK, c->b, c->S), SM_NT, 0, st>>>(p);
    count_launch();
    if (int rc = check_launch("smooth_sum_kernel")) return rc;
    smooth_fwd_kernel<<<dim3(p.blk_off[c->S], c->b), SM_NT, 0, st>>>(p);
    count_launch();
    if (int rc = check_launch("smooth_fwd_kernel")) return rc;
    smooth_finalize_kernel<<<1, 1024, 0, st>>>(p, loss);
    count_launch();
    return check_launch("smooth_finalize_kernel");
}

extern "C" int stv_smooth_bwd(const stv_smooth_cfg* c, const float* const* disp, const float* img, const float* grad_loss,
                              float* const* g_disp, void* ws, size_t ws_bytes, void* stream) {
    SmoothParams p{};
    if (int rc = smooth_setup(c, p, ws, ws_bytes, "stv_smooth_bwd")) return rc;
    STV_REQUIRE(disp && img && grad_loss && g_disp, "stv_smooth_bwd: NULL pointer");
    for (int s = 0; s < c->S; ++s) {
        STV_REQUIRE(disp[s] && g_disp[s], "stv_smooth_bwd: disp/g_disp[%d] is NULL", s);
        p.disp[s] = disp[s]; p.g_disp[s] = g_disp[s];
    }
    p.img = img; p.grad_loss = grad_loss;
    smooth_bwd_kernel<<<dim3(p.blk_off[c->S], c->b), SM_NT, 0, (cudaStream_t)stream>>>(p);
    count_launch();
    return check_launch("smooth_bwd_kernel");
}

extern "C" int stv_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, size_t n_decay,
                              float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, int step,
                              void* stream) {
    STV_REQUIRE(param && grad && exp_avg && exp_avg_sq, "stv_adamw_step: NULL pointer");
    STV_REQUIRE(step >= 1, "stv_adamw_step: step must be >= 1 (got %d)", step);
    STV_REQUIRE(n_decay <= n, "stv_adamw_step: n_decay > n");
    if (n == 0) return STV_OK;
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    size_t blocks = (n/4 + 255)/256 + 1;
    if (blocks > 148*16) blocks = 148*16;
    adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, n_decay, lr, beta1, beta2, eps,
                                                                    weight_decay, grad_scale, (float)(1.0/bc1), (float)(1.0/sqrt(bc2)));
    count_launch();
    return check_launch("adamw_kernel");
}
