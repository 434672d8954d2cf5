// SmoothReg with every constructor flag (src/regularizers/smooth.py:51-97): first- or second-order (use_laplacian) absolute
// gradients, optional 3x3 Gaussian pre-blur (use_blur; kornia.filters.gaussian_blur2d((3,3),(1,1)), reflect border), optional
// edge-aware weights — SURVEY 8f rank 4. The KBR hot path (first order, no blur, all scales at once) is stv_smooth_fwd/bwd; this
// is the general single-scale form behind `SmoothReg.forward` for the other configurations: a handful of one-thread-per-pixel
// passes over (b, H, W) planes, orchestrated inside one C-ABI call, intermediates in the caller's workspace.
//
// Building block  A_axis(x)[p] = | B(x)[p] - B(x)[p + e_axis] |  (0 on the last column / row; B = blur or identity), i.e.
// compute_grad (smooth.py:12-30). Its adjoint for an upstream gradient g:  B^T D^T (sign(.) g).
#include "stv_common.cuh"

namespace stv {

__device__ __forceinline__ float sx_blur_at(const float* __restrict__ x, int y, int xx, int H, int W) {
    // normalised taps exp(-d^2/2), d in {-1,0,1}: (g1, g0, g1), reflect border
    const float e = 0.60653065971263342f, g0 = 1.f/(1.f + 2.f*e), g1 = e*g0;
    float acc = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int ya = min(max(reflect_idx(y + dy, H), 0), H - 1);
        const float wy = dy == 0 ? g0 : g1;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int xa = min(max(reflect_idx(xx + dx, W), 0), W - 1);
            acc = fmaf(wy*(dx == 0 ? g0 : g1), __ldg(x + (size_t)ya*W + xa), acc);
        }
    }
    return acc;
}
__device__ __forceinline__ float sx_val(const float* __restrict__ x, int y, int xx, int H, int W, int blur) {
    return blur ? sx_blur_at(x, y, xx, H, W) : __ldg(x + (size_t)y*W + xx);
}
// signed difference B(x)[p] - B(x)[p+e]; 0 outside the valid range
__device__ __forceinline__ float sx_diff(const float* __restrict__ x, int y, int xx, int H, int W, int axis, int blur) {
    if (axis == 0) { if (xx < 0 || xx >= W - 1) return 0.f; return sx_val(x, y, xx, H, W, blur) - sx_val(x, y, xx + 1, H, W, blur); }
    if (y < 0 || y >= H - 1) return 0.f;
    return sx_val(x, y, xx, H, W, blur) - sx_val(x, y + 1, xx, H, W, blur);
}

// out[plane][p] = A_axis(x[plane])[p]
__global__ void __launch_bounds__(256) sx_absdiff_kernel(int H, int W, int axis, int blur, const float* __restrict__ x, float* __restrict__ out) {
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= H*W) return;
    const size_t pl = (size_t)blockIdx.y*H*W;
    out[pl + q] = fabsf(sx_diff(x + pl, q/W, q % W, H, W, axis, blur));
}

// u = D^T (sign(D B x) g):  u[p] = t[p] - t[p - e],  t = sign(diff) * g (t = 0 where the forward output is the zero border)
__global__ void __launch_bounds__(256) sx_absdiff_bwd_kernel(int H, int W, int axis, int blur, const float* __restrict__ x,
                                                             const float* __restrict__ g, float* __restrict__ u) {
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= H*W) return;
    const size_t pl = (size_t)blockIdx.y*H*W;
    const int y = q/W, xx = q % W;
    auto t_at = [&](int yy, int xc) -> float {
        if (yy < 0 || xc < 0 || yy >= H || xc >= W) return 0.f;
        const float d = sx_diff(x + pl, yy, xc, H, W, axis, blur);
        return d > 0.f ? __ldg(g + pl + (size_t)yy*W + xc) : (d < 0.f ? -__ldg(g + pl + (size_t)yy*W + xc) : 0.f);
    };
    u[pl + q] = t_at(y, xx) - (axis == 0 ? t_at(y, xx - 1) : t_at(y - 1, xx));
}

// out (+)= B^T u: the adjoint of the reflect-padded blur (border taps fold back onto the interior)
__global__ void __launch_bounds__(256) sx_blur_adj_kernel(int H, int W, const float* __restrict__ u, float* __restrict__ out, int accumulate) {
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= H*W) return;
    const size_t pl = (size_t)blockIdx.y*H*W;
    const int y = q/W, xx = q % W;
    const float e = 0.60653065971263342f, g0 = 1.f/(1.f + 2.f*e), g1 = e*g0;
    float acc = 0.f;
    for (int py = max(y - 1, 0); py <= min(y + 1, H - 1); ++py) {
        float wy = 0.f;   // total weight with which output row py reads input row y
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) if (min(max(reflect_idx(py + dy, H), 0), H - 1) == y) wy += dy == 0 ? g0 : g1;
        for (int px = max(xx - 1, 0); px <= min(xx + 1, W - 1); ++px) {
            float wx = 0.f;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) if (min(max(reflect_idx(px + dx, W), 0), W - 1) == xx) wx += dx == 0 ? g0 : g1;
            acc = fmaf(wy*wx, __ldg(u + pl + (size_t)py*W + px), acc);
        }
    }
    out[pl + q] = accumulate ? out[pl + q] + acc : acc;
}

__global__ void __launch_bounds__(256) sx_add_kernel(long long n, const float* __restrict__ a, float* __restrict__ out, int accumulate) {
    const long long q = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (q < n) out[q] = accumulate ? out[q] + a[q] : a[q];
}

// per-image sums: out[i] = sum_p a[i][p] * (b ? b[i][p] : 1)   (one block per image, double accumulation)
__global__ void __launch_bounds__(256) sx_image_dot_kernel(int HW, const float* __restrict__ a, const float* __restrict__ b2, double* __restrict__ out) {
    __shared__ double sh[256];
    const size_t pl = (size_t)blockIdx.x*HW;
    double acc = 0.0;
    for (int q = threadIdx.x; q < HW; q += blockDim.x) acc += (double)a[pl + q]*(b2 ? (double)b2[pl + q] : 1.0);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x/2; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// dn = disp / max(mean, eps)    (ops.mean_normalize, src/tools/ops.py)
__global__ void __launch_bounds__(256) sx_normalize_kernel(int HW, const float* __restrict__ disp, const double* __restrict__ sum, float* __restrict__ dn) {
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= HW) return;
    const float m = fmaxf((float)(sum[blockIdx.y]/(double)HW), STV_EPS32);
    dn[(size_t)blockIdx.y*HW + q] = disp[(size_t)blockIdx.y*HW + q]/m;
}

// g_disp = g_dn / m' - [mean > eps] sum(g_dn * disp) / (m'^2 HW)
__global__ void __launch_bounds__(256) sx_normalize_bwd_kernel(int HW, const float* __restrict__ g_dn, const double* __restrict__ sum,
                                                               const double* __restrict__ dot, float* __restrict__ g_disp) {
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= HW) return;
    const float mean = (float)(sum[blockIdx.y]/(double)HW), m = fmaxf(mean, STV_EPS32);
    const float corr = mean >= STV_EPS32 ? (float)(dot[blockIdx.y]/((double)m*(double)m*(double)HW)) : 0.f;
    g_disp[(size_t)blockIdx.y*HW + q] = g_dn[(size_t)blockIdx.y*HW + q]/m - corr;
}

// channel mean of the image gradient planes: out[i][p] = mean_c a[i][c][p]
__global__ void __launch_bounds__(256) sx_chmean_kernel(int HW, int C, const float* __restrict__ a, float* __restrict__ out) {
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= HW) return;
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += a[((size_t)blockIdx.y*C + c)*HW + q];
    out[(size_t)blockIdx.y*HW + q] = acc/(float)C;
}

// loss partials + the two logging maps; bwd == 1 instead writes the upstream gradients of ddx / ddy
__global__ void __launch_bounds__(256) sx_loss_kernel(int HW, int use_edges, const float* __restrict__ ddx, const float* __restrict__ ddy,
                                                      const float* __restrict__ ix, const float* __restrict__ iy, float* __restrict__ partial,
                                                      float* __restrict__ disp_grad, float* __restrict__ image_grad) {
    __shared__ float red[32];
    const int q = blockIdx.x*blockDim.x + threadIdx.x;
    const size_t o = (size_t)blockIdx.y*HW + q;
    float v = 0.f;
    if (q < HW) {
        const float ax = ddx[o], ay = ddy[o], bx = ix[o], by = iy[o];
        v = use_edges ? ax*__expf(-bx) + ay*__expf(-by) : ax + ay;
        if (disp_grad) disp_grad[o] = sqrtf(fmaxf(ax*ax + ay*ay, STV_EPS32));
        if (image_grad) image_grad[o] = sqrtf(fmaxf(bx*bx + by*by, STV_EPS32));
    }
    v = block_sum(v, red);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y*gridDim.x + blockIdx.x] = v;
}

__global__ void __launch_bounds__(256) sx_loss_bwd_kernel(long long n, int use_edges, const float* __restrict__ grad_loss, float inv_count,
                                                          const float* __restrict__ ix, const float* __restrict__ iy,
                                                          float* __restrict__ g_ddx, float* __restrict__ g_ddy) {
    const long long q = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (q >= n) return;
    const float g = __ldg(grad_loss)*inv_count;
    g_ddx[q] = use_edges ? g*__expf(-ix[q]) : g;
    g_ddy[q] = use_edges ? g*__expf(-iy[q]) : g;
}

__global__ void sx_reduce_kernel(const float* __restrict__ partial, int n, double inv_count, float* __restrict__ out) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) a += (double)partial[q];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x/2; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) *out = (float)(sh[0]*inv_count);
}

}  // namespace stv

using namespace stv;

namespace {
struct SxPlan {   // workspace layout (floats unless stated); planes are b*HW floats, image planes b*C*HW
    size_t plane, iplane, total;
    size_t sum, dot;                 // doubles [b] each
    size_t dn, a1x, a1y, ddx, ddy;   // disparity chain (a1* only with the Laplacian)
    size_t ix, iy;                   // channel-mean image gradients (saved for the backward)
    size_t t0, t1;                   // image-sized temporaries (forward), plane-sized temporaries (backward)
    size_t partial;
};
SxPlan sx_plan(int b, int C, int H, int W) {
    SxPlan p{};
    const size_t al = 64;
    auto take = [&](size_t& cur, size_t n) { const size_t at = cur; cur += (n + al - 1)/al*al; return at; };
    p.plane = (size_t)b*H*W; p.iplane = p.plane*C;
    size_t cur = 0;
    p.sum = take(cur, 2*(size_t)b); p.dot = take(cur, 2*(size_t)b);
    p.dn = take(cur, p.plane); p.a1x = take(cur, p.plane); p.a1y = take(cur, p.plane); p.ddx = take(cur, p.plane); p.ddy = take(cur, p.plane);
    p.ix = take(cur, p.plane); p.iy = take(cur, p.plane);
    const size_t tmp = p.iplane > 2*p.plane ? p.iplane : 2*p.plane;   // the backward keeps two planes in each
    p.t0 = take(cur, tmp); p.t1 = take(cur, tmp);
    p.partial = take(cur, (size_t)b*((H*W + 255)/256));
    p.total = cur;
    return p;
}
int sx_check(int b, int C, int H, int W) {
    STV_REQUIRE(b > 0 && C > 0 && H >= 3 && W >= 3, "stv_smooth_ex: bad shape (b=%d C=%d H=%d W=%d; H, W >= 3)", b, C, H, W);
    STV_REQUIRE(b*C <= 65535, "stv_smooth_ex: too many planes for one launch");
    return STV_OK;
}
}  // namespace

extern "C" size_t stv_smooth_ex_workspace_bytes(int b, int C, int H, int W) {
    if (b <= 0 || C <= 0 || H < 3 || W < 3) return 0;
    return sx_plan(b, C, H, W).total*sizeof(float);
}

#define SX_LAUNCH(name, ...) do { name<<<__VA_ARGS__; count_launch(); if (int rc_ = check_launch(#name)) return rc_; } while (0)

extern "C" int stv_smooth_ex_fwd(int b, int C, int H, int W, int use_edges, int use_laplacian, int use_blur, const float* disp,
                                 const float* img, float* loss, float* disp_grad, float* image_grad, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = sx_check(b, C, H, W)) return rc;
    STV_REQUIRE(disp && img && loss, "stv_smooth_ex_fwd: NULL pointer");
    if (!ws || ws_bytes < stv_smooth_ex_workspace_bytes(b, C, H, W)) { set_error("stv_smooth_ex_fwd: workspace too small"); return STV_E_WORKSPACE; }
    const SxPlan p = sx_plan(b, C, H, W);
    float* w = (float*)ws;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H*W, nb = (HW + 255)/256;
    double* sum = (double*)(w + p.sum);
    SX_LAUNCH(sx_image_dot_kernel, b, 256, 0, st>>>(HW, disp, nullptr, sum));
    SX_LAUNCH(sx_normalize_kernel, dim3(nb, b), 256, 0, st>>>(HW, disp, sum, w + p.dn));
    const dim3 gd(nb, b), gi(nb, b*C);
    if (use_laplacian) {
        SX_LAUNCH(sx_absdiff_kernel, gd, 256, 0, st>>>(H, W, 0, use_blur, w + p.dn, w + p.a1x));
        SX_LAUNCH(sx_absdiff_kernel, gd, 256, 0, st>>>(H, W, 1, use_blur, w + p.dn, w + p.a1y));
        SX_LAUNCH(sx_absdiff_kernel, gd, 256, 0, st>>>(H, W, 0, use_blur, w + p.a1x, w + p.ddx));
        SX_LAUNCH(sx_absdiff_kernel, gd, 256, 0, st>>>(H, W, 1, use_blur, w + p.a1y, w + p.ddy));
    } else {
        SX_LAUNCH(sx_absdiff_kernel, gd, 256, 0, st>>>(H, W, 0, use_blur, w + p.dn, w + p.ddx));
        SX_LAUNCH(sx_absdiff_kernel, gd, 256, 0, st>>>(H, W, 1, use_blur, w + p.dn, w + p.ddy));
    }
    for (int axis = 0; axis < 2; ++axis) {   // image gradients: per channel, then the channel mean (compute_grad(ch_mean=True))
        float* dst = w + (axis == 0 ? p.ix : p.iy);
        SX_LAUNCH(sx_absdiff_kernel, gi, 256, 0, st>>>(H, W, axis, use_blur, img, w + p.t0));
        if (use_laplacian) {
            SX_LAUNCH(sx_absdiff_kernel, gi, 256, 0, st>>>(H, W, axis, use_blur, w + p.t0, w + p.t1));
            SX_LAUNCH(sx_chmean_kernel, gd, 256, 0, st>>>(HW, C, w + p.t1, dst));
        } else SX_LAUNCH(sx_chmean_kernel, gd, 256, 0, st>>>(HW, C, w + p.t0, dst));
    }
    SX_LAUNCH(sx_loss_kernel, gd, 256, 0, st>>>(HW, use_edges, w + p.ddx, w + p.ddy, w + p.ix, w + p.iy, w + p.partial, disp_grad, image_grad));
    SX_LAUNCH(sx_reduce_kernel, 1, 256, 0, st>>>(w + p.partial, nb*b, 1.0/((double)b*HW), loss));
    return STV_OK;
}

/* `ws` must be the workspace filled by the matching stv_smooth_ex_fwd call. */
extern "C" int stv_smooth_ex_bwd(int b, int C, int H, int W, int use_edges, int use_laplacian, int use_blur, const float* disp,
                                 const float* grad_loss, float* g_disp, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = sx_check(b, C, H, W)) return rc;
    STV_REQUIRE(disp && grad_loss && g_disp, "stv_smooth_ex_bwd: NULL pointer");
    if (!ws || ws_bytes < stv_smooth_ex_workspace_bytes(b, C, H, W)) { set_error("stv_smooth_ex_bwd: workspace too small"); return STV_E_WORKSPACE; }
    const SxPlan p = sx_plan(b, C, H, W);
    float* w = (float*)ws;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H*W, nb = (HW + 255)/256;
    const long long n = (long long)b*HW;
    const dim3 gd(nb, b);
    const unsigned nl = (unsigned)((n + 255)/256);
    // plane-sized scratch inside the image-sized temporaries: gx, gy (upstream of ddx, ddy), u (D^T output), ga (gradient of a1*), gdn
    float *gx = w + p.t0, *gy = gx + p.plane, *u = w + p.t1, *ga = u + p.plane, *gdn = w + p.ddx;   // ddx is dead once gx, gy exist
    SX_LAUNCH(sx_loss_bwd_kernel, nl, 256, 0, st>>>(n, use_edges, grad_loss, (float)(1.0/((double)b*HW)), w + p.ix, w + p.iy, gx, gy));
    // adjoint of A_axis at input `xin` for upstream `gin`, written / accumulated into `gout`
    auto adj = [&](int axis, const float* xin, const float* gin, float* gout, int accumulate) -> int {
        if (use_blur) {
            SX_LAUNCH(sx_absdiff_bwd_kernel, gd, 256, 0, st>>>(H, W, axis, 1, xin, gin, u));
            SX_LAUNCH(sx_blur_adj_kernel, gd, 256, 0, st>>>(H, W, u, gout, accumulate));
        } else if (accumulate) {
            SX_LAUNCH(sx_absdiff_bwd_kernel, gd, 256, 0, st>>>(H, W, axis, 0, xin, gin, u));
            SX_LAUNCH(sx_add_kernel, nl, 256, 0, st>>>(n, u, gout, 1));
        } else SX_LAUNCH(sx_absdiff_bwd_kernel, gd, 256, 0, st>>>(H, W, axis, 0, xin, gin, gout));
        return STV_OK;
    };
    if (use_laplacian) {
        if (int rc = adj(0, w + p.a1x, gx, ga, 0)) return rc;        // d/d a1x
        if (int rc = adj(0, w + p.dn, ga, gdn, 0)) return rc;        // -> d/d dn
        if (int rc = adj(1, w + p.a1y, gy, ga, 0)) return rc;        // d/d a1y
        if (int rc = adj(1, w + p.dn, ga, gdn, 1)) return rc;
    } else {
        if (int rc = adj(0, w + p.dn, gx, gdn, 0)) return rc;
        if (int rc = adj(1, w + p.dn, gy, gdn, 1)) return rc;
    }
    double* sum = (double*)(w + p.sum); double* dot = (double*)(w + p.dot);
    SX_LAUNCH(sx_image_dot_kernel, b, 256, 0, st>>>(HW, gdn, disp, dot));
    SX_LAUNCH(sx_normalize_bwd_kernel, gd, 256, 0, st>>>(HW, gdn, sum, dot, g_disp));
    return STV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Feature regularisers (src/regularizers/smooth.py:100-176; registry keys `feat_peaky`, `feat_smooth`): the same building block on
// C-channel feature maps, no mean normalisation, image-edge weights broadcast over the channels.
//   order 1 (FeatPeakReg)    loss = -( mean(A_x f * wx) + mean(A_y f * wy) ),                 w = exp(-mean_c A(img)) | 1
//   order 2 (FeatSmoothReg)  loss = mean(A_x A_x f * wxx) + mean(A_y A_y f * wyy) + mean(A_y A_x f * wxy) + mean(A_x A_y f * wyx),
//                            w = exp(-mean_c of the same second differences of the image) | 1
// feat (b, C, H, W), img (b, Ci, H, W) at the same resolution; feat_grad (b, C, H, W) = sqrt(clamp(t0^2 + t1^2, eps)) of the first two
// terms (the logging map). The backward needs the workspace the forward filled.
// ---------------------------------------------------------------------------------------------------------------------
namespace stv {

// partial[plane][block] = sum over the block's pixels of term * (w ? exp(-w[image of the plane]) : 1)
__global__ void __launch_bounds__(256) fr_wsum_kernel(int HW, int C, const float* __restrict__ term, const float* __restrict__ w,
                                                      float* __restrict__ partial) {
    __shared__ float red[32];
    const int q = blockIdx.x*blockDim.x + threadIdx.x, pl = blockIdx.y;
    float v = 0.f;
    if (q < HW) v = term[(size_t)pl*HW + q]*(w ? __expf(-w[(size_t)(pl/C)*HW + q]) : 1.f);
    v = block_sum(v, red);
    if (threadIdx.x == 0) partial[(size_t)pl*gridDim.x + blockIdx.x] = v;
}

// upstream gradient of a term: g[plane][q] = dL/dloss * coef * (w ? exp(-w) : 1)
__global__ void __launch_bounds__(256) fr_upstream_kernel(int HW, int C, const float* __restrict__ grad_loss, float coef,
                                                          const float* __restrict__ w, float* __restrict__ g) {
    const int q = blockIdx.x*blockDim.x + threadIdx.x, pl = blockIdx.y;
    if (q >= HW) return;
    g[(size_t)pl*HW + q] = __ldg(grad_loss)*coef*(w ? __expf(-w[(size_t)(pl/C)*HW + q]) : 1.f);
}

__global__ void __launch_bounds__(256) fr_gradmap_kernel(long long n, const float* __restrict__ t0, const float* __restrict__ t1,
                                                         float* __restrict__ out) {
    const long long q = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (q < n) out[q] = sqrtf(fmaxf(t0[q]*t0[q] + t1[q]*t1[q], STV_EPS32));
}

}  // namespace stv

namespace {
struct FrPlan {   // floats; P = b*C feature planes, I = b*Ci image planes, B = b planes
    size_t P, I, B, total, npart;
    size_t a1x, a1y;          // first differences of the features (order 2 keeps them for the backward)
    size_t t[4];              // the terms: order 1: A_x f, A_y f; order 2: xx, yy, xy, yx
    size_t w[4];              // channel-mean image terms (b planes each)
    size_t i0, i1;            // image-sized temporaries
    size_t g0, g1, g2;        // feature-sized temporaries of the backward
    size_t partial;
};
FrPlan fr_plan(int b, int C, int Ci, int H, int W) {
    FrPlan p{};
    const size_t al = 64, HW = (size_t)H*W;
    auto take = [&](size_t& cur, size_t n) { const size_t at = cur; cur += (n + al - 1)/al*al; return at; };
    p.P = (size_t)b*C*HW; p.I = (size_t)b*Ci*HW; p.B = (size_t)b*HW;
    p.npart = (size_t)b*C*((HW + 255)/256);
    size_t cur = 0;
    p.a1x = take(cur, p.P); p.a1y = take(cur, p.P);
    for (int k = 0; k < 4; ++k) p.t[k] = take(cur, p.P);
    for (int k = 0; k < 4; ++k) p.w[k] = take(cur, p.B);
    p.i0 = take(cur, p.I); p.i1 = take(cur, p.I);
    p.g0 = take(cur, p.P); p.g1 = take(cur, p.P); p.g2 = take(cur, p.P);
    p.partial = take(cur, 4*p.npart);
    p.total = cur;
    return p;
}
int fr_check(int b, int C, int Ci, int H, int W, int order, const char* who) {
    STV_REQUIRE(b > 0 && C > 0 && Ci > 0 && H >= 2 && W >= 2, "%s: bad shape (b=%d C=%d Ci=%d H=%d W=%d)", who, b, C, Ci, H, W);
    STV_REQUIRE((long long)b*C <= 65535 && (long long)b*Ci <= 65535, "%s: too many planes for one launch", who);
    STV_REQUIRE(order == 1 || order == 2, "%s: order must be 1 (feat_peaky) or 2 (feat_smooth)", who);
    return STV_OK;
}
// the four second-order terms xx, yy, xy (= A_y A_x), yx (= A_x A_y): axis of the first and of the second difference
const int FR_AX1[4] = {0, 1, 0, 1}, FR_AX2[4] = {0, 1, 1, 0};
}  // namespace

extern "C" size_t stv_feat_reg_workspace_bytes(int b, int C, int Ci, int H, int W) {
    if (b <= 0 || C <= 0 || Ci <= 0 || H < 2 || W < 2) return 0;
    return fr_plan(b, C, Ci, H, W).total*sizeof(float);
}

extern "C" int stv_feat_reg_fwd(int b, int C, int Ci, int H, int W, int order, int use_edges, const float* feat, const float* img,
                                float* loss, float* feat_grad, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = fr_check(b, C, Ci, H, W, order, "stv_feat_reg_fwd")) return rc;
    STV_REQUIRE(feat && img && loss, "stv_feat_reg_fwd: NULL pointer");
    if (!ws || ws_bytes < stv_feat_reg_workspace_bytes(b, C, Ci, H, W)) { set_error("stv_feat_reg_fwd: workspace too small"); return STV_E_WORKSPACE; }
    const FrPlan p = fr_plan(b, C, Ci, H, W);
    float* w = (float*)ws;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H*W, nb = (HW + 255)/256, nterm = order == 1 ? 2 : 4;
    const dim3 gf(nb, b*C), gi(nb, b*Ci), gb(nb, b);
    if (order == 1) {
        SX_LAUNCH(sx_absdiff_kernel, gf, 256, 0, st>>>(H, W, 0, 0, feat, w + p.t[0]));
        SX_LAUNCH(sx_absdiff_kernel, gf, 256, 0, st>>>(H, W, 1, 0, feat, w + p.t[1]));
    } else {
        SX_LAUNCH(sx_absdiff_kernel, gf, 256, 0, st>>>(H, W, 0, 0, feat, w + p.a1x));
        SX_LAUNCH(sx_absdiff_kernel, gf, 256, 0, st>>>(H, W, 1, 0, feat, w + p.a1y));
        for (int k = 0; k < 4; ++k)
            SX_LAUNCH(sx_absdiff_kernel, gf, 256, 0, st>>>(H, W, FR_AX2[k], 0, w + (FR_AX1[k] == 0 ? p.a1x : p.a1y), w + p.t[k]));
    }
    if (use_edges) {
        for (int k = 0; k < nterm; ++k) {
            SX_LAUNCH(sx_absdiff_kernel, gi, 256, 0, st>>>(H, W, order == 1 ? k : FR_AX1[k], 0, img, w + p.i0));
            if (order == 2) {
                SX_LAUNCH(sx_absdiff_kernel, gi, 256, 0, st>>>(H, W, FR_AX2[k], 0, w + p.i0, w + p.i1));
                SX_LAUNCH(sx_chmean_kernel, gb, 256, 0, st>>>(HW, Ci, w + p.i1, w + p.w[k]));
            } else SX_LAUNCH(sx_chmean_kernel, gb, 256, 0, st>>>(HW, Ci, w + p.i0, w + p.w[k]));
        }
    }
    for (int k = 0; k < nterm; ++k)
        SX_LAUNCH(fr_wsum_kernel, gf, 256, 0, st>>>(HW, C, w + p.t[k], use_edges ? w + p.w[k] : nullptr, w + p.partial + (size_t)k*p.npart));
    const double sign = order == 1 ? -1.0 : 1.0;   // peakiness is maximised
    SX_LAUNCH(sx_reduce_kernel, 1, 256, 0, st>>>(w + p.partial, (int)(nterm*p.npart), sign/((double)b*C*HW), loss));
    if (feat_grad) {
        const long long n = (long long)b*C*HW;
        SX_LAUNCH(fr_gradmap_kernel, (unsigned)((n + 255)/256), 256, 0, st>>>(n, w + p.t[0], w + p.t[1], feat_grad));
    }
    return STV_OK;
}

extern "C" int stv_feat_reg_bwd(int b, int C, int Ci, int H, int W, int order, int use_edges, const float* feat, const float* grad_loss,
                                float* g_feat, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = fr_check(b, C, Ci, H, W, order, "stv_feat_reg_bwd")) return rc;
    STV_REQUIRE(feat && grad_loss && g_feat, "stv_feat_reg_bwd: NULL pointer");
    if (!ws || ws_bytes < stv_feat_reg_workspace_bytes(b, C, Ci, H, W)) { set_error("stv_feat_reg_bwd: workspace too small"); return STV_E_WORKSPACE; }
    const FrPlan p = fr_plan(b, C, Ci, H, W);
    float* w = (float*)ws;
    cudaStream_t st = (cudaStream_t)stream;
    const int HW = H*W, nb = (HW + 255)/256;
    const long long n = (long long)b*C*HW;
    const dim3 gf(nb, b*C);
    const unsigned nl = (unsigned)((n + 255)/256);
    const float coef = (order == 1 ? -1.f : 1.f)/((float)b*(float)C*(float)HW);
    float *g0 = w + p.g0, *g1 = w + p.g1, *g2 = w + p.g2;
    auto upstream = [&](int k) -> int { SX_LAUNCH(fr_upstream_kernel, gf, 256, 0, st>>>(HW, C, grad_loss, coef, use_edges ? w + p.w[k] : nullptr, g0)); return STV_OK; };
    // out (+)= A_axis^T (xin; gin)
    auto adj = [&](int axis, const float* xin, const float* gin, float* out, int accumulate) -> int {
        if (!accumulate) { SX_LAUNCH(sx_absdiff_bwd_kernel, gf, 256, 0, st>>>(H, W, axis, 0, xin, gin, out)); return STV_OK; }
        SX_LAUNCH(sx_absdiff_bwd_kernel, gf, 256, 0, st>>>(H, W, axis, 0, xin, gin, g2));
        SX_LAUNCH(sx_add_kernel, nl, 256, 0, st>>>(n, g2, out, 1));
        return STV_OK;
    };
    if (order == 1) {
        if (int rc = upstream(0)) return rc;
        if (int rc = adj(0, feat, g0, g_feat, 0)) return rc;
        if (int rc = upstream(1)) return rc;
        return adj(1, feat, g0, g_feat, 1);
    }
    // order 2: gradients of the first differences (g1 = d/d a1x, then d/d a1y), each pulled back to the features
    for (int first = 0; first < 2; ++first) {
        const float* a1 = w + (first == 0 ? p.a1x : p.a1y);
        int seen = 0;
        for (int k = 0; k < 4; ++k) {
            if (FR_AX1[k] != first) continue;
            if (int rc = upstream(k)) return rc;
            if (int rc = adj(FR_AX2[k], a1, g0, g1, seen++)) return rc;
        }
        if (int rc = adj(first, feat, g1, g_feat, first)) return rc;
    }
    return STV_OK;
}
