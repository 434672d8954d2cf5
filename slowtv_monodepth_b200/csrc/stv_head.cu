// Disparity heads of the Monodepth decoder: reflection-padded 3x3 convolution to ONE output channel followed by the output
// activation (sigmoid) — `outconv_i`, src/networks/decoders/monodepth.py:66-69, applied at :86-87 through conv3x3 of utils.py:44-46.
// With a single output channel this is a 9*C-long dot product per pixel: 2 flop per byte read, HBM/L1-bound (SURVEY App. B,
// "N=1 => a dot-product reduction, not a GEMM"), so it runs on the CUDA cores and reads the feature map once per tap through
// L1 instead of wasting a 128 x 32 tensor-core tile on one column. Channels-last fp32, C a multiple of 4, C <= 128.
//
// Thread mapping (all three kernels): LP = C/4 adjacent lanes share one pixel, each owning 4 channels (one 128-bit load per tap);
// a warp therefore reads 32/LP neighbouring pixels x 16*LP contiguous bytes per tap. The nine filter taps of a lane live in registers.
#include "stv_common.cuh"
#include "stv_epi.cuh"

namespace stv {

constexpr int HD_THREADS = 256;

__device__ __forceinline__ float group_sum(float v, int lp) {  // sum over the lp (power of two) adjacent lanes of a pixel
    for (int o = lp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void load_taps(const float* __restrict__ w, int C, int c, float4 (&wt)[9]) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = __ldg((const float4*)(w + t*C + c));
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w*b.w))); }

// y[n,p,q] = act(bias + sum_{r,s,c} x[n, refl(p+r-1), refl(q+s-1), c] * w[r,s,c])
__global__ void __launch_bounds__(HD_THREADS) head3x3_fwd_kernel(int N, int H, int W, int C, const float* __restrict__ x,
                                                                 const float* __restrict__ w, const float* __restrict__ bias, int act,
                                                                 float* __restrict__ y) {
    const int lp = C >> 2, ppb = HD_THREADS/lp;  // lanes per pixel, pixels per block pass
    const int g = threadIdx.x/lp, c = (threadIdx.x % lp)*4;
    float4 wt[9];
    load_taps(w, C, c, wt);
    const float b = bias ? __ldg(bias) : 0.f;
    const long long npix = (long long)N*H*W;
    const long long iters = (npix + (long long)gridDim.x*ppb - 1)/((long long)gridDim.x*ppb);  // same trip count for every lane: warp shuffles below
    for (long long it = 0; it < iters; ++it) {
        const long long pix_ = (it*gridDim.x + blockIdx.x)*ppb + g;
        const bool valid = pix_ < npix;
        const long long pix = valid ? pix_ : 0;
        const int ipix = (int)pix, n = ipix/(H*W), rem = ipix - n*H*W, p = rem/W, q = rem - p*W;  // 32-bit divisions (npix < 2^31)
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int yy = reflect_idx(p + r - 1, H);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int xx = reflect_idx(q + s - 1, W);
                acc += dot4(__ldg((const float4*)(x + ((size_t)(n*H + yy)*W + xx)*C + c)), wt[r*3 + s]);
            }
        }
        acc = group_sum(acc, lp);
        if (c == 0 && valid) y[pix] = act_fwd(act, acc + b);
    }
}

// dz[n,p,q] = dA*act'(y). dx[n,yy,xx,c] = sum over output pixels (p,q) and taps (r,s) with refl(p+r-1) = yy, refl(q+s-1) = xx of
// dz[n,p,q]*w[r,s,c]. Enumerated as: padded positions (yp,xp) that reflection maps onto (yy,xx) (1..2 per axis), taps, p = yp-r, q = xp-s.
__global__ void __launch_bounds__(HD_THREADS) head3x3_dgrad_kernel(int N, int H, int W, int C, const float* __restrict__ dA,
                                                                   const float* __restrict__ y, int act, const float* __restrict__ w,
                                                                   float* __restrict__ dx) {
    const int lp = C >> 2, ppb = HD_THREADS/lp;
    const int g = threadIdx.x/lp, c = (threadIdx.x % lp)*4;
    float4 wt[9];
    load_taps(w, C, c, wt);
    const long long npix = (long long)N*H*W;
    for (long long pix = (long long)blockIdx.x*ppb + g; pix < npix; pix += (long long)gridDim.x*ppb) {
        const int ipix = (int)pix, n = ipix/(H*W), rem = ipix - n*H*W, yy = rem/W, xx = rem - yy*W;
        // padded coordinates (pad 1) that read input row yy: yy+1, plus 0 when yy == 1, plus H+1 when yy == H-2
        int ys[3], xs[3], ny = 0, nx = 0;
        ys[ny++] = yy + 1; if (yy == 1) ys[ny++] = 0; if (yy == H - 2) ys[ny++] = H + 1;
        xs[nx++] = xx + 1; if (xx == 1) xs[nx++] = 0; if (xx == W - 2) xs[nx++] = W + 1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int a = 0; a < ny; ++a)
            for (int bq = 0; bq < nx; ++bq) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int p = ys[a] - r;
                    if (p < 0 || p >= H) continue;
#pragma unroll
                    for (int s = 0; s < 3; ++s) {
                        const int q = xs[bq] - s;
                        if (q < 0 || q >= W) continue;
                        const size_t o = (size_t)(n*H + p)*W + q;
                        const float dz = __ldg(dA + o)*act_bwd(act, __ldg(y + o));
                        const float4 t = wt[r*3 + s];
                        acc.x = fmaf(dz, t.x, acc.x); acc.y = fmaf(dz, t.y, acc.y); acc.z = fmaf(dz, t.z, acc.z); acc.w = fmaf(dz, t.w, acc.w);
                    }
                }
            }
        *(float4*)(dx + (size_t)pix*C + c) = acc;
    }
}

// dw[r,s,c] += sum_{n,p,q} dz[n,p,q] * x[n, refl(p+r-1), refl(q+s-1), c];  db += sum dz.
__global__ void __launch_bounds__(HD_THREADS) head3x3_wgrad_kernel(int N, int H, int W, int C, const float* __restrict__ x,
                                                                   const float* __restrict__ dA, const float* __restrict__ y, int act,
                                                                   float* __restrict__ dw, float* __restrict__ db) {
    __shared__ __align__(16) float red[HD_THREADS/32][9*128 + 4];  // per-warp partial (9 taps x up to 128 channels) + bias term
    const int lp = C >> 2, ppb = HD_THREADS/lp;
    const int g = threadIdx.x/lp, c = (threadIdx.x % lp)*4;
    float4 acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    float bsum = 0.f;
    const long long npix = (long long)N*H*W;
    for (long long pix = (long long)blockIdx.x*ppb + g; pix < npix; pix += (long long)gridDim.x*ppb) {  // no shuffles inside: ragged trip counts are fine
        const int ipix = (int)pix, n = ipix/(H*W), rem = ipix - n*H*W, p = rem/W, q = rem - p*W;
        const float dz = __ldg(dA + pix)*act_bwd(act, __ldg(y + pix));
        bsum += dz;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int yy = reflect_idx(p + r - 1, H);
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int xx = reflect_idx(q + s - 1, W);
                const float4 v = __ldg((const float4*)(x + ((size_t)(n*H + yy)*W + xx)*C + c));
                float4& a = acc[r*3 + s];
                a.x = fmaf(dz, v.x, a.x); a.y = fmaf(dz, v.y, a.y); a.z = fmaf(dz, v.z, a.z); a.w = fmaf(dz, v.w, a.w);
            }
        }
    }
    // combine the pixel groups of a warp (lanes with the same channel offset are lp apart), then the warps, then one atomic per value
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = lp; o < 32; o <<= 1) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, o); acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, o);
            acc[t].z += __shfl_xor_sync(0xffffffffu, acc[t].z, o); acc[t].w += __shfl_xor_sync(0xffffffffu, acc[t].w, o);
        }
    }
    if (c != 0) bsum = 0.f;  // every lane of a pixel group saw the same dz: count it once
    bsum = warp_sum(bsum);
    if (lane < lp) {
#pragma unroll
        for (int t = 0; t < 9; ++t) *(float4*)&red[wid][t*C + c] = acc[t];
    }
    if (lane == 0) red[wid][9*128] = bsum;
    __syncthreads();
    for (int i = threadIdx.x; i < 9*C; i += HD_THREADS) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < HD_THREADS/32; ++k) v += red[k][i];
        atomicAdd(dw + i, v);
    }
    if (threadIdx.x == 0 && db) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < HD_THREADS/32; ++k) v += red[k][9*128];
        atomicAdd(db, v);
    }
}

static int head_check(int N, int H, int W, int C, const char* what) {
    STV_REQUIRE(N > 0 && H >= 2 && W >= 2 && (long long)N*H*W < (1ll << 31), "%s: bad image size", what);
    STV_REQUIRE(C >= 4 && C <= 128 && (C & (C - 1)) == 0, "%s: channels must be a power of two in [4, 128] (got %d)", what, C);
    return STV_OK;
}

static int head_blocks(long long npix, int C) {
    const long long ppb = HD_THREADS/(C/4), want = (npix + ppb - 1)/ppb;
    return (int)(want < 148ll*16 ? want : 148ll*16);
}

}  // namespace stv

using namespace stv;

extern "C" int stv_head3x3_fwd(int N, int H, int W, int C, const float* x, const float* w, const float* bias, int act, float* y, void* stream) {
    if (int rc = head_check(N, H, W, C, "stv_head3x3_fwd")) return rc;
    STV_REQUIRE(x && w && y, "stv_head3x3_fwd: null pointer");
    head3x3_fwd_kernel<<<head_blocks((long long)N*H*W, C), HD_THREADS, 0, (cudaStream_t)stream>>>(N, H, W, C, x, w, bias, act, y);
    count_launch();
    return check_launch("stv_head3x3_fwd");
}

extern "C" int stv_head3x3_bwd(int N, int H, int W, int C, const float* x, const float* w, const float* da, const float* y, int act, float* dx,
                               float* dw, float* db, void* stream) {
    if (int rc = head_check(N, H, W, C, "stv_head3x3_bwd")) return rc;
    STV_REQUIRE(x && w && da && y, "stv_head3x3_bwd: null pointer");
    const int blocks = head_blocks((long long)N*H*W, C);
    if (dx) { head3x3_dgrad_kernel<<<blocks, HD_THREADS, 0, (cudaStream_t)stream>>>(N, H, W, C, da, y, act, w, dx); count_launch(); }
    if (dw) {
        // fewer, longer-running blocks: every block ends with 9*C atomics
        const int wb = blocks < 148*4 ? blocks : 148*4;
        head3x3_wgrad_kernel<<<wb, HD_THREADS, 0, (cudaStream_t)stream>>>(N, H, W, C, x, da, y, act, dw, db);
        count_launch();
    }
    return check_launch("stv_head3x3_bwd");
}
