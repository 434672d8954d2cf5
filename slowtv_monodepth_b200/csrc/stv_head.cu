// Disparity heads of the Monodepth decoder: reflection-padded 3x3 convolution to ONE output channel followed by the output
// activation (sigmoid) — `outconv_i`, src/networks/decoders/monodepth.py:66-69, applied at :86-87 through conv3x3 of utils.py:44-46.
// With a single output channel this is a 9*C-long dot product per pixel: 2 flop per byte read, HBM/L1-bound (SURVEY App. B,
// "N=1 => a dot-product reduction, not a GEMM"), so it runs on the CUDA cores and reads the feature map once per tap through
// L1 instead of wasting a 128 x 32 tensor-core tile on one column. Channels-last fp32, C a multiple of 4, C <= 128.
//
// Thread mapping (all kernels): LP = C/4 adjacent lanes share a RUN of HD_RUN = 4 horizontally adjacent pixels, each lane owning
// 4 channels (one 128-bit load per tap column). A block owns one image row (blockIdx.x = n*H + y: no per-pixel divisions) and
// 256/LP runs of it (blockIdx.y = x tile). Along a run the 3x3 windows overlap, so a lane loads 3 x 6 columns for 4 pixels
// (4.5 loads per pixel instead of 9) and the reflection arithmetic is done once per row / per column of the run.
// v1 of these kernels (one pixel per lane group, grid-stride over pixels, activation derivative recomputed per tap) ran at
// ~0.9 TB/s on the 16-channel full-resolution head: ~190 instructions per lane and pixel, most of them index arithmetic.
#include "stv_common.cuh"
#include "stv_epi.cuh"

namespace stv {

constexpr int HD_THREADS = 256;
constexpr int HD_RUN = 4;

__device__ __forceinline__ float group_sum(float v, int lp) {  // sum over the lp (power of two) adjacent lanes of a run
    for (int o = lp >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void load_taps(const float* __restrict__ w, int C, int c, float4 (&wt)[9]) {
#pragma unroll
    for (int t = 0; t < 9; ++t) wt[t] = __ldg((const float4*)(w + t*C + c));
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w*b.w))); }
__device__ __forceinline__ void axpy4(float4& acc, float a, const float4& v) {
    acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y); acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
}
__device__ __forceinline__ int clamp_refl(int i, int n) { return min(max(reflect_idx(i, n), 0), n - 1); }

// Column offsets (elements, incl. the lane's channel offset) of the HD_RUN + 2 reflect-padded window columns of a run.
__device__ __forceinline__ void window_cols(int x0, int W, int C, int c, int (&xo)[HD_RUN + 2]) {
#pragma unroll
    for (int j = 0; j < HD_RUN + 2; ++j) xo[j] = clamp_refl(x0 + j - 1, W)*C + c;   // columns past the image are clamped: never stored
}
// One window row (padded row index y + r - 1) of channels c..c+3.
__device__ __forceinline__ void load_window_row(const float* __restrict__ x, int n, int yr, int H, int W, int C, const int (&xo)[HD_RUN + 2],
                                                float4 (&v)[HD_RUN + 2]) {
    const float* row = x + (size_t)(n*H + clamp_refl(yr, H))*W*C;
#pragma unroll
    for (int j = 0; j < HD_RUN + 2; ++j) v[j] = __ldg((const float4*)(row + xo[j]));
}

// y[n,p,q] = act(bias + sum_{r,s,c} x[n, refl(p+r-1), refl(q+s-1), c] * w[r,s,c]).   grid = (N*H, x tiles)
__global__ void __launch_bounds__(HD_THREADS) head3x3_fwd_kernel(int H, int W, int C, const float* __restrict__ x,
                                                                 const float* __restrict__ w, const float* __restrict__ bias, int act,
                                                                 float* __restrict__ y) {
    const int lp = C >> 2, rpb = HD_THREADS/lp;
    const int run = threadIdx.x/lp, c = (threadIdx.x % lp)*4;
    const int row = blockIdx.x, n = row/H, p = row - n*H;
    const int x0 = (blockIdx.y*rpb + run)*HD_RUN;
    float4 wt[9];
    load_taps(w, C, c, wt);
    int xo[HD_RUN + 2];
    window_cols(min(x0, W - 1), W, C, c, xo);   // runs past the row still take part in the shuffles below
    float acc[HD_RUN];
#pragma unroll
    for (int u = 0; u < HD_RUN; ++u) acc[u] = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float4 v[HD_RUN + 2];
        load_window_row(x, n, p + r - 1, H, W, C, xo, v);
#pragma unroll
        for (int u = 0; u < HD_RUN; ++u)
#pragma unroll
            for (int s2 = 0; s2 < 3; ++s2) acc[u] += dot4(v[u + s2], wt[r*3 + s2]);
    }
#pragma unroll
    for (int u = 0; u < HD_RUN; ++u) acc[u] = group_sum(acc[u], lp);
    if (c == 0 && x0 < W) {
        const float b = bias ? __ldg(bias) : 0.f;
#pragma unroll
        for (int u = 0; u < HD_RUN; ++u)
            if (x0 + u < W) y[(size_t)row*W + x0 + u] = act_fwd(act, acc[u] + b);
    }
}

// dz = dA * act'(y): one pass over the one-channel maps, so that the gradient kernels read it instead of re-deriving it per tap.
__global__ void __launch_bounds__(256) head_dz_kernel(long long n, int act, const float* __restrict__ dA, const float* __restrict__ y,
                                                      float* __restrict__ dz) {
    const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) dz[i] = __ldg(dA + i)*act_bwd(act, __ldg(y + i));
}

// dx[n,yy,xx,c] = sum over output pixels (p,q) and taps (r,s) with refl(p+r-1) = yy, refl(q+s-1) = xx of dz[n,p,q]*w[r,s,c].
// Per axis the (output index, tap) pairs feeding input index i are (i+1,0), (i,1), (i-1,2) where the output exists, plus (0,0)
// when i == 1 and (last,2) when i == last-1 (the two reflected border rows / columns). Rows are uniform per block; the main
// column terms come from a 6-wide dz window of the run, the two reflected columns are a rare per-pixel extra.
__global__ void __launch_bounds__(HD_THREADS) head3x3_dgrad_kernel(int H, int W, int C, const float* __restrict__ dz,
                                                                   const float* __restrict__ w, float* __restrict__ dx) {
    const int lp = C >> 2, rpb = HD_THREADS/lp;
    const int run = threadIdx.x/lp, c = (threadIdx.x % lp)*4;
    const int row = blockIdx.x, n = row/H, yy = row - n*H;
    const int x0 = (blockIdx.y*rpb + run)*HD_RUN;
    if (x0 >= W) return;   // no shuffles in this kernel
    float4 wt[9];
    load_taps(w, C, c, wt);
    float4 acc[HD_RUN];
#pragma unroll
    for (int u = 0; u < HD_RUN; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* dzn = dz + (size_t)n*H*W;
    // One (output row p, vertical tap R) pair: R is a compile-time constant so the taps stay in registers.
    auto row_term = [&](int p, const float4& w0, const float4& w1, const float4& w2) {
        const float* dr = dzn + (size_t)p*W;
        float d[HD_RUN + 2];   // dz[p, x0-1 .. x0+4], zero outside the row
#pragma unroll
        for (int j = 0; j < HD_RUN + 2; ++j) {
            const int q = x0 + j - 1;
            d[j] = (q >= 0 && q < W) ? __ldg(dr + q) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < HD_RUN; ++u) {   // input column xx = x0+u: outputs q = xx+1 (s=0), xx (s=1), xx-1 (s=2)
            axpy4(acc[u], d[u + 2], w0); axpy4(acc[u], d[u + 1], w1); axpy4(acc[u], d[u], w2);
            const int xx = x0 + u;
            if (xx == 1) axpy4(acc[u], __ldg(dr), w0);                  // padded column 0 reflects onto column 1
            if (xx == W - 2) axpy4(acc[u], __ldg(dr + W - 1), w2);      // padded column W+1 reflects onto column W-2
        }
    };
    // (output row, tap) pairs of input row yy — block-uniform branches
    if (yy + 1 < H) row_term(yy + 1, wt[0], wt[1], wt[2]);
    row_term(yy, wt[3], wt[4], wt[5]);
    if (yy >= 1) row_term(yy - 1, wt[6], wt[7], wt[8]);
    if (yy == 1) row_term(0, wt[0], wt[1], wt[2]);               // padded row 0 reflects onto row 1
    if (yy == H - 2) row_term(H - 1, wt[6], wt[7], wt[8]);       // padded row H+1 reflects onto row H-2
    float* o = dx + ((size_t)row*W + x0)*C + c;
#pragma unroll
    for (int u = 0; u < HD_RUN; ++u)
        if (x0 + u < W) *(float4*)(o + (size_t)u*C) = acc[u];
}

// dw[r,s,c] += sum_{n,p,q} dz[n,p,q] * x[n, refl(p+r-1), refl(q+s-1), c];  db += sum dz.   grid = (row groups, x tiles):
// a block walks rows blockIdx.x, blockIdx.x + gridDim.x, ... so that it ends with ONE set of 9*C atomics.
__global__ void __launch_bounds__(HD_THREADS) head3x3_wgrad_kernel(int NH, int H, int W, int C, const float* __restrict__ x,
                                                                   const float* __restrict__ dz, float* __restrict__ dw,
                                                                   float* __restrict__ db, int ytile0) {
    __shared__ __align__(16) float red[HD_THREADS/32][9*128 + 4];  // per-warp partial (9 taps x up to 128 channels) + bias term
    const int lp = C >> 2, rpb = HD_THREADS/lp;
    const int run = threadIdx.x/lp, c = (threadIdx.x % lp)*4;
    const int x0 = ((ytile0 + blockIdx.y)*rpb + run)*HD_RUN;
    float4 acc[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    float bsum = 0.f;
    if (x0 < W) {
        int xo[HD_RUN + 2];
        window_cols(x0, W, C, c, xo);
        for (int row = blockIdx.x; row < NH; row += gridDim.x) {
            const int n = row/H, p = row - n*H;
            float g[HD_RUN];
#pragma unroll
            for (int u = 0; u < HD_RUN; ++u) g[u] = x0 + u < W ? __ldg(dz + (size_t)row*W + x0 + u) : 0.f;
#pragma unroll
            for (int u = 0; u < HD_RUN; ++u) bsum += g[u];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float4 v[HD_RUN + 2];
                load_window_row(x, n, p + r - 1, H, W, C, xo, v);
#pragma unroll
                for (int u = 0; u < HD_RUN; ++u)
#pragma unroll
                    for (int s2 = 0; s2 < 3; ++s2) axpy4(acc[r*3 + s2], g[u], v[u + s2]);
            }
        }
    }
    // combine the runs of a warp (lanes with the same channel offset are lp apart), then the warps, then one atomic per value
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = lp; o < 32; o <<= 1) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, o); acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, o);
            acc[t].z += __shfl_xor_sync(0xffffffffu, acc[t].z, o); acc[t].w += __shfl_xor_sync(0xffffffffu, acc[t].w, o);
        }
    }
    if (c != 0) bsum = 0.f;  // every lane of a run saw the same dz: count it once
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

static int head_xtiles(int W, int C) {
    const int per_block = (HD_THREADS/(C/4))*HD_RUN;   // pixels of one row covered by a block
    return (W + per_block - 1)/per_block;
}

}  // namespace stv

using namespace stv;

extern "C" int stv_head3x3_fwd(int N, int H, int W, int C, const float* x, const float* w, const float* bias, int act, float* y, void* stream) {
    if (int rc = head_check(N, H, W, C, "stv_head3x3_fwd")) return rc;
    STV_REQUIRE(x && w && y, "stv_head3x3_fwd: null pointer");
    const int xt = head_xtiles(W, C);
    STV_REQUIRE(xt <= 65535, "stv_head3x3_fwd: row too wide");
    head3x3_fwd_kernel<<<dim3(N*H, xt), HD_THREADS, 0, (cudaStream_t)stream>>>(H, W, C, x, w, bias, act, y);
    count_launch();
    return check_launch("stv_head3x3_fwd");
}

extern "C" int stv_head3x3_bwd(int N, int H, int W, int C, const float* x, const float* w, const float* da, const float* y, int act, float* dx,
                               float* dw, float* db, float* dz_ws, void* stream) {
    if (int rc = head_check(N, H, W, C, "stv_head3x3_bwd")) return rc;
    STV_REQUIRE(x && w && da && y && dz_ws, "stv_head3x3_bwd: null pointer (dz_ws: N*H*W floats of scratch)");
    const int xt = head_xtiles(W, C);
    STV_REQUIRE(xt <= 65535, "stv_head3x3_bwd: row too wide");
    cudaStream_t st = (cudaStream_t)stream;
    const long long npix = (long long)N*H*W;
    head_dz_kernel<<<(unsigned)((npix + 255)/256), 256, 0, st>>>(npix, act, da, y, dz_ws);
    count_launch();
    if (dx) { head3x3_dgrad_kernel<<<dim3(N*H, xt), HD_THREADS, 0, st>>>(H, W, C, dz_ws, w, dx); count_launch(); }
    if (dw) {
        // few, long-running blocks: every block ends with 9*C atomics
        int rows = (148*4 + xt - 1)/xt;
        rows = rows < N*H ? rows : N*H;
        if (stv_deterministic()) {   // one block per launch, x-tile after x-tile: one contributor per weight at a time
            for (int t = 0; t < xt; ++t) { head3x3_wgrad_kernel<<<dim3(1, 1), HD_THREADS, 0, st>>>(N*H, H, W, C, x, dz_ws, dw, db, t); count_launch(); }
        } else {
            head3x3_wgrad_kernel<<<dim3(rows, xt), HD_THREADS, 0, st>>>(N*H, H, W, C, x, dz_ws, dw, db, 0);
            count_launch();
        }
    }
    return check_launch("stv_head3x3_bwd");
}
