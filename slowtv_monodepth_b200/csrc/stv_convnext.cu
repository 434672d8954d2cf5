// ConvNeXt block, memory-bound half: depthwise 7x7 convolution (forward / data-gradient / weight-gradient) and LayerNorm
// over channels (forward / backward), all on channels-last (N, H, W, C) fp32 tensors.
//
// Why hand-written: the library path for the depthwise weight gradient in channels-last launches one scalar wgrad engine
// plus two layout transposes PER CHANNEL GROUP (289 + 581 launches, ~30 ms of a 59 ms training step at 8x384x640 on B200;
// profiles/round1_step_breakdown.md). These kernels are HBM/L2-bound by construction: every thread owns one channel, a warp
// reads 128 contiguous bytes of a pixel, and the 7x7 window slides through registers.
//
// Reference call sites: the timm ConvNeXt encoder built at src/networks/depth.py:97 (third-party, see encoders.py).
#include <cstdlib>
#include "stv_common.cuh"

namespace stv {

constexpr int DW_L = 8;    // output pixels along x per thread (forward / dgrad)
constexpr int DW_RY = 4;   // output rows per block (weights stay in registers across them)

// y[n,yy,xx,c] = (bias[c]) + sum_{ky,kx} w[c][ky][kx] * x[n, yy+ky-3, xx+kx-3, c]   (zero padding)  [+ res[n,yy,xx,c]]
// FLIP = true uses w[c][6-ky][6-kx]: the data gradient of the same convolution.
constexpr int DW_CB = 128;  // channels (= threads) per block; blockIdx.x = x_tile * n_channel_blocks + channel_block

// Vertical sliding window: a block owns a strip of DW_L columns x `rows` output rows of one image; every thread owns one
// channel. Input rows are streamed top to bottom: each row is loaded ONCE (DW_L + 6 values per thread) and scattered into the
// 7 output rows it contributes to, held in a 7-slot ring of register accumulators; when the last contributing input row of
// an output row has been consumed the row is written out. Global loads per output drop from 12.25 (7 input rows re-read per
// output row) to (DW_L+6)/DW_L * (rows+6)/rows ~ 2.4; the slot indices are compile-time constants (rows processed in groups of 7).
template <bool FLIP, int L>
__global__ void __launch_bounds__(DW_CB) dwconv7_kernel(int H, int W, int C, int ncb, int rows, const float* __restrict__ x,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        const float* __restrict__ res, float* __restrict__ y) {
    const int c = (blockIdx.x % ncb)*DW_CB + threadIdx.x;
    if (c >= C) return;
    const int x0 = (blockIdx.x/ncb)*L, y0 = blockIdx.y*rows, n = blockIdx.z;
    const int nrows = min(rows, H - y0);  // output rows of this block
    float wr[49];
#pragma unroll
    for (int t = 0; t < 49; ++t) wr[t] = __ldg(w + (size_t)c*49 + (FLIP ? 48 - t : t));
    const float b = bias ? __ldg(bias + c) : 0.f;
    const size_t img = (size_t)n*H*W*C;
    // Column offsets (in elements, relative to the start of an image row of this channel) of the L + 6 input columns,
    // computed once: -1 marks a column outside the image. Every load / store below is then base + 32-bit offset.
    int coff[L + 6];
#pragma unroll
    for (int q = 0; q < L + 6; ++q) {
        const int xx = x0 + q - 3;
        coff[q] = (xx >= 0 && xx < W) ? xx*C : -1;
    }
    float acc[7][L];
#pragma unroll
    for (int sl = 0; sl < 7; ++sl)
#pragma unroll
        for (int j = 0; j < L; ++j) acc[sl][j] = b;
    // The residual (`res`, the data-gradient variant) is folded into the accumulator's initial value: it is requested when the
    // row's slot is recycled, one input row before the row's first product, instead of as a dependent load at the store.
    if (res != nullptr) {
        const size_t rowoff = img + (size_t)y0*W*C + c;
#pragma unroll
        for (int j = 0; j < L; ++j) if (coff[j + 3] >= 0) acc[0][j] = b + __ldg(res + rowoff + coff[j + 3]);
    }
    // input row i (image row y0 - 3 + i) feeds output rows o = i - ky (ky = 0..6), kept in slot o mod 7
    for (int g = 0; g*7 < nrows + 6; ++g) {
#pragma unroll
        for (int jr = 0; jr < 7; ++jr) {
            const int i = g*7 + jr;
            if (i >= nrows + 6) break;
            const int yin = y0 - 3 + i;
            if (yin >= 0 && yin < H) {
                const float* row = x + img + (size_t)yin*W*C + c;
                float v[L + 6];
#pragma unroll
                for (int q = 0; q < L + 6; ++q) v[q] = coff[q] >= 0 ? __ldg(row + coff[q]) : 0.f;
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    const int o = i - ky;
                    if (o < 0 || o >= nrows) continue;
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx)
#pragma unroll
                        for (int j = 0; j < L; ++j) acc[(jr - ky + 7) % 7][j] = fmaf(wr[ky*7 + kx], v[j + kx], acc[(jr - ky + 7) % 7][j]);
                }
            }
            const int o = i - 6;  // complete: its last input row (ky = 6) was row i
            if (o >= 0 && o < nrows) {
                const size_t rowoff = img + (size_t)(y0 + o)*W*C + c;
#pragma unroll
                for (int j = 0; j < L; ++j)
                    if (coff[j + 3] >= 0) y[rowoff + coff[j + 3]] = acc[(jr + 1) % 7][j];
            }
            // the slot of row o is reused by row o + 7 = i + 1
            const int on = i + 1;
            if (res != nullptr && on < nrows) {
                const size_t rowoff = img + (size_t)(y0 + on)*W*C + c;
#pragma unroll
                for (int j = 0; j < L; ++j) acc[(jr + 1) % 7][j] = coff[j + 3] >= 0 ? b + __ldg(res + rowoff + coff[j + 3]) : b;
            } else {
#pragma unroll
                for (int j = 0; j < L; ++j) acc[(jr + 1) % 7][j] = b;
            }
        }
    }
}

// Weight gradient: gw[c][ky][kx] = sum_{n,y,x} gy[n,y,x,c] * x[n,y+ky-3,x+kx-3,c];  gb[c] = sum gy.
// Each block owns a (WG_RY rows x WG_XW columns) patch of one image; a thread owns one channel and slides a 7x7 register
// window of x along the row. Per-block partials go to the workspace; a second kernel adds them in a fixed order.
constexpr int WG_RY = 4, WG_XW = 32, WG_U = 4;

__global__ void __launch_bounds__(DW_CB) dwconv7_wgrad_kernel(int H, int W, int C, int ncb, const float* __restrict__ x,
                                                              const float* __restrict__ gy, float* __restrict__ partial) {
    const int c = (blockIdx.x % ncb)*DW_CB + threadIdx.x;
    if (c >= C) return;
    const int xb = (blockIdx.x/ncb)*WG_XW, yb = blockIdx.y*WG_RY, n = blockIdx.z;
    const size_t img = (size_t)n*H*W*C;
    float acc[49], gsum = 0.f;
#pragma unroll
    for (int t = 0; t < 49; ++t) acc[t] = 0.f;
    for (int r = 0; r < WG_RY; ++r) {
        const int yo = yb + r;
        if (yo >= H) break;
        // WG_U output pixels per iteration: their 7 x WG_U new window columns and WG_U gradients are loaded first (independent
        // loads in flight), then 49 x WG_U FMAs run on the (7 + WG_U - 1)-wide register window; the window shifts once per iteration.
        float win[7][6 + WG_U];  // win[ky][q] = x[yo+ky-3][xo+q-3] for the iteration starting at column xo
        const float* rowp[7];    // start of input row yo+ky-3 for this channel (nullptr: outside the image -> zeros)
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
            const int yy = yo + ky - 3;
            rowp[ky] = (yy >= 0 && yy < H) ? x + img + (size_t)yy*W*C + c : nullptr;
        }
        // one column offset (xx*C) and validity test per column, shared by the 7 rows of the window
        auto ldcol = [&](int xx, int q) {
            const bool okx = xx >= 0 && xx < W;
            const int off = xx*C;
#pragma unroll
            for (int ky = 0; ky < 7; ++ky) win[ky][q] = (okx && rowp[ky]) ? __ldg(rowp[ky] + off) : 0.f;
        };
#pragma unroll
        for (int q = 0; q < 6; ++q) ldcol(xb + q - 3, q);
        const int xe = min(xb + WG_XW, W);
        for (int xo = xb; xo < xe; xo += WG_U) {
#pragma unroll
            for (int u = 0; u < WG_U; ++u) ldcol(xo + 3 + u, 6 + u);
            float g[WG_U];
#pragma unroll
            for (int u = 0; u < WG_U; ++u) g[u] = xo + u < xe ? __ldg(gy + img + ((size_t)yo*W + xo + u)*C + c) : 0.f;
#pragma unroll
            for (int u = 0; u < WG_U; ++u) {
                gsum += g[u];
#pragma unroll
                for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 7; ++kx) acc[ky*7 + kx] = fmaf(g[u], win[ky][u + kx], acc[ky*7 + kx]);
            }
#pragma unroll
            for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                for (int q = 0; q < 6; ++q) win[ky][q] = win[ky][q + WG_U];
        }
    }
    const size_t blk = ((size_t)blockIdx.z*gridDim.y + blockIdx.y)*(gridDim.x/ncb) + blockIdx.x/ncb;  // spatial patch index
    float* out = partial + blk*50*C;
#pragma unroll
    for (int t = 0; t < 49; ++t) out[(size_t)t*C + c] = acc[t];
    out[(size_t)49*C + c] = gsum;
}

// Strip variant (round 2): the forward kernel's vertical sweep applied to the weight gradient. A block owns DW_L columns x `rows`
// gradient rows of one image; the thread's 49 partial sums stay in registers for the whole strip. Input row i (image row
// y0 - 3 + i) is loaded ONCE (DW_L + 6 values) and multiplied with the 7 gradient rows o = i - ky it pairs with, which wait in a
// 7-slot register ring (each gradient row is loaded once, too): 2 DW_L + 6 loads per 49 DW_L FMAs instead of 8 per 49 in the
// window kernel above, and one partial record per 8 x rows pixels instead of one per 4 x 32.
template <int L>
__global__ void __launch_bounds__(DW_CB) dwconv7_wgrad_strip_kernel(int H, int W, int C, int ncb, int rows, const float* __restrict__ x,
                                                                    const float* __restrict__ gy, float* __restrict__ partial) {
    const int c = (blockIdx.x % ncb)*DW_CB + threadIdx.x;
    if (c >= C) return;
    const int x0 = (blockIdx.x/ncb)*L, y0 = blockIdx.y*rows, n = blockIdx.z;
    const int nrows = min(rows, H - y0);
    const size_t img = (size_t)n*H*W*C;
    int coff[L + 6];
#pragma unroll
    for (int q = 0; q < L + 6; ++q) {
        const int xx = x0 + q - 3;
        coff[q] = (xx >= 0 && xx < W) ? xx*C : -1;
    }
    float acc[49], gsum = 0.f;
#pragma unroll
    for (int t = 0; t < 49; ++t) acc[t] = 0.f;
    float g[7][L];
#pragma unroll
    for (int sl = 0; sl < 7; ++sl)
#pragma unroll
        for (int j = 0; j < L; ++j) g[sl][j] = 0.f;
    for (int gi = 0; gi*7 < nrows + 6; ++gi) {
#pragma unroll
        for (int jr = 0; jr < 7; ++jr) {
            const int i = gi*7 + jr;
            if (i >= nrows + 6) break;
            // gradient row i enters slot i mod 7 (the row it replaces, i - 7, met its last input row one step ago)
            if (i < nrows) {
                const float* grow = gy + img + (size_t)(y0 + i)*W*C + c;
#pragma unroll
                for (int j = 0; j < L; ++j) { g[jr][j] = coff[j + 3] >= 0 ? __ldg(grow + coff[j + 3]) : 0.f; gsum += g[jr][j]; }
            } else {
#pragma unroll
                for (int j = 0; j < L; ++j) g[jr][j] = 0.f;
            }
            const int yin = y0 - 3 + i;
            if (yin < 0 || yin >= H) continue;
            const float* row = x + img + (size_t)yin*W*C + c;
            float v[L + 6];
#pragma unroll
            for (int q = 0; q < L + 6; ++q) v[q] = coff[q] >= 0 ? __ldg(row + coff[q]) : 0.f;
#pragma unroll
            for (int ky = 0; ky < 7; ++ky) {
                const int o = i - ky;   // gradient row paired with this input row through filter row ky
                if (o < 0 || o >= nrows) continue;
#pragma unroll
                for (int kx = 0; kx < 7; ++kx)
#pragma unroll
                    for (int j = 0; j < L; ++j) acc[ky*7 + kx] = fmaf(g[(jr - ky + 7) % 7][j], v[j + kx], acc[ky*7 + kx]);
            }
        }
    }
    const size_t blk = ((size_t)blockIdx.z*gridDim.y + blockIdx.y)*(gridDim.x/ncb) + blockIdx.x/ncb;
    float* out = partial + blk*50*C;
#pragma unroll
    for (int t = 0; t < 49; ++t) out[(size_t)t*C + c] = acc[t];
    out[(size_t)49*C + c] = gsum;
}

// out[c*49 + t] = sum_blk partial[blk][t][c];  gb[c] = sum_blk partial[blk][49][c]. A block owns 32 consecutive entries (t, c);
// its 8 warps stride over the partial blocks (8 independent, coalesced load streams) and are combined in a fixed order.
__global__ void __launch_bounds__(256) dwconv7_wgrad_reduce_kernel(int C, int nblk, const float* __restrict__ partial,
                                                                   float* __restrict__ gw, float* __restrict__ gb, int accumulate) {
    __shared__ double red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int e = blockIdx.x*32 + tx;
    double a = 0.0;
    if (e < 50*C)
        for (int b = ty; b < nblk; b += 8) a += (double)__ldg(partial + (size_t)b*50*C + e);
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && e < 50*C) {
#pragma unroll
        for (int k = 1; k < 8; ++k) a += red[k][tx];
        const int t = e/C, c = e - t*C;
        if (t < 49) { float* o = gw + (size_t)c*49 + t; *o = accumulate ? *o + (float)a : (float)a; }
        else if (gb) gb[c] = accumulate ? gb[c] + (float)a : (float)a;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm over the last (channel) axis of a (P, C) matrix. One warp per row; rows are re-read from L1 between passes.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(long long P, int C, float eps, const float* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ y, float* __restrict__ mean,
                                                            float* __restrict__ rstd) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= P) return;
    const float* xr = x + row*C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mu = warp_sum(s)/(float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mu; v = fmaf(d, d, v); }
    const float rs = rsqrtf(warp_sum(v)/(float)C + eps);
    float* yr = y + row*C;
    for (int c = lane; c < C; c += 32) yr[c] = fmaf((xr[c] - mu)*rs, __ldg(gamma + c), __ldg(beta + c));
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

// Register-resident variant (C % 4 == 0, C <= 128*NV): a lane holds its NV float4 of the row, so the row is read from memory ONCE
// (all loads issued up front) and the three passes (mean, variance, normalise) run on registers.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_fwd_vec_kernel(long long P, int C, float eps, const float* __restrict__ x,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= P) return;
    const int c4 = C >> 2;
    const float4* xr = (const float4*)(x + row*C);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = lane + 32*i < c4 ? __ldg(xr + lane + 32*i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mu = warp_sum(s)/(float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (lane + 32*i < c4) {
            const float a = v[i].x - mu, b = v[i].y - mu, c = v[i].z - mu, d = v[i].w - mu;
            q += fmaf(a, a, b*b) + fmaf(c, c, d*d);
        }
    const float rs = rsqrtf(warp_sum(q)/(float)C + eps);
    float4* yr = (float4*)(y + row*C);
#pragma unroll
    for (int i = 0; i < NV; ++i)
        if (lane + 32*i < c4) {
            const float4 g = __ldg((const float4*)gamma + lane + 32*i), b = __ldg((const float4*)beta + lane + 32*i);
            yr[lane + 32*i] = make_float4(fmaf((v[i].x - mu)*rs, g.x, b.x), fmaf((v[i].y - mu)*rs, g.y, b.y),
                                          fmaf((v[i].z - mu)*rs, g.z, b.z), fmaf((v[i].w - mu)*rs, g.w, b.w));
        }
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

// dx = rstd * (dy*gamma - mean_c(dy*gamma) - xhat * mean_c(dy*gamma*xhat));  dgamma = sum_rows dy*xhat;  dbeta = sum_rows dy.
// Each warp walks `rows` (<= LN_ROWS) consecutive rows and keeps its lanes' dgamma/dbeta slices in registers (C <= 32*LN_MAXPL).
// `rows` shrinks for short matrices (the deep ConvNeXt stages have only 2-8 k rows) so that the grid still covers the SMs.
constexpr int LN_ROWS = 32, LN_MAXPL = 32;

template <int PL>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(long long P, int C, int rows, const float* __restrict__ dy,
                                                            const float* __restrict__ x, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            float* __restrict__ dx, float* __restrict__ partial) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const long long row0 = ((long long)blockIdx.x*nw + wid)*rows;
    float dg[PL], db[PL], g[PL];
#pragma unroll
    for (int i = 0; i < PL; ++i) { dg[i] = db[i] = 0.f; const int c = lane + 32*i; g[i] = c < C ? __ldg(gamma + c) : 0.f; }
    for (int r = 0; r < rows; ++r) {
        const long long row = row0 + r;
        if (row >= P) break;
        const float mu = mean[row], rs = rstd[row];
        const float* xr = x + row*C;
        const float* dr = dy + row*C;
        float xh[PL], dyv[PL], s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < PL; ++i) {
            const int c = lane + 32*i;
            const bool ok = c < C;
            xh[i] = ok ? (xr[c] - mu)*rs : 0.f;
            dyv[i] = ok ? dr[c] : 0.f;
            const float t = dyv[i]*g[i];
            s1 += t; s2 = fmaf(t, xh[i], s2);
            dg[i] = fmaf(dyv[i], xh[i], dg[i]); db[i] += dyv[i];
        }
        s1 = warp_sum(s1)/(float)C; s2 = warp_sum(s2)/(float)C;
        float* dxr = dx + row*C;
#pragma unroll
        for (int i = 0; i < PL; ++i) {
            const int c = lane + 32*i;
            if (c < C) dxr[c] = rs*(dyv[i]*g[i] - s1 - xh[i]*s2);
        }
    }
    // Block partial: the 8 warps are added in a fixed order through shared memory, one row [2][C] per block.
    extern __shared__ float sred[];  // [nw][2*C]
#pragma unroll
    for (int i = 0; i < PL; ++i) {
        const int c = lane + 32*i;
        if (c < C) { sred[(size_t)wid*2*C + c] = dg[i]; sred[(size_t)wid*2*C + C + c] = db[i]; }
    }
    __syncthreads();
    float* out = partial + (size_t)blockIdx.x*2*C;
    for (int e = threadIdx.x; e < 2*C; e += blockDim.x) {
        float a = 0.f;
        for (int w2 = 0; w2 < nw; ++w2) a += sred[(size_t)w2*2*C + e];
        out[e] = a;
    }
}

// A block owns 32 consecutive entries of [dgamma | dbeta]; its 8 warps stride over the partial rows (8 coalesced load streams in
// flight instead of one serial chain) and are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256) layernorm_bwd_reduce_kernel(int C, int nrows, const float* __restrict__ partial,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
    __shared__ double red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int e = blockIdx.x*32 + tx;
    double a0 = 0.0, a1 = 0.0;
    if (e < 2*C) {
        int r = ty;
        for (; r + 8 < nrows; r += 16) { a0 += (double)__ldg(partial + (size_t)r*2*C + e); a1 += (double)__ldg(partial + (size_t)(r + 8)*2*C + e); }
        if (r < nrows) a0 += (double)__ldg(partial + (size_t)r*2*C + e);
    }
    red[ty][tx] = a0 + a1;
    __syncthreads();
    if (ty == 0 && e < 2*C) {
        double a = red[0][tx];
#pragma unroll
        for (int k = 1; k < 8; ++k) a += red[k][tx];
        float* o = e < C ? dgamma + e : dbeta + (e - C);
        *o = accumulate ? *o + (float)a : (float)a;
    }
}

}  // namespace stv

using namespace stv;

static int round32(int c) { return (c + 31)/32*32; }

// Rows per block of the strip kernels: tall strips amortise the 6-row halo; shrink them while the grid is smaller than ~6 blocks
// per SM (blocks are only 3-4 warps).
// Rows per block: tall strips amortise the 6 warm-up rows; enough blocks for ~6 per SM. (Measured and rejected, profiles/r2_rejected_variants.txt:
// 4 instead of 8 columns per thread for the small deep-stage maps — four times the blocks, but 0.39 vs 0.26 ms for the twelve
// stage-2/3 data gradients: the halo loads per output grow from 1.75x to 2.5x and the 49 weights are re-read by twice the threads.)
static int dw_rows(int N, int H, int W, int C) {
    const int ncb = (C + DW_CB - 1)/DW_CB;
    const long long cols = (long long)((W + DW_L - 1)/DW_L)*ncb*N;
    int rows = 32;
    while (rows > 8 && cols*((H + rows - 1)/rows) < 6*148) rows >>= 1;
    return rows;
}

extern "C" int stv_dwconv7_fwd(int N, int H, int W, int C, const float* x, const float* w, const float* bias, const float* res,
                               float* y, int flip, void* stream) {
    STV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 1024, "stv_dwconv7_fwd: bad shape (N=%d H=%d W=%d C=%d; C <= 1024)", N, H, W, C);
    STV_REQUIRE(N <= 65535, "stv_dwconv7_fwd: grid too large");
    STV_REQUIRE(x && w && y, "stv_dwconv7_fwd: NULL pointer");
    const int ncb = (C + DW_CB - 1)/DW_CB, nt = C < DW_CB ? round32(C) : DW_CB;
    const int rows = dw_rows(N, H, W, C);
    dim3 grid(((W + DW_L - 1)/DW_L)*ncb, (H + rows - 1)/rows, N);
    if (flip) dwconv7_kernel<true, DW_L><<<grid, nt, 0, (cudaStream_t)stream>>>(H, W, C, ncb, rows, x, w, bias, res, y);
    else dwconv7_kernel<false, DW_L><<<grid, nt, 0, (cudaStream_t)stream>>>(H, W, C, ncb, rows, x, w, bias, res, y);
    count_launch();
    return check_launch("dwconv7_kernel");
}

extern "C" size_t stv_dwconv7_wgrad_workspace_bytes(int N, int H, int W, int C) {
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
    const size_t window = (size_t)N*((H + WG_RY - 1)/WG_RY)*((W + WG_XW - 1)/WG_XW);
    const size_t strip = (size_t)N*((H + 7)/8)*((W + DW_L - 1)/DW_L);   // strips are at least 8 rows tall
    return (window > strip ? window : strip)*50*C*sizeof(float);
}

// Weight-gradient variant: STV_DW_WGRAD=window selects the round-1 window kernel (developer A/B switch).
static bool dw_wgrad_strip() {
    static const bool v = !(getenv("STV_DW_WGRAD") && getenv("STV_DW_WGRAD")[0] == 'w');
    return v;
}

extern "C" int stv_dwconv7_wgrad(int N, int H, int W, int C, const float* x, const float* gy, float* gw, float* gb, int accumulate,
                                 void* ws, size_t ws_bytes, void* stream) {
    STV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 1024, "stv_dwconv7_wgrad: bad shape");
    STV_REQUIRE(N <= 65535 && (H + WG_RY - 1)/WG_RY <= 65535, "stv_dwconv7_wgrad: grid too large");
    STV_REQUIRE(x && gy && gw, "stv_dwconv7_wgrad: NULL pointer");
    const size_t need = stv_dwconv7_wgrad_workspace_bytes(N, H, W, C);
    if (!ws || ws_bytes < need) { set_error("stv_dwconv7_wgrad: workspace too small (%zu < %zu bytes)", ws_bytes, need); return STV_E_WORKSPACE; }
    const int ncb = (C + DW_CB - 1)/DW_CB, nt = C < DW_CB ? round32(C) : DW_CB;
    dim3 grid(((W + WG_XW - 1)/WG_XW)*ncb, (H + WG_RY - 1)/WG_RY, N);
    if (dw_wgrad_strip()) {
        const int rows = dw_rows(N, H, W, C);
        grid = dim3(((W + DW_L - 1)/DW_L)*ncb, (H + rows - 1)/rows, N);
        dwconv7_wgrad_strip_kernel<DW_L><<<grid, nt, 0, (cudaStream_t)stream>>>(H, W, C, ncb, rows, x, gy, (float*)ws);
    } else dwconv7_wgrad_kernel<<<grid, nt, 0, (cudaStream_t)stream>>>(H, W, C, ncb, x, gy, (float*)ws);
    count_launch();
    if (int rc = check_launch("dwconv7_wgrad_kernel")) return rc;
    const int nblk = (grid.x/ncb)*grid.y*grid.z;
    dwconv7_wgrad_reduce_kernel<<<(50*C + 31)/32, 256, 0, (cudaStream_t)stream>>>(C, nblk, (const float*)ws, gw, gb, accumulate);
    count_launch();
    return check_launch("dwconv7_wgrad_reduce_kernel");
}

extern "C" int stv_layernorm_fwd(long long P, int C, const float* x, const float* gamma, const float* beta, float eps, float* y,
                                 float* mean, float* rstd, void* stream) {
    STV_REQUIRE(P > 0 && C > 0, "stv_layernorm_fwd: bad shape");
    STV_REQUIRE(x && gamma && beta && y && mean && rstd, "stv_layernorm_fwd: NULL pointer");
    const long long blocks = (P + 7)/8;
    STV_REQUIRE(blocks < (1ll << 31), "stv_layernorm_fwd: too many rows");
    const bool vec = C % 4 == 0 && C <= 1024 && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0;
    const int nv = (C/4 + 31)/32;
    cudaStream_t st = (cudaStream_t)stream;
    if (vec && nv <= 1) layernorm_fwd_vec_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(P, C, eps, x, gamma, beta, y, mean, rstd);
    else if (vec && nv == 2) layernorm_fwd_vec_kernel<2><<<(unsigned)blocks, 256, 0, st>>>(P, C, eps, x, gamma, beta, y, mean, rstd);
    else if (vec && nv <= 4) layernorm_fwd_vec_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(P, C, eps, x, gamma, beta, y, mean, rstd);
    else if (vec && nv <= 8) layernorm_fwd_vec_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(P, C, eps, x, gamma, beta, y, mean, rstd);
    else layernorm_fwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(P, C, eps, x, gamma, beta, y, mean, rstd);
    count_launch();
    return check_launch("layernorm_fwd_kernel");
}

// Rows per warp: LN_ROWS for long matrices, fewer while the grid would be smaller than ~2 blocks (of 8 warps) per SM.
static int ln_bwd_rows(long long P) {
    int rows = LN_ROWS;
    while (rows > 1 && (P + rows - 1)/rows < 2ll*148*8) rows >>= 1;
    return rows;
}
static long long ln_bwd_warps(long long P) { const int r = ln_bwd_rows(P); return (P + r - 1)/r; }

extern "C" size_t stv_layernorm_bwd_workspace_bytes(long long P, int C) {
    if (P <= 0 || C <= 0) return 0;
    const long long blocks = (ln_bwd_warps(P) + 7)/8;
    return (size_t)blocks*2*C*sizeof(float);
}

extern "C" int stv_layernorm_bwd(long long P, int C, const float* dy, const float* x, const float* mean, const float* rstd,
                                 const float* gamma, float* dx, float* dgamma, float* dbeta, int accumulate, void* ws,
                                 size_t ws_bytes, void* stream) {
    STV_REQUIRE(P > 0 && C > 0 && C <= 32*LN_MAXPL, "stv_layernorm_bwd: bad shape (C <= %d)", 32*LN_MAXPL);
    STV_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta, "stv_layernorm_bwd: NULL pointer");
    const size_t need = stv_layernorm_bwd_workspace_bytes(P, C);
    if (!ws || ws_bytes < need) { set_error("stv_layernorm_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need); return STV_E_WORKSPACE; }
    const long long blocks = (ln_bwd_warps(P) + 7)/8;
    STV_REQUIRE(blocks < (1ll << 31), "stv_layernorm_bwd: too many rows");
    cudaStream_t st = (cudaStream_t)stream;
    // Rows beyond P inside the last block write zero partials (their accumulators stay 0), so every slot is defined.
    const int pl = (C + 31)/32;
    static bool attr_done = false;
    if (!attr_done) {  // C > 768 needs more than the default 48 KB of dynamic shared memory for the block reduction
        const int mx = 8*2*32*LN_MAXPL*(int)sizeof(float);
        cudaFuncSetAttribute(layernorm_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        cudaFuncSetAttribute(layernorm_bwd_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx);
        attr_done = true;
    }
#define STV_LN_BWD(PL) layernorm_bwd_kernel<PL><<<(unsigned)blocks, 256, (size_t)8*2*C*sizeof(float), st>>>(P, C, ln_bwd_rows(P), dy, x, mean, rstd, gamma, dx, (float*)ws)
    if (pl <= 3) STV_LN_BWD(3); else if (pl <= 4) STV_LN_BWD(4); else if (pl <= 6) STV_LN_BWD(6); else if (pl <= 8) STV_LN_BWD(8);
    else if (pl <= 12) STV_LN_BWD(12); else if (pl <= 16) STV_LN_BWD(16); else if (pl <= 24) STV_LN_BWD(24); else STV_LN_BWD(32);
#undef STV_LN_BWD
    count_launch();
    if (int rc = check_launch("layernorm_bwd_kernel")) return rc;
    layernorm_bwd_reduce_kernel<<<(2*C + 31)/32, 256, 0, st>>>(C, (int)blocks, (const float*)ws, dgamma, dbeta, accumulate);
    count_launch();
    return check_launch("layernorm_bwd_reduce_kernel");
}
