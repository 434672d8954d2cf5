// Fused view-synthesis photometric loss (forward + backward) for sm_100a.
//
// One launch covers every (scale s, image i) pair of a batch: grid = (tiles, b, S). A block owns a TH x TW pixel tile
// of the target frame and walks the n support frames:
//   phase 1  back-project -> rigid transform -> project -> bilinear border sample (rows 9-11 of SURVEY 8a) for the tile
//            plus its SSIM halo, straight into shared memory (the warped image never exists in HBM);
//   phase 2  3x3 reflect-padded SSIM + L1 from shared memory (rows 12-13), min / mean over support frames and the
//            auto-mask comparison against the pre-computed identity error (row 14), block partial sum of the loss.
// The backward kernel re-warps with a halo of 2, rebuilds the SSIM window sums, pushes the loss gradient through the
// box filters (adjoint of reflect pad + 3x3 mean), the bilinear sampler and the projection, and emits d/d depth per
// pixel plus per-block partials of d/dT, d/dK, d/dKinv that a finalize kernel adds up in a fixed order (deterministic).
//
// HBM traffic per target pixel (fp32): target 12 B + depth 4 B per scale + gathered support texels (L1/L2 resident
// between neighbouring pixels) + 1 B decision byte per scale; see DESIGN.md for the roofline accounting.
#include "stv_common.cuh"
#include "stv_f2.cuh"

namespace stv {

constexpr int TW = 64;    // tile width  (two warps per tile row)
constexpr int TH = 16;    // tile height
constexpr int NT = 256;   // threads per block
constexpr int RUN = 4;    // consecutive rows handled by one thread in the per-pixel phases (TH*TW == NT*RUN)
constexpr int PW1 = TW + 2, PH1 = TH + 2;  // tile + halo 1
constexpr int PW2 = TW + 4, PH2 = TH + 4;  // tile + halo 2
static_assert(TH*TW == NT*RUN, "tile/threads mismatch");

struct PhotoParams {
    int b, n, S, H, W;
    float w_ssim, w_l1;
    int use_min, use_automask;
    uint64_t seed;
    int tiles_x, tiles_y;
    const float* depth[STV_MAX_SCALES];
    float* g_depth[STV_MAX_SCALES];
    const float *tgt, *supp, *T, *K, *Kinv, *noise, *e0, *grad_loss;
    const uint8_t* sel_in;
    const unsigned long long* step;  // device-side call counter added to `seed` (advanced by the loss reduction kernel)
    float* partial;   // fwd: one float per block; bwd: n_acc floats per (block, k)
    uint8_t* sel;
    float* warp0;
    float* coef;             // fwd out: (S,b,9,H,W) masked d err/d(S1,S2,S3) per channel of the SELECTED support (nullable)
    const float* coef_in;    // bwd in
    int coef_tma;            // bwd: coefficient tiles arrive by TMA (row pitch 16-byte granular)
    cudaTextureObject_t supp_tex;  // supp viewed as one (n*b*3*H) x W single-channel texture (0 = not available)
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// ---------------------------------------------------------------------------------------------------------------------
// Shared per-pixel machinery
// ---------------------------------------------------------------------------------------------------------------------
// Loads a (rows x cols) window of the 3-channel image `img` (one sample, planes of HW floats) whose top-left padded
// coordinate is (y0, x0) into smem planes dst[c][rows][cols]; coordinates are reflected (pad 1) and then clamped.
template <int ROWS, int COLS>
__device__ __forceinline__ void load_tile3(float (*dst)[ROWS][COLS], const float* __restrict__ img, int y0, int x0, int H,
                                           int W) {
    const int HW = H*W;
    for (int q = threadIdx.x; q < ROWS*COLS; q += NT) {
        const int py = q/COLS, px = q - py*COLS;
        const int ya = clampi(reflect_idx(y0 + py, H), 0, H - 1), xa = clampi(reflect_idx(x0 + px, W), 0, W - 1);
        const float* p = img + ya*W + xa;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[c][py][px] = __ldg(p + c*HW);
    }
}

// Window sums of a 3x3 neighbourhood for RUN vertically consecutive centres; `top` is the smem row of the first
// centre's upper neighbour, `col` the smem column of the left neighbour.
template <int ROWS, int COLS, int NRUN>
__device__ __forceinline__ void target_sums(const float (*st)[COLS], int top, int col, float* T1, float* T2) {
    float h1[NRUN + 2], h2[NRUN + 2];
#pragma unroll
    for (int r = 0; r < NRUN + 2; ++r) {
        const float a = st[top + r][col], b = st[top + r][col + 1], c = st[top + r][col + 2];
        h1[r] = a + b + c;
        h2[r] = fmaf(a, a, fmaf(b, b, c*c));
    }
#pragma unroll
    for (int j = 0; j < NRUN; ++j) {
        T1[j] = h1[j] + h1[j + 1] + h1[j + 2];
        T2[j] = h2[j] + h2[j + 1] + h2[j + 2];
    }
}

template <int COLS, int NRUN>
__device__ __forceinline__ void pair_sums(const float (*sw)[COLS], const float (*st)[COLS], int top, int col, float* S1,
                                          float* S2, float* S3) {
    float h1[NRUN + 2], h2[NRUN + 2], h3[NRUN + 2];
#pragma unroll
    for (int r = 0; r < NRUN + 2; ++r) {
        const float a = sw[top + r][col], b = sw[top + r][col + 1], c = sw[top + r][col + 2];
        const float ta = st[top + r][col], tb = st[top + r][col + 1], tc = st[top + r][col + 2];
        h1[r] = a + b + c;
        h2[r] = fmaf(a, a, fmaf(b, b, c*c));
        h3[r] = fmaf(a, ta, fmaf(b, tb, c*tc));
    }
#pragma unroll
    for (int j = 0; j < NRUN; ++j) {
        S1[j] = h1[j] + h1[j + 1] + h1[j + 2];
        S2[j] = h2[j] + h2[j + 1] + h2[j + 2];
        S3[j] = h3[j] + h3[j + 1] + h3[j + 2];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward-side machinery (v2): transposed shared tiles + packed fp32x2 math + texture-gather sampling.
//
// Shared tiles are stored [channel][column][row] so that the RUN+2 = 6 vertically consecutive rows a thread needs from one
// column are 3 aligned 8-byte words: LDS.64 loads whose two lanes are exactly the row pairs the packed FADD2/FFMA2 box-filter
// arithmetic consumes. With a column stride of PH1 = 18 words, the 16 lanes of an LDS.64 phase hit all 32 banks once.
// ---------------------------------------------------------------------------------------------------------------------
// Forward-side tiling: many small blocks per SM (each block alternates a memory phase and a compute phase separated by
// barriers, so independent blocks are what overlaps the two).
#ifndef STV_FWD_TH
#define STV_FWD_TH 8
#endif
#ifndef STV_FWD_NT
#define STV_FWD_NT 256
#endif
constexpr int FTW = 64, FTH = STV_FWD_TH;    // forward tile
constexpr int FNT = STV_FWD_NT;              // threads per block of the forward-side kernels
constexpr int FRUN = FTH*FTW/FNT;            // vertically consecutive centres per thread (2 or 4)
constexpr int FNP = FRUN/2;                  // packed centre pairs per thread
static_assert(FRUN == 2 || FRUN == 4, "forward kernels support 2 or 4 centres per thread");
constexpr int FPW = FTW + 2, FPH = FTH + 2;  // tile + halo 1
constexpr int NPOS = (FPH*FPW + FNT - 1)/FNT;  // halo-1 positions per thread in phase 1
typedef float TileT[FPW][FPH];               // one channel, transposed

// Window sums (3x3, as packed row pairs) of one channel for the FRUN centres of a thread:
//   rows top..top+FRUN+1 of columns col..col+2 ; centres are rows top+1..top+FRUN of column col+1.
struct WinSums { f2 S1[FNP], S2[FNP], S3[FNP]; f2 wc[FNP + 1], tc[FNP + 1]; };

__device__ __forceinline__ f2 mid2(f2 a, f2 b) { return mk2(hi2(a), lo2(b)); }

__device__ __forceinline__ void pair_sums_t(const TileT& sw, const TileT& st, int top, int col, WinSums& o) {
    f2 h1[FNP + 1], h2[FNP + 1], h3[FNP + 1];
#pragma unroll
    for (int j = 0; j < FNP + 1; ++j) {
        const f2 a = ld2(&sw[col][top + 2*j]), b = ld2(&sw[col + 1][top + 2*j]), c = ld2(&sw[col + 2][top + 2*j]);
        const f2 ta = ld2(&st[col][top + 2*j]), tb = ld2(&st[col + 1][top + 2*j]), tc = ld2(&st[col + 2][top + 2*j]);
        h1[j] = a + b + c;
        h2[j] = fma2(a, a, fma2(b, b, c*c));
        h3[j] = fma2(a, ta, fma2(b, tb, c*tc));
        o.wc[j] = b; o.tc[j] = tb;
    }
#pragma unroll
    for (int j = 0; j < FNP; ++j) {
        o.S1[j] = h1[j] + h1[j + 1] + mid2(h1[j], h1[j + 1]);
        o.S2[j] = h2[j] + h2[j + 1] + mid2(h2[j], h2[j + 1]);
        o.S3[j] = h3[j] + h3[j + 1] + mid2(h3[j], h3[j + 1]);
    }
}

__device__ __forceinline__ void target_sums_t(const TileT& st, int top, int col, f2* T1, f2* T2) {
    f2 h1[FNP + 1], h2[FNP + 1];
#pragma unroll
    for (int j = 0; j < FNP + 1; ++j) {
        const f2 a = ld2(&st[col][top + 2*j]), b = ld2(&st[col + 1][top + 2*j]), c = ld2(&st[col + 2][top + 2*j]);
        h1[j] = a + b + c;
        h2[j] = fma2(a, a, fma2(b, b, c*c));
    }
#pragma unroll
    for (int j = 0; j < FNP; ++j) {
        T1[j] = h1[j] + h1[j + 1] + mid2(h1[j], h1[j + 1]);
        T2[j] = h2[j] + h2[j + 1] + mid2(h2[j], h2[j + 1]);
    }
}

// SSIM error of a pixel pair (packed), clamp to [0,1] by the saturating FMA.  (src/losses/photometric.py:40-50)
__device__ __forceinline__ void ssim_pair(f2 S1, f2 S2, f2 S3, f2 T1, f2 T2, float& e0, float& e1) {
    const f2 k = splat2(1.f/9.f), two = splat2(2.f), nk = splat2(-1.f/9.f);
    const f2 mx = S1*k, my = T1*k;
    const f2 mxy = mx*my, mxx = mx*mx, myy = my*my;
    // -(sigma) = mean^2 - E[.]  keeps everything in FFMA2 form (no packed negate exists)
    const f2 nsxx = fma2(S2, nk, mxx), nsyy = fma2(T2, nk, myy), nsxy = fma2(S3, nk, mxy);
    const f2 num = fma2(two, mxy, splat2(STV_C1))*fma2(splat2(-2.f), nsxy, splat2(STV_C2));
    const f2 den = (mxx + myy + splat2(STV_C1))*fma2(splat2(-1.f), nsxx + nsyy, splat2(STV_C2));
    e0 = __saturatef(fmaf(-0.5f*lo2(num), rcp_fast(lo2(den)), 0.5f));
    e1 = __saturatef(fmaf(-0.5f*hi2(num), rcp_fast(hi2(den)), 0.5f));
}

// Same, plus d err/d(S1,S2,S3) of both pixels (zero where the clamp is active, as ssim_err_grad) sharing num/den/rcp:
//   r = A1 A2/(B1 B2);  a = k (r mx (B2-B1) - my (A2-A1))/(B1 B2);  b = k r /(2 B2);  c = -k A1/(B1 B2),  k = 1/9.
__device__ __forceinline__ void ssim_pair_coef(f2 S1, f2 S2, f2 S3, f2 T1, f2 T2, float& e0, float& e1, f2& ca, f2& cb, f2& cc) {
    const f2 k = splat2(1.f/9.f), two = splat2(2.f), nk = splat2(-1.f/9.f), neg = splat2(-1.f);
    const f2 mx = S1*k, my = T1*k;
    const f2 mxy = mx*my, mxx = mx*mx, myy = my*my;
    const f2 nsxx = fma2(S2, nk, mxx), nsyy = fma2(T2, nk, myy), nsxy = fma2(S3, nk, mxy);
    const f2 A1 = fma2(two, mxy, splat2(STV_C1)), A2 = fma2(splat2(-2.f), nsxy, splat2(STV_C2));
    const f2 B1 = mxx + myy + splat2(STV_C1), B2 = fma2(neg, nsxx + nsyy, splat2(STV_C2));
    const f2 num = A1*A2, den = B1*B2;
    const f2 iD = mk2(rcp_fast(lo2(den)), rcp_fast(hi2(den)));
    const f2 r = num*iD;
    e0 = __saturatef(fmaf(-0.5f, lo2(r), 0.5f));
    e1 = __saturatef(fmaf(-0.5f, hi2(r), 0.5f));
    const f2 kid = iD*k;
    const f2 dA = fma2(A1, neg, A2), dB = fma2(B1, neg, B2);
    const f2 a = kid*fma2(my*dA, neg, (r*mx)*dB);
    const f2 b = (kid*splat2(0.5f))*(r*B1);
    const f2 c = kid*(A1*neg);
    const bool v0 = fabsf(lo2(r)) <= 1.f, v1 = fabsf(hi2(r)) <= 1.f;  // un-clamped error inside [0,1]; NaN -> false
    ca = mk2(v0 ? lo2(a) : 0.f, v1 ? hi2(a) : 0.f);
    cb = mk2(v0 ? lo2(b) : 0.f, v1 ? hi2(b) : 0.f);
    cc = mk2(v0 ? lo2(c) : 0.f, v1 ? hi2(c) : 0.f);
}

// Photometric error of the thread's FRUN centres for one support frame (COEF: plus the SSIM coefficients per channel,
// coef[c*3 + q][pixel], un-scaled).
template <bool COEF>
__device__ __forceinline__ void photo_run_t(const TileT* sw, const TileT* st, int top, int col, const f2 (*T1)[FNP],
                                            const f2 (*T2)[FNP], float w_ssim, float w_l1, float* ek, float (*coef)[FRUN]) {
#pragma unroll
    for (int j = 0; j < FRUN; ++j) ek[j] = 0.f;
    const float ws = w_ssim*(1.f/3.f), wl = w_l1*(1.f/3.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        WinSums w;
        pair_sums_t(sw[c], st[c], top, col, w);
        if (w_ssim > 0.f) {
#pragma unroll
            for (int j = 0; j < FNP; ++j) {
                float e0, e1;
                if (COEF) {
                    f2 ca, cb, cc;
                    ssim_pair_coef(w.S1[j], w.S2[j], w.S3[j], T1[c][j], T2[c][j], e0, e1, ca, cb, cc);
                    coef[c*3 + 0][2*j] = lo2(ca); coef[c*3 + 0][2*j + 1] = hi2(ca);
                    coef[c*3 + 1][2*j] = lo2(cb); coef[c*3 + 1][2*j + 1] = hi2(cb);
                    coef[c*3 + 2][2*j] = lo2(cc); coef[c*3 + 2][2*j + 1] = hi2(cc);
                } else ssim_pair(w.S1[j], w.S2[j], w.S3[j], T1[c][j], T2[c][j], e0, e1);
                ek[2*j] = fmaf(ws, e0, ek[2*j]);
                ek[2*j + 1] = fmaf(ws, e1, ek[2*j + 1]);
            }
        }
        if (w_l1 > 0.f) {
#pragma unroll
            for (int j = 0; j < FNP; ++j) {  // centre 2j is the high lane of pair j, centre 2j+1 the low lane of pair j+1
                ek[2*j] = fmaf(wl, fabsf(hi2(w.wc[j]) - hi2(w.tc[j])), ek[2*j]);
                ek[2*j + 1] = fmaf(wl, fabsf(lo2(w.wc[j + 1]) - lo2(w.tc[j + 1])), ek[2*j + 1]);
            }
        }
    }
}

// Loads the halo-1 window of a 3-channel image into transposed tiles.
__device__ __forceinline__ void load_tile3_t(TileT* dst, const float* __restrict__ img, int y0, int x0, int H, int W) {
    const int HW = H*W;
    for (int q = threadIdx.x; q < FPH*FPW; q += FNT) {
        const int py = q/FPW, px = q - py*FPW;
        const int ya = clampi(reflect_idx(y0 + py, H), 0, H - 1), xa = clampi(reflect_idx(x0 + px, W), 0, W - 1);
        const float* p = img + ya*W + xa;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[c][px][py] = __ldg(p + c*HW);
    }
}

// Bilinear border sample of the 3 channels of support frame `plane0/3` at pixel-unit position (ix, iy).
//   TEX: one TLD4 (texture gather) per channel returns the 2x2 footprint with hardware address clamping; the footprint is
//        addressed at its texel centre (x0+1, y0+1) so the fixed-point coordinate conversion cannot move it.
//   !TEX: four read-only loads per channel (used when the frames do not meet the texture alignment rules).
template <bool TEX>
__device__ __forceinline__ void sample3(const PhotoParams& p, const float* __restrict__ sp, int plane0, float ix, float iy,
                                        float* out) {
    const int H = p.H, W = p.W;
    ix = fminf(fmaxf(ix, 0.f), (float)(W - 1));
    iy = fminf(fmaxf(iy, 0.f), (float)(H - 1));
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx = ix - x0f, wy = iy - y0f;
    if (TEX) {
        const float xc = x0f + 1.f, yc = y0f + 1.f + (float)(plane0*H);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 g = tex2Dgather<float4>(p.supp_tex, xc, yc + (float)(c*H), 0);  // x=(0,1) y=(1,1) z=(1,0) w=(0,0)
            const float top = fmaf(wx, g.z - g.w, g.w), bot = fmaf(wx, g.y - g.x, g.x);
            out[c] = fmaf(wy, bot - top, top);
        }
    } else {
        const int x0 = (int)x0f, y0 = (int)y0f, x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
        const int o00 = y0*W + x0, o01 = y0*W + x1, o10 = y1*W + x0, o11 = y1*W + x1, HW = H*W;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* q = sp + c*HW;
            const float a = __ldg(q + o00), b = __ldg(q + o01), cc = __ldg(q + o10), d = __ldg(q + o11);
            const float top = fmaf(wx, b - a, a), bot = fmaf(wx, d - cc, cc);
            out[c] = fmaf(wy, bot - top, top);
        }
    }
}

// Projection of pixel (u, v) with depth d to the sample position in the support frame (rows 9-11 of SURVEY 8a).
__device__ __forceinline__ void project_fast(const Cam& c, float u, float v, float d, float sx, float sy, float& ix, float& iy) {
    float P[3], Q[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) P[r] = fmaf(c.Ki[r*3], u, fmaf(c.Ki[r*3 + 1], v, c.Ki[r*3 + 2]))*d;
#pragma unroll
    for (int r = 0; r < 3; ++r) Q[r] = fmaf(c.R[r*3], P[0], fmaf(c.R[r*3 + 1], P[1], fmaf(c.R[r*3 + 2], P[2], c.t[r])));
    const float inv = rcp_fast(fmaxf(Q[2], STV_MIN_Z));  // max(max(z, eps), 0.1) == max(z, 0.1)
    const float nx = Q[0]*inv, ny = Q[1]*inv, nz = Q[2]*inv;
    ix = fmaf(fmaf(c.K0[0], nx, fmaf(c.K0[1], ny, c.K0[2]*nz)), sx, -0.5f);
    iy = fmaf(fmaf(c.K1[0], nx, fmaf(c.K1[1], ny, c.K1[2]*nz)), sy, -0.5f);
}

// ---------------------------------------------------------------------------------------------------------------------
// compute_photo on un-warped frames: identity (static) error for the auto-mask, and the stand-alone entry point.
// grid = (tiles, b)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FNT) photo_error_kernel(PhotoParams p, const float* __restrict__ pred,
                                                            float* __restrict__ err) {
    __shared__ __align__(16) float st[3][FPW][FPH];
    __shared__ __align__(16) float sw[3][FPW][FPH];
    const int tile = blockIdx.x, i = blockIdx.y;
    const int tx0 = (tile % p.tiles_x)*FTW, ty0 = (tile/p.tiles_x)*FTH;
    const int H = p.H, W = p.W, HW = H*W;
    load_tile3_t(st, p.tgt + (size_t)i*3*HW, ty0 - 1, tx0 - 1, H, W);
    __syncthreads();
    const int lx = threadIdx.x & (FTW - 1), top = (threadIdx.x/FTW)*FRUN;
    f2 T1[3][FNP], T2[3][FNP];
#pragma unroll
    for (int c = 0; c < 3; ++c) target_sums_t(st[c], top, lx, T1[c], T2[c]);

    float ered[FRUN];
#pragma unroll
    for (int j = 0; j < FRUN; ++j) ered[j] = p.use_min ? INFINITY : 0.f;
    for (int k = 0; k < p.n; ++k) {
        load_tile3_t(sw, pred + ((size_t)k*p.b + i)*3*HW, ty0 - 1, tx0 - 1, H, W);
        __syncthreads();
        float ek[FRUN];
        photo_run_t<false>(sw, st, top, lx, T1, T2, p.w_ssim, p.w_l1, ek, nullptr);
#pragma unroll
        for (int j = 0; j < FRUN; ++j) ered[j] = p.use_min ? fminf(ered[j], ek[j]) : ered[j] + ek[j];
        __syncthreads();
    }
    const int x = tx0 + lx;
#pragma unroll
    for (int j = 0; j < FRUN; ++j) {
        const int y = ty0 + top + j;
        if (y < H && x < W) err[(size_t)i*HW + y*W + x] = p.use_min ? ered[j] : ered[j]/(float)p.n;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward. grid = (tiles, b, S)
// ---------------------------------------------------------------------------------------------------------------------
#ifndef STV_FWD_MINB
#define STV_FWD_MINB 4  // measured on B200 at 8x384x640, n=2, S=4: (TH, NT, MINB) = (8, 256, 4) is the fastest of six variants
#endif
#ifndef STV_FWD_MINB_COEF
#define STV_FWD_MINB_COEF 4  // measured (config 3): 2 -> 0.62 ms, 3 -> 0.53, 4 -> 0.51: residency beats the 150 B of spills
#endif
template <bool TEX, bool COEF, int MINB = (COEF ? STV_FWD_MINB_COEF : STV_FWD_MINB)>
__global__ void __launch_bounds__(FNT, MINB) photo_fwd_kernel(PhotoParams p) {
    __shared__ __align__(16) float st[3][FPW][FPH];
    __shared__ __align__(16) float sw[3][FPW][FPH];
    __shared__ float red[32];
    const int tile = blockIdx.x, i = blockIdx.y, s = blockIdx.z;
    const int tx0 = (tile % p.tiles_x)*FTW, ty0 = (tile/p.tiles_x)*FTH;
    const int H = p.H, W = p.W, HW = H*W;
    const float sx = (float)W/(float)(W - 1), sy = (float)H/(float)(H - 1);

    // Per-thread halo positions (branch-free: surplus threads of the last round redo the final position). The mirrored
    // pixel and its depth do not depend on the support frame, so they are resolved once.
    const float* __restrict__ dp = p.depth[s] + (size_t)i*HW;
    const float* __restrict__ tg = p.tgt + (size_t)i*3*HW;
    float pu[NPOS], pv[NPOS], pd[NPOS];
    int so[NPOS];  // offset of the position inside one transposed tile
#pragma unroll
    for (int it = 0; it < NPOS; ++it) {
        const int q = min((int)threadIdx.x + it*FNT, FPH*FPW - 1);
        const int py = q/FPW, px = q - py*FPW;
        const int ya = clampi(reflect_idx(ty0 - 1 + py, H), 0, H - 1), xa = clampi(reflect_idx(tx0 - 1 + px, W), 0, W - 1);
        const int o = ya*W + xa;
        so[it] = px*FPH + py;
        pu[it] = (float)xa; pv[it] = (float)ya;
        pd[it] = __ldg(dp + o);
#pragma unroll
        for (int c = 0; c < 3; ++c) (&st[c][0][0])[so[it]] = __ldg(tg + c*HW + o);
    }
    __syncthreads();
    const int lx = threadIdx.x & (FTW - 1), top = (threadIdx.x/FTW)*FRUN;
    f2 T1[3][FNP], T2[3][FNP];
    if (p.w_ssim > 0.f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) target_sums_t(st[c], top, lx, T1[c], T2[c]);
    }

    float ered[FRUN];
    int ksel[FRUN];
#pragma unroll
    for (int j = 0; j < FRUN; ++j) { ered[j] = p.use_min ? INFINITY : 0.f; ksel[j] = p.use_min ? 0 : STV_SEL_MEAN; }

    // Identity error (+ tie-break noise) of the thread's centres: requested now, consumed after the support-frame loop.
    float e0v[FRUN];
    const uint64_t seed = p.seed ? p.seed + (p.step ? *p.step : 0ull) : 0ull;
    {
        const int x = min(tx0 + lx, W - 1);
#pragma unroll
        for (int j = 0; j < FRUN; ++j) {
            const int y = min(ty0 + top + j, H - 1);
            const size_t pix = (size_t)i*HW + y*W + x, nidx = (size_t)s*p.b*HW + pix;
            e0v[j] = 0.f;
            if (p.use_automask) {
                e0v[j] = __ldg(p.e0 + pix);
                if (p.noise) e0v[j] = fmaf(STV_EPS32, __ldg(p.noise + nidx), e0v[j]);
                else if (seed) e0v[j] = fmaf(STV_EPS32, hash_normal(seed, nidx), e0v[j]);
            }
        }
    }

    // COEF: the SSIM coefficients of a support frame are stored whenever it becomes the pixel's running minimum (a later
    // winner overwrites them in L2); pixels the static frame wins keep stale values that the backward masks out by `sel`.
    float* const coefp = COEF ? p.coef + ((size_t)s*p.b + i)*9*HW : nullptr;
    const bool want_warp = p.warp0 != nullptr && s == 0;
    for (int k = 0; k < p.n; ++k) {
        Cam cam;
        load_cam(cam, p.T + ((size_t)k*p.b + i)*16, p.K + (size_t)i*16, p.Kinv + (size_t)i*16);
        const int plane0 = (k*p.b + i)*3;
        const float* __restrict__ sp = p.supp + (size_t)plane0*HW;
        // phase 1: warp the tile + halo into shared memory. Sample positions first, then all gathers, then the lerps, so
        // that every thread keeps several texture requests in flight.
        float ix[NPOS], iy[NPOS];
#pragma unroll
        for (int it = 0; it < NPOS; ++it) project_fast(cam, pu[it], pv[it], pd[it], sx, sy, ix[it], iy[it]);
#pragma unroll
        for (int it = 0; it < NPOS; ++it) {
            float v[3];
            sample3<TEX>(p, sp, plane0, ix[it], iy[it], v);
#pragma unroll
            for (int c = 0; c < 3; ++c) (&sw[c][0][0])[so[it]] = v[c];
        }
        __syncthreads();
        if (want_warp) {  // logging path (scale 0 only): the block's interior of the warped frame goes to HBM
            const int x = tx0 + lx;
#pragma unroll
            for (int j = 0; j < FRUN; ++j) {
                const int y = ty0 + top + j;
                if (y < H && x < W) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) p.warp0[((size_t)plane0 + c)*HW + y*W + x] = sw[c][lx + 1][top + 1 + j];
                }
            }
        }
        // phase 2: photometric error, reduce over support frames (first index wins ties, as torch.min)
        float ek[FRUN];
        float cur[COEF ? 9 : 1][FRUN];
        if (COEF) {
#pragma unroll
            for (int q = 0; q < 9; ++q) {
#pragma unroll
                for (int j = 0; j < FRUN; ++j) cur[q][j] = 0.f;
            }
        }
        photo_run_t<COEF>(sw, st, top, lx, T1, T2, p.w_ssim, p.w_l1, ek, cur);
#pragma unroll
        for (int j = 0; j < FRUN; ++j) {
            if (p.use_min) {
                if (ek[j] < ered[j]) {
                    ered[j] = ek[j]; ksel[j] = k;
                    if (COEF) {
                        const int y = ty0 + top + j, x = tx0 + lx;
                        if (y < H && x < W) {
#pragma unroll
                            for (int q = 0; q < 9; ++q) coefp[(size_t)q*HW + y*W + x] = cur[q][j];
                        }
                    }
                }
            } else ered[j] += ek[j];
        }
        __syncthreads();
    }

    float acc = 0.f;
    const int x = tx0 + lx;
#pragma unroll
    for (int j = 0; j < FRUN; ++j) {
        const int y = ty0 + top + j;
        if (y < H && x < W) {
            float e = p.use_min ? ered[j] : ered[j]/(float)p.n;
            int sel = ksel[j];
            const size_t pix = (size_t)i*HW + y*W + x;
            if (p.use_automask && !(e <= e0v[j])) { e = e0v[j]; sel = STV_SEL_STATIC; }  // torch.min(cat(err, static)): index 0 wins ties
            p.sel[(size_t)s*p.b*HW + pix] = (uint8_t)sel;
            acc += e;
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) p.partial[((size_t)blockIdx.z*gridDim.y + blockIdx.y)*gridDim.x + blockIdx.x] = acc;
}

// loss = sum(partials) / count, accumulated in double in a fixed order. One block.
__global__ void reduce_mean_kernel(const float* __restrict__ partial, int n, double inv_count, float* __restrict__ out,
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
    if (threadIdx.x == 0) {
        *out = (float)(sh[0]*inv_count);
        if (step) *step += 1ull;  // every reader of this call's value ran before this kernel (stream order)
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward. grid = (tiles, b, S); dynamic shared memory.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef STV_BWD_MINB
#define STV_BWD_MINB 3
#endif
constexpr int N_ACC_T = 12, N_ACC_K = 15;  // dT (3x4) | dK rows 0-1 (2x3) + dKinv (3x3)

struct BwdSmem {
    float st[3][PH2][PW2];   // target, halo 2
    float sw[3][PH2][PW2];   // warped support, halo 2
    float sc[9][PH1][PW1];   // masked d err/d(S1,S2,S3) per channel at every centre of tile + halo 1
    uint8_t ssel[PH1][PW1];  // decisions at the centres
    float red[8][N_ACC_T + N_ACC_K];
};

template <bool NEED_K>
__global__ void __launch_bounds__(NT, 2) photo_bwd_kernel(PhotoParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);
    constexpr int NACC = N_ACC_T + (NEED_K ? N_ACC_K : 0);

    const int tile = blockIdx.x, i = blockIdx.y, s = blockIdx.z;
    const int tx0 = (tile % p.tiles_x)*TW, ty0 = (tile/p.tiles_x)*TH;
    const int H = p.H, W = p.W, HW = H*W;
    const float sx = (float)W/(float)(W - 1), sy = (float)H/(float)(H - 1);
    float g = __ldg(p.grad_loss)/((float)p.S*(float)p.b*(float)HW);
    if (!p.use_min) g /= (float)p.n;
    const float gs = g*p.w_ssim*(1.f/3.f), gl = g*p.w_l1*(1.f/3.f);

    load_tile3<PH2, PW2>(sm.st, p.tgt + (size_t)i*3*HW, ty0 - 2, tx0 - 2, H, W);
    const uint8_t* __restrict__ selp = p.sel_in + ((size_t)s*p.b + i)*HW;
    for (int q = threadIdx.x; q < PH1*PW1; q += NT) {
        const int py = q/PW1, px = q - py*PW1;
        const int y = ty0 - 1 + py, x = tx0 - 1 + px;
        sm.ssel[py][px] = (y >= 0 && y < H && x >= 0 && x < W) ? selp[y*W + x] : (uint8_t)STV_SEL_STATIC;
    }
    __syncthreads();

    // phase-2 mapping: vertical runs of 6 centres over the (TH+2) x (TW+2) centre grid
    constexpr int CRUN = 6;
    static_assert(PH1 % CRUN == 0 && (PH1/CRUN)*PW1 <= NT, "centre run mapping");
    const bool c_active = threadIdx.x < (PH1/CRUN)*PW1;
    const int ccol = threadIdx.x % PW1, ctop = (threadIdx.x/PW1)*CRUN;  // centre rows ctop.., smem(halo2) rows ctop..ctop+CRUN+1
    float T1[3][CRUN], T2[3][CRUN];
    if (c_active && p.w_ssim > 0.f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) target_sums<PH2, PW2, CRUN>(sm.st[c], ctop, ccol, T1[c], T2[c]);
    }

    // phase-3 mapping: vertical runs of RUN interior pixels
    const int lx = threadIdx.x & (TW - 1), top = (threadIdx.x/TW)*RUN;
    const int x = tx0 + lx;
    float gd[RUN];
#pragma unroll
    for (int j = 0; j < RUN; ++j) gd[j] = 0.f;
    const float* __restrict__ dp = p.depth[s] + (size_t)i*HW;

    for (int k = 0; k < p.n; ++k) {
        Cam cam;
        load_cam(cam, p.T + ((size_t)k*p.b + i)*16, p.K + (size_t)i*16, p.Kinv + (size_t)i*16);
        const float* __restrict__ sp = p.supp + ((size_t)k*p.b + i)*3*HW;
        // phase 1: warp tile + halo 2
        for (int q = threadIdx.x; q < PH2*PW2; q += NT) {
            const int py = q/PW2, px = q - py*PW2;
            const int ya = clampi(reflect_idx(ty0 - 2 + py, H), 0, H - 1), xa = clampi(reflect_idx(tx0 - 2 + px, W), 0, W - 1);
            const float d = __ldg(dp + ya*W + xa);
            Proj pr;
            project(cam, (float)xa, (float)ya, d, sx, sy, pr);
            Taps t;
            make_taps(pr.ix, pr.iy, H, W, t);
#pragma unroll
            for (int c = 0; c < 3; ++c) sm.sw[c][py][px] = sample_plane(sp + c*HW, t);
        }
        __syncthreads();
        // phase 2: SSIM coefficient planes at the centres that selected support k
        if (c_active) {
            bool on[CRUN];
#pragma unroll
            for (int j = 0; j < CRUN; ++j) {
                const uint8_t sv = sm.ssel[ctop + j][ccol];
                on[j] = p.w_ssim > 0.f && (sv == k || sv == STV_SEL_MEAN);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float S1[CRUN], S2[CRUN], S3[CRUN];
                if (p.w_ssim > 0.f) pair_sums<PW2, CRUN>(sm.sw[c], sm.st[c], ctop, ccol, S1, S2, S3);
#pragma unroll
                for (int j = 0; j < CRUN; ++j) {
                    float a = 0.f, bq = 0.f, cq = 0.f;
                    if (on[j]) {
                        ssim_err_grad(S1[j], S2[j], S3[j], T1[c][j], T2[c][j], a, bq, cq);
                        a *= gs; bq *= gs; cq *= gs;
                    }
                    sm.sc[c*3 + 0][ctop + j][ccol] = a;
                    sm.sc[c*3 + 1][ctop + j][ccol] = bq;
                    sm.sc[c*3 + 2][ctop + j][ccol] = cq;
                }
            }
        }
        __syncthreads();
        // phase 3: d loss / d warped pixel -> sampler -> projection
        float acc[NACC];
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[q] = 0.f;
        if (x < W) {
            const float mxl = (x == 1) ? 2.f : 1.f, mxr = (x == W - 2) ? 2.f : 1.f;
            float gw[3][RUN];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // horizontal (reflect-adjoint) sums of the three coefficient planes on rows top .. top+RUN+1 of the centre grid
                float ha[RUN + 2], hb[RUN + 2], hc[RUN + 2];
#pragma unroll
                for (int r = 0; r < RUN + 2; ++r) {
                    const float* ra = sm.sc[c*3 + 0][top + r] + lx;
                    const float* rb = sm.sc[c*3 + 1][top + r] + lx;
                    const float* rc = sm.sc[c*3 + 2][top + r] + lx;
                    ha[r] = fmaf(mxl, ra[0], fmaf(mxr, ra[2], ra[1]));
                    hb[r] = fmaf(mxl, rb[0], fmaf(mxr, rb[2], rb[1]));
                    hc[r] = fmaf(mxl, rc[0], fmaf(mxr, rc[2], rc[1]));
                }
#pragma unroll
                for (int j = 0; j < RUN; ++j) {
                    const int y = ty0 + top + j;
                    const float myu = (y == 1) ? 2.f : 1.f, myd = (y == H - 2) ? 2.f : 1.f;
                    const float A = fmaf(myu, ha[j], fmaf(myd, ha[j + 2], ha[j + 1]));
                    const float B = fmaf(myu, hb[j], fmaf(myd, hb[j + 2], hb[j + 1]));
                    const float C = fmaf(myu, hc[j], fmaf(myd, hc[j + 2], hc[j + 1]));
                    const float wv = sm.sw[c][top + j + 2][lx + 2], tv = sm.st[c][top + j + 2][lx + 2];
                    float gwv = fmaf(2.f*wv, B, fmaf(tv, C, A));
                    const uint8_t sv = sm.ssel[top + j + 1][lx + 1];
                    if (sv == k || sv == STV_SEL_MEAN) {
                        const float df = wv - tv;
                        gwv += df > 0.f ? gl : (df < 0.f ? -gl : 0.f);
                    }
                    gw[c][j] = gwv;
                }
            }
#pragma unroll
            for (int j = 0; j < RUN; ++j) {
                const int y = ty0 + top + j;
                if (y >= H) continue;
                if (gw[0][j] == 0.f && gw[1][j] == 0.f && gw[2][j] == 0.f) continue;
                const float d = __ldg(dp + y*W + x);
                Proj pr;
                project(cam, (float)x, (float)y, d, sx, sy, pr);
                Taps t;
                make_taps(pr.ix, pr.iy, H, W, t);
                float gix = 0.f, giy = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float dx, dy;
                    sample_plane_grad(sp + c*HW, t, dx, dy);
                    gix = fmaf(gw[c][j], dx, gix);
                    giy = fmaf(gw[c][j], dy, giy);
                }
                const float gqx = gix*t.gx*sx, gqy = giy*t.gy*sy;
                // q = K[:2,:3] nrm
                float gn[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) gn[r] = fmaf(cam.K0[r], gqx, cam.K1[r]*gqy);
                // nrm = Q/zc ; zc = max(Qz, 0.1)
                float gQ[3];
                float gz = 0.f;
#pragma unroll
                for (int r = 0; r < 3; ++r) { gQ[r] = gn[r]*pr.inv; gz = fmaf(gn[r], pr.Q[r], gz); }
                if (pr.Q[2] >= STV_MIN_Z) gQ[2] -= gz*pr.inv*pr.inv;
                // Q = R P + t ; P = d ray
                float gP[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) gP[r] = fmaf(cam.R[r], gQ[0], fmaf(cam.R[3 + r], gQ[1], cam.R[6 + r]*gQ[2]));
                gd[j] += fmaf(gP[0], pr.ray[0], fmaf(gP[1], pr.ray[1], gP[2]*pr.ray[2]));
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    acc[r*4 + 0] = fmaf(gQ[r], pr.P[0], acc[r*4 + 0]);
                    acc[r*4 + 1] = fmaf(gQ[r], pr.P[1], acc[r*4 + 1]);
                    acc[r*4 + 2] = fmaf(gQ[r], pr.P[2], acc[r*4 + 2]);
                    acc[r*4 + 3] += gQ[r];
                }
                if (NEED_K) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        acc[12 + r] = fmaf(gqx, pr.nrm[r], acc[12 + r]);
                        acc[15 + r] = fmaf(gqy, pr.nrm[r], acc[15 + r]);
                        const float gr = gP[r]*d;  // d loss / d ray_r
                        acc[18 + r*3 + 0] = fmaf(gr, (float)x, acc[18 + r*3 + 0]);
                        acc[18 + r*3 + 1] = fmaf(gr, (float)y, acc[18 + r*3 + 1]);
                        acc[18 + r*3 + 2] += gr;
                    }
                }
            }
        }
        // block reduction of the pose / intrinsics partials for this support frame
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            const float v = warp_sum(acc[q]);
            if (lane == 0) sm.red[wid][q] = v;
        }
        __syncthreads();
        if (threadIdx.x < NACC) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < NT/32; ++w) v += sm.red[w][threadIdx.x];
            const size_t blk = ((size_t)blockIdx.z*gridDim.y + blockIdx.y)*gridDim.x + blockIdx.x;
            p.partial[(blk*p.n + k)*(N_ACC_T + N_ACC_K) + threadIdx.x] = v;
        }
        // (the next iteration's phase 1 only writes sw, whose last readers finished before the barrier above)
    }
    if (x < W) {
#pragma unroll
        for (int j = 0; j < RUN; ++j) {
            const int y = ty0 + top + j;
            if (y < H) p.g_depth[s][(size_t)i*HW + y*W + x] = gd[j];
        }
    }
}

// Sums the per-block partials in a fixed order (double accumulation).
//   gT[k,i,r,c]   (r<3)   = sum_{s,tile} partial[((s*b+i)*tiles+tile)*n + k][r*4+c]
//   gK[i,r,c]     (r<2)   = sum_{s,tile,k} partial[...][12 + r*3 + c]
//   gKinv[i,r,c]  (r<3)   = sum_{s,tile,k} partial[...][18 + r*3 + c]
__global__ void __launch_bounds__(128) photo_bwd_finalize_kernel(const float* __restrict__ partial, int b, int n, int S, int tiles,
                                                                 float* __restrict__ gT, float* __restrict__ gK,
                                                                 float* __restrict__ gKinv) {
    // One warp per output element; lanes stride over the partial rows, then a fixed-shape shuffle tree (deterministic).
    constexpr int NA = N_ACC_T + N_ACC_K;
    const int nT = n*b*16, nK = b*16;
    const int idx = (blockIdx.x*blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (idx >= nT + 2*nK) return;
    float* out;
    int off, i, k0, k1, e;
    bool live;
    if (idx < nT) {
        const int c = idx & 3, r = (idx >> 2) & 3, ki = idx >> 4;
        k0 = ki/b; k1 = k0 + 1; i = ki - k0*b;
        out = gT; e = idx; off = r*4 + c; live = r < 3;
    } else {
        const int which = (idx - nT)/nK;
        e = (idx - nT) - which*nK;
        out = which == 0 ? gK : gKinv;
        const int c = e & 3, r = (e >> 2) & 3;
        i = e >> 4; k0 = 0; k1 = n;
        live = c < 3 && (which == 0 ? r < 2 : r < 3);
        off = which == 0 ? 12 + r*3 + c : 18 + r*3 + c;
    }
    if (out == nullptr) return;
    double a = 0.0;
    if (live) {
        const int rows = S*tiles;
        for (int q = lane; q < rows; q += 32) {
            const int s = q/tiles, t = q - s*tiles;
            for (int k = k0; k < k1; ++k) a += (double)partial[((((size_t)s*b + i)*tiles + t)*n + k)*NA + off];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) out[e] = (float)a;
}

}  // namespace stv

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
using namespace stv;

static int check_cfg(const stv_photo_cfg* c) {
    STV_REQUIRE(c != nullptr, "stv_photo: cfg is NULL");
    STV_REQUIRE(c->b > 0 && c->n > 0 && c->S > 0, "stv_photo: b, n, S must be positive (b=%d n=%d S=%d)", c->b, c->n, c->S);
    STV_REQUIRE(c->S <= STV_MAX_SCALES, "stv_photo: S=%d exceeds STV_MAX_SCALES=%d", c->S, STV_MAX_SCALES);
    STV_REQUIRE(c->n <= STV_MAX_SUPPORT, "stv_photo: n=%d exceeds STV_MAX_SUPPORT=%d", c->n, STV_MAX_SUPPORT);
    STV_REQUIRE(c->H >= 3 && c->W >= 3, "stv_photo: H, W must be >= 3 for reflection padding (H=%d W=%d)", c->H, c->W);
    STV_REQUIRE((long long)c->H*c->W*3 < (1ll << 31), "stv_photo: image too large for 32-bit plane offsets");
    STV_REQUIRE(c->w_ssim >= 0.f && c->w_l1 >= 0.f, "stv_photo: negative loss weights");
    STV_REQUIRE(c->b <= 65535 && c->S <= 65535, "stv_photo: batch too large for one launch");
    return STV_OK;
}

static void fill_params(PhotoParams& p, const stv_photo_cfg* c) {
    p.b = c->b; p.n = c->n; p.S = c->S; p.H = c->H; p.W = c->W;
    p.w_ssim = c->w_ssim; p.w_l1 = c->w_l1;
    p.use_min = c->use_min; p.use_automask = c->use_automask; p.seed = c->noise_seed;
    p.tiles_x = (c->W + TW - 1)/TW; p.tiles_y = (c->H + TH - 1)/TH;
}

static void use_fwd_tiles(PhotoParams& p) { p.tiles_x = (p.W + FTW - 1)/FTW; p.tiles_y = (p.H + FTH - 1)/FTH; }

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int stv::photo_identity_error(const stv_photo_cfg* c, const float* tgt, const float* supp, float* e0, cudaStream_t st) {
    PhotoParams p{};
    fill_params(p, c);
    use_fwd_tiles(p);
    p.tgt = tgt;
    photo_error_kernel<<<dim3(p.tiles_x*p.tiles_y, c->b), FNT, 0, st>>>(p, supp, e0);
    count_launch();
    return check_launch("photo_error_kernel");
}

extern "C" size_t stv_photo_workspace_bytes(const stv_photo_cfg* c) {
    if (check_cfg(c) != STV_OK) return 0;
    const size_t tiles_b = (size_t)((c->W + TW - 1)/TW)*((c->H + TH - 1)/TH);
    const size_t tiles_f = (size_t)((c->W + FTW - 1)/FTW)*((c->H + FTH - 1)/FTH);
    const size_t e0 = align256((size_t)c->b*c->H*c->W*sizeof(float));
    const size_t part_fwd = align256(tiles_f*c->b*c->S*sizeof(float));
    const size_t part_bwd = align256(tiles_b*c->b*c->S*c->n*(N_ACC_T + N_ACC_K)*sizeof(float));
    return e0 + (part_fwd > part_bwd ? part_fwd : part_bwd);
}

extern "C" int stv_photo_error(const stv_photo_cfg* c, const float* pred, const float* tgt, float* err, void* stream) {
    if (int rc = check_cfg(c)) return rc;
    STV_REQUIRE(pred && tgt && err, "stv_photo_error: NULL pointer");
    PhotoParams p{};
    fill_params(p, c);
    use_fwd_tiles(p);
    p.tgt = tgt;
    dim3 grid(p.tiles_x*p.tiles_y, c->b);
    photo_error_kernel<<<grid, FNT, 0, (cudaStream_t)stream>>>(p, pred, err);
    count_launch();
    return check_launch("photo_error_kernel");
}

extern "C" int stv_photo_fwd(const stv_photo_cfg* c, const float* const* depth, const float* tgt, const float* supp,
                             const float* T, const float* K, const float* Kinv, const float* noise,
                             unsigned long long* noise_step, float* loss, uint8_t* sel, float* warp0, void* ws,
                             size_t ws_bytes, void* stream) {
    if (int rc = check_cfg(c)) return rc;
    STV_REQUIRE(depth && tgt && supp && T && K && Kinv && loss && sel, "stv_photo_fwd: NULL pointer");
    for (int s = 0; s < c->S; ++s) STV_REQUIRE(depth[s] != nullptr, "stv_photo_fwd: depth[%d] is NULL", s);
    const size_t need = stv_photo_workspace_bytes(c);
    if (ws == nullptr || ws_bytes < need) {
        set_error("stv_photo_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return STV_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PhotoParams p{};
    fill_params(p, c);
    use_fwd_tiles(p);
    for (int s = 0; s < c->S; ++s) p.depth[s] = depth[s];
    p.tgt = tgt; p.supp = supp; p.T = T; p.K = K; p.Kinv = Kinv; p.noise = noise; p.step = noise_step;
    float* e0 = (float*)ws;
    p.e0 = e0;
    p.partial = (float*)((char*)ws + align256((size_t)c->b*c->H*c->W*sizeof(float)));
    p.sel = sel; p.warp0 = warp0;
    const int tiles = p.tiles_x*p.tiles_y;
    if (c->use_automask) {
        photo_error_kernel<<<dim3(tiles, c->b), FNT, 0, st>>>(p, supp, e0);
        count_launch();
        if (int rc = check_launch("photo_error_kernel")) return rc;
    }
    const dim3 grid(tiles, c->b, c->S);
    photo_fwd_kernel<false, false><<<grid, FNT, 0, st>>>(p);
    count_launch();
    if (int rc = check_launch("photo_fwd_kernel")) return rc;
    const int nblk = tiles*c->b*c->S;
    reduce_mean_kernel<<<1, 256, 0, st>>>(p.partial, nblk, 1.0/((double)c->S*c->b*c->H*c->W), loss,
                                          (c->use_automask && noise == nullptr && c->noise_seed) ? noise_step : nullptr);
    count_launch();
    return check_launch("reduce_mean_kernel");
}

extern "C" int stv_photo_bwd(const stv_photo_cfg* c, const float* const* depth, const float* tgt, const float* supp,
                             const float* T, const float* K, const float* Kinv, const uint8_t* sel,
                             const float* grad_loss, float* const* g_depth, float* gT, float* gK, float* gKinv, void* ws,
                             size_t ws_bytes, void* stream) {
    if (int rc = check_cfg(c)) return rc;
    STV_REQUIRE(depth && tgt && supp && T && K && Kinv && sel && grad_loss && g_depth && gT, "stv_photo_bwd: NULL pointer");
    for (int s = 0; s < c->S; ++s)
        STV_REQUIRE(depth[s] != nullptr && g_depth[s] != nullptr, "stv_photo_bwd: depth/g_depth[%d] is NULL", s);
    const size_t need = stv_photo_workspace_bytes(c);
    if (ws == nullptr || ws_bytes < need) {
        set_error("stv_photo_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return STV_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PhotoParams p{};
    fill_params(p, c);
    for (int s = 0; s < c->S; ++s) { p.depth[s] = depth[s]; p.g_depth[s] = g_depth[s]; }
    p.tgt = tgt; p.supp = supp; p.T = T; p.K = K; p.Kinv = Kinv; p.sel_in = sel; p.grad_loss = grad_loss;
    p.partial = (float*)((char*)ws + align256((size_t)c->b*c->H*c->W*sizeof(float)));
    const int tiles = p.tiles_x*p.tiles_y;
    const bool need_k = gK != nullptr || gKinv != nullptr;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(photo_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
        cudaFuncSetAttribute(photo_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem));
        attr_done = true;
    }
    dim3 grid(tiles, c->b, c->S);
    if (need_k) photo_bwd_kernel<true><<<grid, NT, sizeof(BwdSmem), st>>>(p);
    else photo_bwd_kernel<false><<<grid, NT, sizeof(BwdSmem), st>>>(p);
    count_launch();
    if (int rc = check_launch("photo_bwd_kernel")) return rc;
    if (!need_k) {
        // The finalize kernel reads the K slots of every partial row; keep them defined.
        // (photo_bwd_kernel<false> writes only the first N_ACC_T entries.)
    }
    const int total = c->n*c->b*16 + 2*c->b*16;
    photo_bwd_finalize_kernel<<<(total*32 + 127)/128, 128, 0, st>>>(p.partial, c->b, c->n, c->S, tiles, gT, need_k ? gK : nullptr,
                                                                 need_k ? gKinv : nullptr);
    count_launch();
    return check_launch("photo_bwd_finalize_kernel");
}
