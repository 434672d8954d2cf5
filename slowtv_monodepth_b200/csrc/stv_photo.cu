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
    float* partial;   // fwd: one float per block; bwd: n_acc floats per (block, k)
    uint8_t* sel;
    float* warp0;
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

// Photometric error of RUN centres (rows top+1.., column col+1 of the halo-1 tiles) for one support frame.
__device__ __forceinline__ void photo_run(const float (*sw)[PH1][PW1], const float (*st)[PH1][PW1], int top, int col,
                                          const float (*T1)[RUN], const float (*T2)[RUN], float w_ssim, float w_l1,
                                          float* ek) {
#pragma unroll
    for (int j = 0; j < RUN; ++j) ek[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (w_ssim > 0.f) {
            float S1[RUN], S2[RUN], S3[RUN];
            pair_sums<PW1, RUN>(sw[c], st[c], top, col, S1, S2, S3);
#pragma unroll
            for (int j = 0; j < RUN; ++j) ek[j] = fmaf(w_ssim*(1.f/3.f), ssim_err(S1[j], S2[j], S3[j], T1[c][j], T2[c][j]), ek[j]);
        }
        if (w_l1 > 0.f) {
#pragma unroll
            for (int j = 0; j < RUN; ++j)
                ek[j] = fmaf(w_l1*(1.f/3.f), fabsf(sw[c][top + 1 + j][col + 1] - st[c][top + 1 + j][col + 1]), ek[j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// compute_photo on un-warped frames: identity (static) error for the auto-mask, and the stand-alone entry point.
// grid = (tiles, b)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) photo_error_kernel(PhotoParams p, const float* __restrict__ pred,
                                                         float* __restrict__ err) {
    __shared__ float st[3][PH1][PW1];
    __shared__ float sw[3][PH1][PW1];
    const int tile = blockIdx.x, i = blockIdx.y;
    const int tx0 = (tile % p.tiles_x)*TW, ty0 = (tile/p.tiles_x)*TH;
    const int H = p.H, W = p.W, HW = H*W;
    load_tile3<PH1, PW1>(st, p.tgt + (size_t)i*3*HW, ty0 - 1, tx0 - 1, H, W);
    __syncthreads();
    const int lx = threadIdx.x & (TW - 1), top = (threadIdx.x/TW)*RUN;
    float T1[3][RUN], T2[3][RUN];
#pragma unroll
    for (int c = 0; c < 3; ++c) target_sums<PH1, PW1, RUN>(st[c], top, lx, T1[c], T2[c]);

    float ered[RUN];
#pragma unroll
    for (int j = 0; j < RUN; ++j) ered[j] = p.use_min ? INFINITY : 0.f;
    for (int k = 0; k < p.n; ++k) {
        load_tile3<PH1, PW1>(sw, pred + ((size_t)k*p.b + i)*3*HW, ty0 - 1, tx0 - 1, H, W);
        __syncthreads();
        float ek[RUN];
        photo_run(sw, st, top, lx, T1, T2, p.w_ssim, p.w_l1, ek);
#pragma unroll
        for (int j = 0; j < RUN; ++j) ered[j] = p.use_min ? fminf(ered[j], ek[j]) : ered[j] + ek[j];
        __syncthreads();
    }
    const int x = tx0 + lx;
#pragma unroll
    for (int j = 0; j < RUN; ++j) {
        const int y = ty0 + top + j;
        if (y < H && x < W) err[(size_t)i*HW + y*W + x] = p.use_min ? ered[j] : ered[j]/(float)p.n;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward. grid = (tiles, b, S)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) photo_fwd_kernel(PhotoParams p) {
    __shared__ float st[3][PH1][PW1];
    __shared__ float sw[3][PH1][PW1];
    __shared__ float red[32];
    const int tile = blockIdx.x, i = blockIdx.y, s = blockIdx.z;
    const int tx0 = (tile % p.tiles_x)*TW, ty0 = (tile/p.tiles_x)*TH;
    const int H = p.H, W = p.W, HW = H*W;
    const float sx = (float)W/(float)(W - 1), sy = (float)H/(float)(H - 1);

    load_tile3<PH1, PW1>(st, p.tgt + (size_t)i*3*HW, ty0 - 1, tx0 - 1, H, W);
    __syncthreads();
    const int lx = threadIdx.x & (TW - 1), top = (threadIdx.x/TW)*RUN;
    float T1[3][RUN], T2[3][RUN];
    if (p.w_ssim > 0.f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) target_sums<PH1, PW1, RUN>(st[c], top, lx, T1[c], T2[c]);
    }

    float ered[RUN];
    int ksel[RUN];
#pragma unroll
    for (int j = 0; j < RUN; ++j) { ered[j] = p.use_min ? INFINITY : 0.f; ksel[j] = p.use_min ? 0 : STV_SEL_MEAN; }

    const float* __restrict__ dp = p.depth[s] + (size_t)i*HW;
    const bool want_warp = p.warp0 != nullptr && s == 0;
    for (int k = 0; k < p.n; ++k) {
        Cam cam;
        load_cam(cam, p.T + ((size_t)k*p.b + i)*16, p.K + (size_t)i*16, p.Kinv + (size_t)i*16);
        const float* __restrict__ sp = p.supp + ((size_t)k*p.b + i)*3*HW;
        // phase 1: warp the tile + halo into shared memory
        for (int q = threadIdx.x; q < PH1*PW1; q += NT) {
            const int py = q/PW1, px = q - py*PW1;
            const int yy = ty0 - 1 + py, xx = tx0 - 1 + px;
            const int ya = clampi(reflect_idx(yy, H), 0, H - 1), xa = clampi(reflect_idx(xx, W), 0, W - 1);
            const float d = __ldg(dp + ya*W + xa);
            Proj pr;
            project(cam, (float)xa, (float)ya, d, sx, sy, pr);
            Taps t;
            make_taps(pr.ix, pr.iy, H, W, t);
            const bool own = want_warp && py >= 1 && py <= TH && px >= 1 && px <= TW && yy < H && xx < W;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = sample_plane(sp + c*HW, t);
                sw[c][py][px] = v;
                if (own) p.warp0[(((size_t)k*p.b + i)*3 + c)*HW + yy*W + xx] = v;
            }
        }
        __syncthreads();
        // phase 2: photometric error, reduce over support frames (first index wins ties, as torch.min)
        float ek[RUN];
        photo_run(sw, st, top, lx, T1, T2, p.w_ssim, p.w_l1, ek);
#pragma unroll
        for (int j = 0; j < RUN; ++j) {
            if (p.use_min) { if (ek[j] < ered[j]) { ered[j] = ek[j]; ksel[j] = k; } }
            else ered[j] += ek[j];
        }
        __syncthreads();
    }

    float acc = 0.f;
    const int x = tx0 + lx;
#pragma unroll
    for (int j = 0; j < RUN; ++j) {
        const int y = ty0 + top + j;
        if (y < H && x < W) {
            float e = p.use_min ? ered[j] : ered[j]/(float)p.n;
            int sel = ksel[j];
            const size_t pix = (size_t)i*HW + y*W + x;
            if (p.use_automask) {
                float e0 = __ldg(p.e0 + pix);
                const size_t nidx = (size_t)s*p.b*HW + pix;
                if (p.noise) e0 = fmaf(STV_EPS32, __ldg(p.noise + nidx), e0);
                else if (p.seed) e0 = fmaf(STV_EPS32, hash_normal(p.seed, nidx), e0);
                if (!(e <= e0)) { e = e0; sel = STV_SEL_STATIC; }  // torch.min(cat(err, static)): index 0 wins ties
            }
            p.sel[(size_t)s*p.b*HW + pix] = (uint8_t)sel;
            acc += e;
        }
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) p.partial[((size_t)blockIdx.z*gridDim.y + blockIdx.y)*gridDim.x + blockIdx.x] = acc;
}

// loss = sum(partials) / count, accumulated in double in a fixed order. One block.
__global__ void reduce_mean_kernel(const float* __restrict__ partial, int n, double inv_count, float* __restrict__ out) {
    __shared__ double sh[256];
    double a = 0.0;
    for (int q = threadIdx.x; q < n; q += blockDim.x) a += (double)partial[q];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x/2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = (float)(sh[0]*inv_count);
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward. grid = (tiles, b, S); dynamic shared memory.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int N_ACC_T = 12, N_ACC_K = 15;  // dT (3x4) | dK rows 0-1 (2x3) + dKinv (3x3)

struct BwdSmem {
    float st[3][PH2][PW2];   // target, halo 2
    float sw[3][PH2][PW2];   // warped support, halo 2
    float sc[9][PH1][PW1];   // masked d err/d(S1,S2,S3) per channel at every centre of tile + halo 1
    uint8_t ssel[PH1][PW1];  // decisions at the centres
    float red[8][N_ACC_T + N_ACC_K];
};

template <bool NEED_K>
__global__ void __launch_bounds__(NT) photo_bwd_kernel(PhotoParams p) {
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
__global__ void photo_bwd_finalize_kernel(const float* __restrict__ partial, int b, int n, int S, int tiles,
                                          float* __restrict__ gT, float* __restrict__ gK, float* __restrict__ gKinv) {
    constexpr int NA = N_ACC_T + N_ACC_K;
    const int nT = n*b*16, nK = b*16;
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx < nT) {
        const int c = idx & 3, r = (idx >> 2) & 3, ki = idx >> 4, k = ki/b, i = ki - k*b;
        double a = 0.0;
        if (r < 3)
            for (int s = 0; s < S; ++s)
                for (int t = 0; t < tiles; ++t)
                    a += (double)partial[((((size_t)s*b + i)*tiles + t)*n + k)*NA + r*4 + c];
        gT[idx] = (float)a;
    } else if (idx < nT + 2*nK) {
        const int which = (idx - nT)/nK, e = (idx - nT) - which*nK;
        float* out = which == 0 ? gK : gKinv;
        if (out == nullptr) return;
        const int c = e & 3, r = (e >> 2) & 3, i = e >> 4;
        double a = 0.0;
        const bool live = c < 3 && (which == 0 ? r < 2 : r < 3);
        if (live) {
            const int off = which == 0 ? 12 + r*3 + c : 18 + r*3 + c;
            for (int s = 0; s < S; ++s)
                for (int t = 0; t < tiles; ++t)
                    for (int k = 0; k < n; ++k)
                        a += (double)partial[((((size_t)s*b + i)*tiles + t)*n + k)*NA + off];
        }
        out[e] = (float)a;
    }
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

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t stv_photo_workspace_bytes(const stv_photo_cfg* c) {
    if (check_cfg(c) != STV_OK) return 0;
    const size_t tiles = (size_t)((c->W + TW - 1)/TW)*((c->H + TH - 1)/TH);
    const size_t nblk = tiles*c->b*c->S;
    const size_t e0 = align256((size_t)c->b*c->H*c->W*sizeof(float));
    const size_t part_fwd = align256(nblk*sizeof(float));
    const size_t part_bwd = align256(nblk*c->n*(N_ACC_T + N_ACC_K)*sizeof(float));
    return e0 + (part_fwd > part_bwd ? part_fwd : part_bwd);
}

extern "C" int stv_photo_error(const stv_photo_cfg* c, const float* pred, const float* tgt, float* err, void* stream) {
    if (int rc = check_cfg(c)) return rc;
    STV_REQUIRE(pred && tgt && err, "stv_photo_error: NULL pointer");
    PhotoParams p{};
    fill_params(p, c);
    p.tgt = tgt;
    dim3 grid(p.tiles_x*p.tiles_y, c->b);
    photo_error_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(p, pred, err);
    count_launch();
    return check_launch("photo_error_kernel");
}

extern "C" int stv_photo_fwd(const stv_photo_cfg* c, const float* const* depth, const float* tgt, const float* supp,
                             const float* T, const float* K, const float* Kinv, const float* noise, float* loss,
                             uint8_t* sel, float* warp0, void* ws, size_t ws_bytes, void* stream) {
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
    for (int s = 0; s < c->S; ++s) p.depth[s] = depth[s];
    p.tgt = tgt; p.supp = supp; p.T = T; p.K = K; p.Kinv = Kinv; p.noise = noise;
    float* e0 = (float*)ws;
    p.e0 = e0;
    p.partial = (float*)((char*)ws + align256((size_t)c->b*c->H*c->W*sizeof(float)));
    p.sel = sel; p.warp0 = warp0;
    const int tiles = p.tiles_x*p.tiles_y;
    if (c->use_automask) {
        photo_error_kernel<<<dim3(tiles, c->b), NT, 0, st>>>(p, supp, e0);
        count_launch();
        if (int rc = check_launch("photo_error_kernel")) return rc;
    }
    photo_fwd_kernel<<<dim3(tiles, c->b, c->S), NT, 0, st>>>(p);
    count_launch();
    if (int rc = check_launch("photo_fwd_kernel")) return rc;
    const int nblk = tiles*c->b*c->S;
    reduce_mean_kernel<<<1, 256, 0, st>>>(p.partial, nblk, 1.0/((double)c->S*c->b*c->H*c->W), loss);
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
    photo_bwd_finalize_kernel<<<(total + 127)/128, 128, 0, st>>>(p.partial, c->b, c->n, c->S, tiles, gT, need_k ? gK : nullptr,
                                                                 need_k ? gKinv : nullptr);
    count_launch();
    return check_launch("photo_bwd_finalize_kernel");
}
