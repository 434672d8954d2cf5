// Single-pass fused view-synthesis photometric loss for sm_100a: loss AND its gradients in ONE sweep over the batch.
//
// Why one pass. The two-pass formulation (stv_photo_fwd + stv_photo_bwd) either re-warps a halo-2 tile in the backward or
// hands nine SSIM coefficient planes per (scale, pixel) through HBM (2x the algorithmic traffic). Here nothing intermediate
// touches HBM: the loss is a mean, so d loss/d(.) is known up to the scalar dL/dloss the moment a pixel's decision is taken.
// The kernel emits UNIT gradients (dL/dloss = 1); the backward entry point only scales them (and, when the kernel was fed the
// network's low-resolution disparities, applies the adjoint of the bilinear up-sampling).
//
// Work decomposition: one WARP owns a strip of 28 columns x `rows` rows of one (image, scale) and sweeps it top to bottom, one
// image row per iteration, lane = column (28 interior lanes + 2 halo lanes on each side). Everything horizontal is a warp
// shuffle, everything vertical lives in registers as a 3-row running sum, so there is no shared-memory tile, no block barrier
// and no halo re-computation in y beyond the 4 warm-up rows of a strip:
//   stage A (row y)     up-sampled depth -> back-project -> rigid transform -> project -> 2x2 gather (TLD4) of each support frame
//                       -> warped RGB + its d/d(ix,iy); horizontal 3-sums of w, w^2, w*t (shuffles) pushed into the vertical sums
//   stage B (row y-1)   3x3 SSIM + L1 per support frame, min-reprojection, auto-mask -> decision byte, loss partial; the SSIM
//                       coefficients d e/d(S1,S2,S3) of the SELECTED frame, masked per frame, horizontally summed with the
//                       reflection multiplicities (shuffles) and pushed into a second set of vertical sums
//   stage C (row y-2)   d loss/d warped pixel = A + 2 w B + t C + L1 sign -> bilinear sampler -> projection -> d/d depth (stored),
//                       d/dT, d/dK (per-lane accumulators, reduced once per strip; d/dKinv follows from the d/dT moments)
// Values a later stage needs from an earlier row (warped RGB, sampler derivatives, depth, target) wait in a 3-slot ring in
// shared memory, lane-private (no bank conflicts, __syncwarp only).
#include "stv_common.cuh"
#include "stv_f2.cuh"

namespace stv {

constexpr int FZ_COLS = 28;      // interior columns per warp (32 lanes - 2 x halo 2)
constexpr int FZ_WARPS = 4;      // independent warps per block
constexpr int FZ_NT = FZ_WARPS*32;
constexpr int FZ_NPART_K = 6;    // d/dK rows 0-1 accumulators (shared by the support frames)

template <int N> struct FzLayout {
    static constexpr int CAM = 12 + 4 + 4 + N*12;      // float4 rows: Kinv rows 0-2 (w = 0), K row 0, K row 1, then per frame (R[r][0..2], t[r]), r = 0..2
    static constexpr int CAM_PAD = (CAM + 31)/32*32;
    static constexpr int RV = 5 + 9*N;                 // ring values per lane: depth, target RGB, d depth/d src, per frame w[3], gx[3], gy[3]
    static constexpr int PER_WARP = CAM_PAD + 3*RV*32; // floats
    static constexpr int NPART = N*12 + FZ_NPART_K;    // per-strip pose / intrinsics partial sums
};

struct FusedParams {
    int b, n, S, H, W;
    float w_ssim, w_l1;
    int use_automask;
    uint64_t seed;
    const unsigned long long* step;
    int mode;                 // 0: src = depth maps (b,1,H,W); 1: src = sigmoid disparities (b,1,h,w), up-sampled + scaled here
    int h[STV_MAX_SCALES], w[STV_MAX_SCALES];
    float d_mul, d_add;       // to_scaled: disp' = d_mul*disp + d_add (geometry.py:73-75)
    int scaled;
    int rows, nsx, nsy;       // strip height, strips per image in x / y
    const float* src[STV_MAX_SCALES];
    float* g_unit[STV_MAX_SCALES];   // d loss/d src_up at full resolution (mode 0: d/d depth, mode 1: d/d disp_up), dL/dloss = 1
    const float *tgt, *supp, *T, *K, *Kinv, *noise, *e0;
    unsigned long long tex;   // cudaTextureObject_t over supp viewed as (n*b*3*H) x W, or 0
    float* loss_partial;      // [strip]
    float* gpart;             // [strip][NPART]
    uint8_t* sel;
    float* warp0;
};

__device__ __forceinline__ int fz_clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// SSIM error + d err/d(S1,S2,S3) for two support frames at once (packed lanes = frames); zero coefficients where the clamp to
// [0,1] is active. Same algebra as the two-pass kernels (src/losses/photometric.py:40-50):
//   r = A1 A2/(B1 B2);  a = k (r mx (B2-B1) - my (A2-A1))/(B1 B2);  b = k r/(2 B2);  c = -k A1/(B1 B2),  k = 1/9.
__device__ __forceinline__ void fz_ssim2(f2 S1, f2 S2, f2 S3, float T1s, float T2s, f2& e, f2& ca, f2& cb, f2& cc) {
    const f2 k = splat2(1.f/9.f), two = splat2(2.f), nk = splat2(-1.f/9.f), neg = splat2(-1.f);
    const f2 T1 = splat2(T1s), T2 = splat2(T2s);
    const f2 mx = S1*k, my = T1*k;
    const f2 mxy = mx*my, mxx = mx*mx, myy = my*my;
    const f2 nsxx = fma2(S2, nk, mxx), nsyy = fma2(T2, nk, myy), nsxy = fma2(S3, nk, mxy);
    const f2 A1 = fma2(two, mxy, splat2(STV_C1)), A2 = fma2(splat2(-2.f), nsxy, splat2(STV_C2));
    const f2 B1 = mxx + myy + splat2(STV_C1), B2 = fma2(neg, nsxx + nsyy, splat2(STV_C2));
    const f2 num = A1*A2, den = B1*B2;
    const f2 iD = mk2(rcp_fast(lo2(den)), rcp_fast(hi2(den)));
    const f2 r = num*iD;
    e = mk2(__saturatef(fmaf(-0.5f, lo2(r), 0.5f)), __saturatef(fmaf(-0.5f, hi2(r), 0.5f)));
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

#ifndef STV_FUSED_MINB
#define STV_FUSED_MINB 1
#endif
template <int N, bool TEX, bool GRAD>
__global__ void __launch_bounds__(FZ_NT, STV_FUSED_MINB) photo_fused_kernel(const FusedParams p) {
    using LY = FzLayout<N>;
    constexpr int NP = (N + 1)/2;  // packed pairs of support frames
    extern __shared__ __align__(16) float fz_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* const cam = fz_smem + wid*LY::PER_WARP;
    float* const ring = cam + LY::CAM_PAD + lane;

    const long long total = (long long)p.b*p.S*p.nsx*p.nsy;
    const long long strip = (long long)blockIdx.x*FZ_WARPS + wid;
    if (strip >= total) return;  // warps are independent: no block-wide barrier below
    int t = (int)(strip % ((long long)p.nsx*p.nsy*p.S));
    const int i = (int)(strip/((long long)p.nsx*p.nsy*p.S));
    const int sxi = t % p.nsx; t /= p.nsx;
    const int syi = t % p.nsy;
    const int s = t/p.nsy;

    const int H = p.H, W = p.W, HW = H*W;
    const int x0 = sxi*FZ_COLS, y0 = syi*p.rows;
    const int x = x0 - 2 + lane;
    const int xa = fz_clampi(reflect_idx(x, W), 0, W - 1);
    const bool col_centre = lane >= 1 && lane <= 30 && x >= 0 && x < W;
    const bool col_inner = lane >= 2 && lane <= 29 && x < W;
    const float mxl = (x == 1) ? 2.f : 1.f, mxr = (x == W - 2) ? 2.f : 1.f;
    const float sx = (float)W/(float)(W - 1), sy = (float)H/(float)(H - 1);
    const float u = (float)xa;

    // camera constants of (i, frames) -> this warp's shared memory (read back as broadcasts)
    {
        const float* __restrict__ Km = p.K + (size_t)i*16;
        const float* __restrict__ Ki = p.Kinv + (size_t)i*16;
        for (int q = lane; q < LY::CAM; q += 32) {
            float v;
            if (q < 12) v = (q & 3) < 3 ? __ldg(Ki + q) : 0.f;       // rows 0-2 of the 4x4 inverse intrinsics
            else if (q < 20) v = __ldg(Km + (q - 12));               // rows 0-1 of the intrinsics
            else {
                const int k = (q - 20)/12, e = (q - 20) % 12;        // rows 0-2 of T_k: (R | t)
                v = __ldg(p.T + ((size_t)k*p.b + i)*16 + e);
            }
            cam[q] = v;
        }
    }
    __syncwarp();
    // read back as 128-bit broadcasts (one LDS per matrix row instead of one per element: the constants do not fit in registers)
    const float4* const cam4 = reinterpret_cast<const float4*>(cam);

    // low-resolution taps of this lane's column (mode 1)
    const int hs = p.h[s], ws = p.w[s];
    const bool resize = p.mode == 1 && (hs != H || ws != W);
    int lx0 = 0, lx1 = 0;
    float llx = 0.f;
    const float ry = (float)hs/(float)H;
    if (resize) {
        const float src = fmaxf(((float)ws/(float)W)*((float)xa + 0.5f) - 0.5f, 0.f);
        lx0 = min((int)src, ws - 1);
        lx1 = lx0 + (lx0 < ws - 1 ? 1 : 0);
        llx = src - (float)lx0;
    }
    const float* __restrict__ srcp = p.src[s] + (size_t)i*(p.mode == 1 ? hs*ws : HW);
    const float* __restrict__ tg = p.tgt + (size_t)i*3*HW;
    float* __restrict__ gup = GRAD ? p.g_unit[s] + (size_t)i*HW : nullptr;
    uint8_t* __restrict__ selp = p.sel + ((size_t)s*p.b + i)*HW;
    const float* __restrict__ e0p = p.e0 + (size_t)i*HW;
    const uint64_t seed = p.seed ? p.seed + (p.step ? *p.step : 0ull) : 0ull;
    const bool want_warp = p.warp0 != nullptr && s == 0;

    const float inv_cnt = 1.f/((float)p.S*(float)p.b*(float)HW);
    const float ws3 = p.w_ssim*(1.f/3.f), wl3 = p.w_l1*(1.f/3.f);
    const float gs = ws3*inv_cnt, gl = wl3*inv_cnt;   // unit-gradient weights of the SSIM / L1 terms of one pixel

    // vertical 3-row running sums: q1 = previous row, q2 = previous two rows
    f2 hS1[NP][3][2], hS2[NP][3][2], hS3[NP][3][2];   // per frame pair, channel: horizontal sums of w, w^2, w*t
    float hT1[3][2], hT2[3][2];                       // target
    f2 vc[NP][9][2];                                  // masked coefficient sums, two frames per packed pair
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        hT1[c][0] = hT1[c][1] = hT2[c][0] = hT2[c][1] = 0.f;
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            hS1[q][c][0] = hS1[q][c][1] = hS2[q][c][0] = hS2[q][c][1] = hS3[q][c][0] = hS3[q][c][1] = splat2(0.f);
        }
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
#pragma unroll
        for (int j = 0; j < 9; ++j) vc[q][j][0] = vc[q][j][1] = splat2(0.f);
    }
    float accA[N][3], accB[N][3], accC[N][3], accK[FZ_NPART_K];
#pragma unroll
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int r = 0; r < 3; ++r) accA[k][r] = accB[k][r] = accC[k][r] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < FZ_NPART_K; ++q) accK[q] = 0.f;
    float loss_acc = 0.f;
    int kcode_prev = 255;      // decision of this lane's centre one row up (needed by stage C for the L1 term)

    const int rows_here = min(p.rows, H - y0);
    const int n_it = rows_here + 4;
    int slot_w = 0;            // ring slot written by this iteration (it % 3)

    // Loads of an iteration are requested one iteration ahead (software pipelining: a warp shares its scheduler with one other
    // warp only, so nothing else hides a dependent chain of three L2 round trips per row): target RGB, the source-map taps and
    // the identity error of row `it` sit in registers when the iteration starts.
    struct RowLoads { float t[3], a00, a01, a10, a11, lly, e0; };
    auto request = [&](int it_n, RowLoads& L) {
        const int yn = y0 - 2 + it_n;
        const int yan = fz_clampi(reflect_idx(yn, H), 0, H - 1);
        const int on = yan*W + xa;
#pragma unroll
        for (int c = 0; c < 3; ++c) L.t[c] = __ldg(tg + c*HW + on);
        if (resize) {
            const float src = fmaxf(ry*((float)yan + 0.5f) - 0.5f, 0.f);
            const int ly0 = min((int)src, hs - 1), ly1 = ly0 + (ly0 < hs - 1 ? 1 : 0);
            L.lly = src - (float)ly0;
            const float* r0 = srcp + ly0*ws; const float* r1 = srcp + ly1*ws;
            L.a00 = __ldg(r0 + lx0); L.a01 = __ldg(r0 + lx1); L.a10 = __ldg(r1 + lx0); L.a11 = __ldg(r1 + lx1);
        } else { L.a00 = __ldg(srcp + on); L.a01 = L.a10 = L.a11 = 0.f; L.lly = 0.f; }
        // identity error of the centre row this iteration will decide (yn - 1)
        const int ycn = yn - 1;
        L.e0 = (p.use_automask && col_centre && ycn >= 0 && ycn < H) ? __ldg(e0p + (size_t)ycn*W + x) : 0.f;
    };
    RowLoads cur;
    request(0, cur);
#pragma unroll 2
    for (int it = 0; it < n_it; ++it) {
        const int y = y0 - 2 + it;
        const int ya = fz_clampi(reflect_idx(y, H), 0, H - 1);
        const float v = (float)ya;
        float* const rw = ring + slot_w*LY::RV*32;
        const int slot_c = slot_w == 0 ? 2 : slot_w - 1;            // row y-1
        const int slot_p = slot_c == 0 ? 2 : slot_c - 1;            // row y-2
        const float* const rc = ring + slot_c*LY::RV*32;
        const float* const rp = ring + slot_p*LY::RV*32;

        // ---------------- stage A: row y ----------------
        float tv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) tv[c] = cur.t[c];
        const float e0_row = cur.e0;
        float d, dchain = 1.f;   // depth and d depth/d src_up
        {
            const float sv = resize ? (1.f - cur.lly)*((1.f - llx)*cur.a00 + llx*cur.a01) + cur.lly*((1.f - llx)*cur.a10 + llx*cur.a11) : cur.a00;
            if (p.mode == 1) {
                const float dp = p.scaled ? __fadd_rn(__fmul_rn(p.d_mul, sv), p.d_add) : sv;
                d = dp > 0.f ? 1.0f/fmaxf(dp, STV_EPS32) : 0.f;
                dchain = (dp > 0.f && dp >= STV_EPS32) ? -(d*d)*(p.scaled ? p.d_mul : 1.f) : 0.f;
            } else d = sv;
        }
        if (it + 1 < n_it) request(it + 1, cur);   // next row's loads fly while this row is processed
        float ray[3], P[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) { const float4 ki = cam4[r]; ray[r] = fmaf(ki.x, u, fmaf(ki.y, v, ki.z)); P[r] = ray[r]*d; }
        rw[0] = d;
#pragma unroll
        for (int c = 0; c < 3; ++c) rw[(1 + c)*32] = tv[c];
        if (GRAD) rw[4*32] = dchain;

        float wv[N][3];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            float Q[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) { const float4 rt = cam4[5 + k*3 + r]; Q[r] = fmaf(rt.x, P[0], fmaf(rt.y, P[1], fmaf(rt.z, P[2], rt.w))); }
            const float inv = rcp_fast(fmaxf(Q[2], STV_MIN_Z));  // max(max(z, eps), 0.1) == max(z, 0.1)
            const float nx = Q[0]*inv, ny = Q[1]*inv, nz = Q[2]*inv;
            const float4 k0 = cam4[3], k1 = cam4[4];
            float ix = fmaf(fmaf(k0.x, nx, fmaf(k0.y, ny, k0.z*nz)), sx, -0.5f);
            float iy = fmaf(fmaf(k1.x, nx, fmaf(k1.y, ny, k1.z*nz)), sy, -0.5f);
            const float mxc = (float)(W - 1), myc = (float)(H - 1);
            const float bx = (ix > 0.f && ix < mxc) ? sx : 0.f, by = (iy > 0.f && iy < myc) ? sy : 0.f;  // d(clamped)/d(raw) * d ix/d qx
            ix = fminf(fmaxf(ix, 0.f), mxc);
            iy = fminf(fmaxf(iy, 0.f), myc);
            const float x0f = floorf(ix), y0f = floorf(iy);
            const float fx = ix - x0f, fy = iy - y0f;
            const int plane0 = (k*p.b + i)*3;
            float ta[3], tb[3], tc[3], td[3];  // taps (x0,y0) (x1,y0) (x0,y1) (x1,y1)
            {   // (measured: 843-852 us with, 855-870 us without this prefetch for the forward entry point at config 3)
                // the next sweep row gathers (approximately) one image row further down: warm L1 with it while this row is processed
                const int xi0 = (int)x0f, yi2 = min((int)y0f + 2, H - 1);
                const float* __restrict__ q = p.supp + (size_t)plane0*HW + yi2*W + xi0;
#pragma unroll
                for (int c = 0; c < 3; ++c) asm volatile("prefetch.global.L1 [%0];" :: "l"(q + c*HW));
            }
            if (TEX) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 t4 = tex2Dgather<float4>((cudaTextureObject_t)p.tex, x0f + 1.f, y0f + 1.f + (float)((plane0 + c)*H), 0);
                    ta[c] = t4.w; tb[c] = t4.z; tc[c] = t4.x; td[c] = t4.y;
                }
            } else {
                const int xi0 = (int)x0f, yi0 = (int)y0f, xi1 = min(xi0 + 1, W - 1), yi1 = min(yi0 + 1, H - 1);
                const float* __restrict__ sp = p.supp + (size_t)plane0*HW;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float* q = sp + c*HW;
                    ta[c] = __ldg(q + yi0*W + xi0); tb[c] = __ldg(q + yi0*W + xi1); tc[c] = __ldg(q + yi1*W + xi0); td[c] = __ldg(q + yi1*W + xi1);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float top = fmaf(fx, tb[c] - ta[c], ta[c]), bot = fmaf(fx, td[c] - tc[c], tc[c]);
                const float w = fmaf(fy, bot - top, top);
                wv[k][c] = w;
                rw[(5 + k*9 + c)*32] = w;
                if (GRAD) {
                    rw[(5 + k*9 + 3 + c)*32] = fmaf(fy, (td[c] - tc[c]) - (tb[c] - ta[c]), tb[c] - ta[c])*bx;
                    rw[(5 + k*9 + 6 + c)*32] = (bot - top)*by;
                }
            }
            if (want_warp && col_inner && it >= 2 && it < rows_here + 2) {
#pragma unroll
                for (int c = 0; c < 3; ++c) p.warp0[((size_t)plane0 + c)*HW + ya*W + xa] = wv[k][c];
            }
        }

        // horizontal 3-sums of this row (neighbour columns by shuffle) and the vertical running sums -> window sums of row y-1
        float T1[3], T2[3];
        f2 S1[NP][3], S2[NP][3], S3[NP][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float tl = __shfl_up_sync(0xffffffffu, tv[c], 1), tr = __shfl_down_sync(0xffffffffu, tv[c], 1);
            const float h1 = tl + tv[c] + tr;
            const float h2 = fmaf(tl, tl, fmaf(tv[c], tv[c], tr*tr));
            T1[c] = hT1[c][1] + h1; T2[c] = hT2[c][1] + h2;
            hT1[c][1] = hT1[c][0] + h1; hT1[c][0] = h1;
            hT2[c][1] = hT2[c][0] + h2; hT2[c][0] = h2;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int k0 = 2*q, k1 = (2*q + 1 < N) ? 2*q + 1 : 2*q;
                const float a0 = wv[k0][c], a1 = wv[k1][c];
                const float l0 = __shfl_up_sync(0xffffffffu, a0, 1), r0 = __shfl_down_sync(0xffffffffu, a0, 1);
                const float l1 = (k1 != k0) ? __shfl_up_sync(0xffffffffu, a1, 1) : l0;
                const float r1 = (k1 != k0) ? __shfl_down_sync(0xffffffffu, a1, 1) : r0;
                const f2 wl = mk2(l0, l1), wc = mk2(a0, a1), wr = mk2(r0, r1);
                const f2 g1 = wl + wc + wr;
                const f2 g2 = fma2(wl, wl, fma2(wc, wc, wr*wr));
                const f2 g3 = fma2(wl, splat2(tl), fma2(wc, splat2(tv[c]), wr*splat2(tr)));
                S1[q][c] = hS1[q][c][1] + g1; S2[q][c] = hS2[q][c][1] + g2; S3[q][c] = hS3[q][c][1] + g3;
                hS1[q][c][1] = hS1[q][c][0] + g1; hS1[q][c][0] = g1;
                hS2[q][c][1] = hS2[q][c][0] + g2; hS2[q][c][0] = g2;
                hS3[q][c][1] = hS3[q][c][0] + g3; hS3[q][c][0] = g3;
            }
        }

        // ---------------- stage B: centre row yc = y-1 (window rows y-2, y-1, y are in the running sums) ----------------
        int kcode = 255;          // decision of this lane's centre: frame index, or 255 = no gradient (static / not a centre)
        float csel[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) csel[j] = 0.f;
        if (it >= 2) {
            const int yc = y - 1;
            const bool centre = col_centre && yc >= 0 && yc < H;
            float e[N];
            f2 ca[NP][3], cb[NP][3], cc[NP][3];
#pragma unroll
            for (int k = 0; k < N; ++k) e[k] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float tcv = rc[(1 + c)*32];
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const int k0 = 2*q, k1 = (2*q + 1 < N) ? 2*q + 1 : 2*q;
                    if (p.w_ssim > 0.f) {
                        f2 eq;
                        fz_ssim2(S1[q][c], S2[q][c], S3[q][c], T1[c], T2[c], eq, ca[q][c], cb[q][c], cc[q][c]);
                        e[k0] = fmaf(ws3, lo2(eq), e[k0]);
                        if (k1 != k0) e[k1] = fmaf(ws3, hi2(eq), e[k1]);
                    } else ca[q][c] = cb[q][c] = cc[q][c] = splat2(0.f);
                    if (p.w_l1 > 0.f) {
                        e[k0] = fmaf(wl3, fabsf(rc[(5 + k0*9 + c)*32] - tcv), e[k0]);
                        if (k1 != k0) e[k1] = fmaf(wl3, fabsf(rc[(5 + k1*9 + c)*32] - tcv), e[k1]);
                    }
                }
            }
            float emin = e[0];
            int ks = 0;
#pragma unroll
            for (int k = 1; k < N; ++k) if (e[k] < emin) { emin = e[k]; ks = k; }  // first index wins ties (torch.min)
            bool is_static = false;
            if (p.use_automask && centre) {
                const size_t pix = (size_t)yc*W + x, nidx = ((size_t)s*p.b + i)*HW + pix;
                float e0v = e0_row;
                if (p.noise) e0v = fmaf(STV_EPS32, __ldg(p.noise + nidx), e0v);
                else if (seed) e0v = fmaf(STV_EPS32, hash_normal(seed, nidx), e0v);
                if (!(emin <= e0v)) { emin = e0v; is_static = true; }  // torch.min(cat(err, static)): index 0 wins ties
            }
            if (centre && !is_static) kcode = ks;
            if (centre && col_inner && yc >= y0 && yc < y0 + rows_here) {   // this strip owns the pixel
                selp[(size_t)yc*W + x] = (uint8_t)(is_static ? STV_SEL_STATIC : ks);
                loss_acc += emin;
            }
            if (GRAD && p.w_ssim > 0.f) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const f2 a2 = ca[k/2][c], b2 = cb[k/2][c], c2 = cc[k/2][c];
                        if (kcode == k) {
                            csel[c*3 + 0] = (k & 1) ? hi2(a2) : lo2(a2);
                            csel[c*3 + 1] = (k & 1) ? hi2(b2) : lo2(b2);
                            csel[c*3 + 2] = (k & 1) ? hi2(c2) : lo2(c2);
                        }
                    }
                }
            }
        }

        if (GRAD) {
            // masked horizontal sums of the selected coefficients per frame (reflection multiplicities on the image border),
            // then the vertical running sums -> coefficient window sums of row y-2
            const int kl = __shfl_up_sync(0xffffffffu, kcode, 1), kr = __shfl_down_sync(0xffffffffu, kcode, 1);
            // Window of pixel row yp = y-2: myu(yp) c[yp-1] + c[yp] + myd(yp) c[yp+1] with myu = 2 iff yp == 1, myd = 2 iff yp == H-2
            // (the reflect-padded window of centre 0 / H-1 holds row 1 / H-2 twice). hk below is c[yp+1]; q2 arrives holding
            // myu(yp) c[yp-1] + c[yp] and leaves holding myu(yp+1) c[yp] + c[yp+1] for the next pixel row.
            const int yp = y - 2;
            const float myd = (yp == H - 2) ? 2.f : 1.f, myu_next = (yp == 0) ? 2.f : 1.f;
            float cs[N][9];
            f2 ml[NP], mm[NP], mr[NP];   // per-frame masks of the left / own / right window, frames (2q, 2q+1) packed
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int k0 = 2*q, k1 = 2*q + 1;   // (k1 == N for an odd frame count: kcode is never N, the lane stays zero)
                ml[q] = mk2(kl == k0 ? mxl : 0.f, kl == k1 ? mxl : 0.f);
                mm[q] = mk2(kcode == k0 ? 1.f : 0.f, kcode == k1 ? 1.f : 0.f);
                mr[q] = mk2(kr == k0 ? mxr : 0.f, kr == k1 ? mxr : 0.f);
            }
            const f2 myd2 = splat2(myd), myu2 = splat2(myu_next);
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const float cl = __shfl_up_sync(0xffffffffu, csel[j], 1), cr = __shfl_down_sync(0xffffffffu, csel[j], 1);
                const f2 cl2 = splat2(cl), cr2 = splat2(cr), cm2 = splat2(csel[j]);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const f2 hk = fma2(ml[q], cl2, fma2(mr[q], cr2, mm[q]*cm2));
                    const f2 c2 = fma2(myd2, hk, vc[q][j][1]);
                    vc[q][j][1] = fma2(myu2, vc[q][j][0], hk);
                    vc[q][j][0] = hk;
                    cs[2*q][j] = lo2(c2);
                    if (2*q + 1 < N) cs[2*q + 1][j] = hi2(c2);
                }
            }
            // ---------------- stage C: pixel row yp = y-2 ----------------
            if (it >= 4 && col_inner) {
                const float dp = rp[0];
                float tp[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) tp[c] = rp[(1 + c)*32];
                const int ypa = yp;  // inside the image by construction (y0 <= yp < y0 + rows_here <= H)
                const float vp = (float)ypa;
                float rayp[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) { const float4 ki = cam4[r]; rayp[r] = fmaf(ki.x, u, fmaf(ki.y, vp, ki.z)); }
                const float4 k0p = cam4[3], k1p = cam4[4];
                const float cK0[3] = {k0p.x, k0p.y, k0p.z}, cK1[3] = {k1p.x, k1p.y, k1p.z};
                float gd = 0.f;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    float gqx = 0.f, gqy = 0.f;
                    bool any = false;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float w = rp[(5 + k*9 + c)*32];
                        float gw = gs*fmaf(2.f*w, cs[k][c*3 + 1], fmaf(tp[c], cs[k][c*3 + 2], cs[k][c*3 + 0]));
                        if (kcode_prev == k) {
                            const float df = w - tp[c];
                            gw += df != 0.f ? copysignf(gl, df) : 0.f;
                        }
                        any = any || gw != 0.f;
                        gqx = fmaf(gw, rp[(5 + k*9 + 3 + c)*32], gqx);
                        gqy = fmaf(gw, rp[(5 + k*9 + 6 + c)*32], gqy);
                    }
                    if (!any) continue;
                    // u = R ray, Q = d u + t. The depth gradient gQ . u is, written out, inv (gn.u) - inv^2 (gn.Q) u_z: two terms of
                    // size d |gn| / z that cancel to size |t| |gn| / z^2 (the d-terms cancel EXACTLY in exact arithmetic) — for far
                    // points a float32 evaluation of the difference loses log2(d/|t|) ~ 11 bits. The closed form below has the
                    // cancellation done analytically:  gd = inv^2 ((gn.u) t_z - (gn.t) u_z)   (unclamped z), inv (gn.u) (clamped).
                    float uu[3], Q[3], ct[3];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float4 rt = cam4[5 + k*3 + r];
                        uu[r] = fmaf(rt.x, rayp[0], fmaf(rt.y, rayp[1], rt.z*rayp[2]));
                        ct[r] = rt.w;
                        Q[r] = fmaf(dp, uu[r], ct[r]);
                    }
                    const float inv = rcp_fast(fmaxf(Q[2], STV_MIN_Z));
                    float gn[3], gQ[3];
                    float gz = 0.f, gnu = 0.f, gnt = 0.f;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        gn[r] = fmaf(cK0[r], gqx, cK1[r]*gqy); gQ[r] = gn[r]*inv;
                        gz = fmaf(gn[r], Q[r], gz); gnu = fmaf(gn[r], uu[r], gnu); gnt = fmaf(gn[r], ct[r], gnt);
                    }
                    const bool unclamped = Q[2] >= STV_MIN_Z;
                    if (unclamped) gQ[2] -= gz*inv*inv;
                    gd += unclamped ? inv*inv*fmaf(gnu, ct[2], -gnt*uu[2]) : inv*gnu;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float m = gQ[r]*dp;
                        accA[k][r] += m;
                        accB[k][r] = fmaf(m, vp, accB[k][r]);
                        accC[k][r] += gQ[r];
                        const float nr = Q[r]*inv;
                        accK[r] = fmaf(gqx, nr, accK[r]);
                        accK[3 + r] = fmaf(gqy, nr, accK[3 + r]);
                    }
                }
                gup[(size_t)ypa*W + x] = gd*rp[4*32];   // d loss/d src_up (mode 1: through to_scaled / to_inv)
            }
        }
        kcode_prev = kcode;   // the centre row of this iteration (y-1) is the pixel row of the next one
        slot_w = slot_w == 2 ? 0 : slot_w + 1;
    }

    // ---- per-strip reductions ----
    loss_acc = warp_sum(loss_acc);
    if (lane == 0) p.loss_partial[strip] = loss_acc;
    if (GRAD) {
        float* __restrict__ gp = p.gpart + (size_t)strip*LY::NPART;
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float m0 = warp_sum(accA[k][r]*u), m1 = warp_sum(accB[k][r]), m2 = warp_sum(accA[k][r]), m3 = warp_sum(accC[k][r]);
                if (lane == 0) { gp[k*12 + r*4 + 0] = m0; gp[k*12 + r*4 + 1] = m1; gp[k*12 + r*4 + 2] = m2; gp[k*12 + r*4 + 3] = m3; }
            }
        }
#pragma unroll
        for (int q = 0; q < FZ_NPART_K; ++q) {
            const float m = warp_sum(accK[q]);
            if (lane == 0) gp[N*12 + q] = m;
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Three-warp variant of the sweep (gradients on). The monolithic kernel above needs 255 registers: 8 resident warps per SM, each
// a chain of dependent instructions (~5 cycles per instruction), issue slots half idle with nothing to switch to. Here a strip
// is owned by THREE warps of a 96-thread block, one per stage, each carrying only its own state across rows:
//   warp G  stage A: loads, up-sampling, projection, gathers, bilinear warp            (state: the taps in flight)
//   warp S  stage B: window sums, SSIM + L1, min-reprojection, auto-mask, decision      (state: 3-row sums of w, w^2, w t)
//   warp C  masked coefficient window sums and stage C (the chain rule)                 (state: coefficient sums, pose moments)
// Rows travel G -> S -> C through shared memory: the lane-private value ring of the monolithic kernel, FZ_R slots deep, plus a
// decision ring (selected coefficients + decision code) S -> C. Hand-over is a producer / consumer ring on shared-memory
// mbarriers (named barriers would do, but 3 x FZ_R of them per block cap the resident blocks: an SM has 32-64 barrier slots):
// FULL_A[slot] G -> S, FULL_B[slot] S -> C, EMPTY[slot] C -> G and S once C no longer reads what they are about to overwrite
// (C is the last reader of everything). G may run FZ_R - 3 rows ahead of C.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FZ_R = 5;          // ring depth
constexpr int FZ_RB = 10;        // decision-ring values per lane: 9 selected coefficients + the decision code
constexpr int FZ_SPLIT_MAX_N = 2;
constexpr int FZ_SPLIT_NT = 96;
constexpr unsigned FZ_WAIT_HINT_NS = 2000;   // the hardware parks a waiting warp for up to this long instead of spinning on issue slots

template <int N> struct FzSplitLayout {
    using LY = FzLayout<N>;
    static constexpr int RING_A = FZ_R*LY::RV*32;
    static constexpr int RING_B = (FZ_R - 2)*FZ_RB*32;   // produced and consumed in the same iteration number
    static constexpr int BARS = 6*FZ_R;                  // 3 x FZ_R mbarriers (8 bytes each)
    static constexpr int PER_BLOCK = LY::CAM_PAD + RING_A + RING_B + BARS;   // floats
};

// Every lane arrives for itself (barriers count 32): each lane's ring accesses are ordered by its OWN releasing arrive, with no
// reliance on __syncwarp + one elected arrival being transitive (which compute-sanitizer's racecheck also cannot follow).
__device__ __forceinline__ void fz_signal(uint32_t bar, int) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void fz_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(FZ_WAIT_HINT_NS) : "memory");
    } while (!done);
}

template <int N, bool TEX, int MINB>
__global__ void __launch_bounds__(FZ_SPLIT_NT, MINB) photo_fused_split_kernel(const FusedParams p) {
    using LY = FzLayout<N>;
    using SL = FzSplitLayout<N>;
    constexpr int NP = (N + 1)/2;
    extern __shared__ __align__(16) float fz_smem[];
    const int lane = threadIdx.x & 31;
    const int role = threadIdx.x >> 5;   // 0: G, 1: S, 2: C
    float* const cam = fz_smem;
    float* const ring = cam + LY::CAM_PAD + lane;
    float* const ringb = cam + LY::CAM_PAD + SL::RING_A + lane;
    const uint32_t bars = (uint32_t)__cvta_generic_to_shared(cam + LY::CAM_PAD + SL::RING_A + SL::RING_B);
    const uint32_t bar_fa = bars, bar_fb = bars + 8*FZ_R, bar_e = bars + 16*FZ_R;   // FULL_A, FULL_B, EMPTY
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 3*FZ_R; ++q) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" :: "r"(bars + 8*q));
    }

    const long long strip = blockIdx.x;
    int t = (int)(strip % ((long long)p.nsx*p.nsy*p.S));
    const int i = (int)(strip/((long long)p.nsx*p.nsy*p.S));
    const int sxi = t % p.nsx; t /= p.nsx;
    const int syi = t % p.nsy;
    const int s = t/p.nsy;

    const int H = p.H, W = p.W, HW = H*W;
    const int x0 = sxi*FZ_COLS, y0 = syi*p.rows;
    const int x = x0 - 2 + lane;
    const int xa = fz_clampi(reflect_idx(x, W), 0, W - 1);
    const bool col_centre = lane >= 1 && lane <= 30 && x >= 0 && x < W;
    const bool col_inner = lane >= 2 && lane <= 29 && x < W;
    const float u = (float)xa;

    if (role == 0) {
        const float* __restrict__ Km = p.K + (size_t)i*16;
        const float* __restrict__ Ki = p.Kinv + (size_t)i*16;
        for (int q = lane; q < LY::CAM; q += 32) {
            float v;
            if (q < 12) v = (q & 3) < 3 ? __ldg(Ki + q) : 0.f;
            else if (q < 20) v = __ldg(Km + (q - 12));
            else {
                const int k = (q - 20)/12, e = (q - 20) % 12;
                v = __ldg(p.T + ((size_t)k*p.b + i)*16 + e);
            }
            cam[q] = v;
        }
    }
    __syncthreads();
    const float4* const cam4 = reinterpret_cast<const float4*>(cam);

    const int rows_here = min(p.rows, H - y0);
    const int n_it = rows_here + 4;
    const float inv_cnt = 1.f/((float)p.S*(float)p.b*(float)HW);
    const float ws3 = p.w_ssim*(1.f/3.f), wl3 = p.w_l1*(1.f/3.f);

    if (role == 0) {
        // =============================== warp G: stage A ===============================
        const float sx = (float)W/(float)(W - 1), sy = (float)H/(float)(H - 1);
        const int hs = p.h[s], ws = p.w[s];
        const bool resize = p.mode == 1 && (hs != H || ws != W);
        int lx0 = 0, lx1 = 0;
        float llx = 0.f;
        const float ry = (float)hs/(float)H;
        if (resize) {
            const float src = fmaxf(((float)ws/(float)W)*((float)xa + 0.5f) - 0.5f, 0.f);
            lx0 = min((int)src, ws - 1);
            lx1 = lx0 + (lx0 < ws - 1 ? 1 : 0);
            llx = src - (float)lx0;
        }
        const float* __restrict__ srcp = p.src[s] + (size_t)i*(p.mode == 1 ? hs*ws : HW);
        const float* __restrict__ tg = p.tgt + (size_t)i*3*HW;
        const bool want_warp = p.warp0 != nullptr && s == 0;

        struct RowLoads { float t[3], a00, a01, a10, a11, lly; };
        auto request = [&](int it_n, RowLoads& L) {
            const int yn = y0 - 2 + it_n;
            const int yan = fz_clampi(reflect_idx(yn, H), 0, H - 1);
            const int on = yan*W + xa;
#pragma unroll
            for (int c = 0; c < 3; ++c) L.t[c] = __ldg(tg + c*HW + on);
            if (resize) {
                const float src = fmaxf(ry*((float)yan + 0.5f) - 0.5f, 0.f);
                const int ly0 = min((int)src, hs - 1), ly1 = ly0 + (ly0 < hs - 1 ? 1 : 0);
                L.lly = src - (float)ly0;
                const float* r0 = srcp + ly0*ws; const float* r1 = srcp + ly1*ws;
                L.a00 = __ldg(r0 + lx0); L.a01 = __ldg(r0 + lx1); L.a10 = __ldg(r1 + lx0); L.a11 = __ldg(r1 + lx1);
            } else { L.a00 = __ldg(srcp + on); L.a01 = L.a10 = L.a11 = 0.f; L.lly = 0.f; }
        };
        // Software pipeline: iteration `it` finishes row `it` from the taps requested one iteration earlier, then computes the
        // sample positions of row it+1 and issues its gathers, then requests the plain loads of row it+2.
        struct Taps { float ta[3], tb[3], tc[3], td[3], fx, fy, bx, by; };
        struct RowA { float d, dchain, tv[3]; Taps tp[N]; };
        auto gather = [&](int it_r, const RowLoads& L, RowA& A) {
            const int yr = y0 - 2 + it_r;
            const float v = (float)fz_clampi(reflect_idx(yr, H), 0, H - 1);
#pragma unroll
            for (int c = 0; c < 3; ++c) A.tv[c] = L.t[c];
            A.dchain = 1.f;
            {
                const float sv = resize ? (1.f - L.lly)*((1.f - llx)*L.a00 + llx*L.a01) + L.lly*((1.f - llx)*L.a10 + llx*L.a11) : L.a00;
                if (p.mode == 1) {
                    const float dp = p.scaled ? __fadd_rn(__fmul_rn(p.d_mul, sv), p.d_add) : sv;
                    A.d = dp > 0.f ? 1.0f/fmaxf(dp, STV_EPS32) : 0.f;
                    A.dchain = (dp > 0.f && dp >= STV_EPS32) ? -(A.d*A.d)*(p.scaled ? p.d_mul : 1.f) : 0.f;
                } else A.d = sv;
            }
            float P[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) { const float4 ki = cam4[r]; P[r] = fmaf(ki.x, u, fmaf(ki.y, v, ki.z))*A.d; }
#pragma unroll
            for (int k = 0; k < N; ++k) {
                Taps& tp = A.tp[k];
                float Q[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) { const float4 rt = cam4[5 + k*3 + r]; Q[r] = fmaf(rt.x, P[0], fmaf(rt.y, P[1], fmaf(rt.z, P[2], rt.w))); }
                const float inv = rcp_fast(fmaxf(Q[2], STV_MIN_Z));  // max(max(z, eps), 0.1) == max(z, 0.1)
                const float nx = Q[0]*inv, ny = Q[1]*inv, nz = Q[2]*inv;
                const float4 k0 = cam4[3], k1 = cam4[4];
                float ix = fmaf(fmaf(k0.x, nx, fmaf(k0.y, ny, k0.z*nz)), sx, -0.5f);
                float iy = fmaf(fmaf(k1.x, nx, fmaf(k1.y, ny, k1.z*nz)), sy, -0.5f);
                const float mxc = (float)(W - 1), myc = (float)(H - 1);
                tp.bx = (ix > 0.f && ix < mxc) ? sx : 0.f; tp.by = (iy > 0.f && iy < myc) ? sy : 0.f;
                ix = fminf(fmaxf(ix, 0.f), mxc);
                iy = fminf(fmaxf(iy, 0.f), myc);
                const float x0f = floorf(ix), y0f = floorf(iy);
                tp.fx = ix - x0f; tp.fy = iy - y0f;
                const int plane0 = (k*p.b + i)*3;
                if (TEX) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float4 t4 = tex2Dgather<float4>((cudaTextureObject_t)p.tex, x0f + 1.f, y0f + 1.f + (float)((plane0 + c)*H), 0);
                        tp.ta[c] = t4.w; tp.tb[c] = t4.z; tp.tc[c] = t4.x; tp.td[c] = t4.y;
                    }
                } else {
                    const int xi0 = (int)x0f, yi0 = (int)y0f, xi1 = min(xi0 + 1, W - 1), yi1 = min(yi0 + 1, H - 1);
                    const float* __restrict__ sp = p.supp + (size_t)plane0*HW;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float* q = sp + c*HW;
                        tp.ta[c] = __ldg(q + yi0*W + xi0); tp.tb[c] = __ldg(q + yi0*W + xi1);
                        tp.tc[c] = __ldg(q + yi1*W + xi0); tp.td[c] = __ldg(q + yi1*W + xi1);
                    }
                }
            }
        };
        RowLoads ld;
        RowA cur;
        request(0, ld);
        gather(0, ld, cur);
        if (n_it > 1) request(1, ld);
        int slot_w = 0, slot_e = 0;
        uint32_t par_e = 0;
#pragma unroll 1   // (unrolling any of the three loops is slower: 83 KB of SASS run concurrently by the three roles)
        for (int it = 0; it < n_it; ++it) {
            // the slot about to be overwritten held row it - FZ_R, last read by warp C in its iteration it - FZ_R + 2
            if (it >= FZ_R - 2) {
                fz_wait(bar_e + 8*slot_e, par_e);
                if (++slot_e == FZ_R) { slot_e = 0; par_e ^= 1u; }
            }
            float* const rw = ring + slot_w*LY::RV*32;
            rw[0] = cur.d;
#pragma unroll
            for (int c = 0; c < 3; ++c) rw[(1 + c)*32] = cur.tv[c];
            rw[4*32] = cur.dchain;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const Taps& tp = cur.tp[k];
                float wv[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float top = fmaf(tp.fx, tp.tb[c] - tp.ta[c], tp.ta[c]), bot = fmaf(tp.fx, tp.td[c] - tp.tc[c], tp.tc[c]);
                    const float w = fmaf(tp.fy, bot - top, top);
                    wv[c] = w;
                    rw[(5 + k*9 + c)*32] = w;
                    rw[(5 + k*9 + 3 + c)*32] = fmaf(tp.fy, (tp.td[c] - tp.tc[c]) - (tp.tb[c] - tp.ta[c]), tp.tb[c] - tp.ta[c])*tp.bx;
                    rw[(5 + k*9 + 6 + c)*32] = (bot - top)*tp.by;
                }
                if (want_warp && col_inner && it >= 2 && it < rows_here + 2) {
                    const int plane0 = (k*p.b + i)*3;
                    const int ya = fz_clampi(reflect_idx(y0 - 2 + it, H), 0, H - 1);
#pragma unroll
                    for (int c = 0; c < 3; ++c) p.warp0[((size_t)plane0 + c)*HW + ya*W + xa] = wv[c];
                }
            }
            fz_signal(bar_fa + 8*slot_w, lane);            // FULL_A[slot]: row y is in the value ring
            if (it + 1 < n_it) {
                gather(it + 1, ld, cur);                   // row it+1: sample positions + gathers
                if (it + 2 < n_it) request(it + 2, ld);    // row it+2: target, source-map taps
            }
            slot_w = slot_w == FZ_R - 1 ? 0 : slot_w + 1;
        }
    } else if (role == 1) {
        // =============================== warp S: window sums and stage B ===============================
        uint8_t* __restrict__ selp = p.sel + ((size_t)s*p.b + i)*HW;
        const float* __restrict__ e0p = p.e0 + (size_t)i*HW;
        const uint64_t seed = p.seed ? p.seed + (p.step ? *p.step : 0ull) : 0ull;
        f2 hS1[NP][3][2], hS2[NP][3][2], hS3[NP][3][2];
        float hT1[3][2], hT2[3][2];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            hT1[c][0] = hT1[c][1] = hT2[c][0] = hT2[c][1] = 0.f;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                hS1[q][c][0] = hS1[q][c][1] = hS2[q][c][0] = hS2[q][c][1] = hS3[q][c][0] = hS3[q][c][1] = splat2(0.f);
            }
        }
        float loss_acc = 0.f;
        // identity error (and explicit tie-break noise) of the centre row an iteration decides: requested one iteration ahead
        auto request_e0 = [&](int it_n, float& e0, float& nz) {
            const int ycn = y0 - 3 + it_n;
            const bool ok = p.use_automask && col_centre && ycn >= 0 && ycn < H;
            e0 = ok ? __ldg(e0p + (size_t)ycn*W + x) : 0.f;
            nz = (ok && p.noise) ? __ldg(p.noise + ((size_t)s*p.b + i)*HW + (size_t)ycn*W + x) : 0.f;
        };
        float e0_next, nz_next;
        request_e0(0, e0_next, nz_next);
        int slot_w = 0, slot_c = FZ_R - 1, slot_b = 0, slot_e = 0;
        uint32_t par_f = 0, par_e = 0;
#pragma unroll 1
        for (int it = 0; it < n_it; ++it) {
            const int y = y0 - 2 + it;
            const float e0_row = e0_next, nz_row = nz_next;
            if (it + 1 < n_it) request_e0(it + 1, e0_next, nz_next);
            fz_wait(bar_fa + 8*slot_w, par_f);
            const float* const rw = ring + slot_w*LY::RV*32;
            const float* const rc = ring + slot_c*LY::RV*32;
            float tv[3], wv[N][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) tv[c] = rw[(1 + c)*32];
#pragma unroll
            for (int k = 0; k < N; ++k) {
#pragma unroll
                for (int c = 0; c < 3; ++c) wv[k][c] = rw[(5 + k*9 + c)*32];
            }

            float T1[3], T2[3];
            f2 S1[NP][3], S2[NP][3], S3[NP][3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float tl = __shfl_up_sync(0xffffffffu, tv[c], 1), tr = __shfl_down_sync(0xffffffffu, tv[c], 1);
                const float h1 = tl + tv[c] + tr;
                const float h2 = fmaf(tl, tl, fmaf(tv[c], tv[c], tr*tr));
                T1[c] = hT1[c][1] + h1; T2[c] = hT2[c][1] + h2;
                hT1[c][1] = hT1[c][0] + h1; hT1[c][0] = h1;
                hT2[c][1] = hT2[c][0] + h2; hT2[c][0] = h2;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const int k0 = 2*q, k1 = (2*q + 1 < N) ? 2*q + 1 : 2*q;
                    const float a0 = wv[k0][c], a1 = wv[k1][c];
                    const float l0 = __shfl_up_sync(0xffffffffu, a0, 1), r0 = __shfl_down_sync(0xffffffffu, a0, 1);
                    const float l1 = (k1 != k0) ? __shfl_up_sync(0xffffffffu, a1, 1) : l0;
                    const float r1 = (k1 != k0) ? __shfl_down_sync(0xffffffffu, a1, 1) : r0;
                    const f2 wl = mk2(l0, l1), wc = mk2(a0, a1), wr = mk2(r0, r1);
                    const f2 g1 = wl + wc + wr;
                    const f2 g2 = fma2(wl, wl, fma2(wc, wc, wr*wr));
                    const f2 g3 = fma2(wl, splat2(tl), fma2(wc, splat2(tv[c]), wr*splat2(tr)));
                    S1[q][c] = hS1[q][c][1] + g1; S2[q][c] = hS2[q][c][1] + g2; S3[q][c] = hS3[q][c][1] + g3;
                    hS1[q][c][1] = hS1[q][c][0] + g1; hS1[q][c][0] = g1;
                    hS2[q][c][1] = hS2[q][c][0] + g2; hS2[q][c][0] = g2;
                    hS3[q][c][1] = hS3[q][c][0] + g3; hS3[q][c][0] = g3;
                }
            }

            // ---------------- stage B: centre row yc = y-1 ----------------
            int kcode = 255;
            float csel[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) csel[j] = 0.f;
            if (it >= 2) {
                const int yc = y - 1;
                const bool centre = col_centre && yc >= 0 && yc < H;
                float e[N];
                f2 ca[NP][3], cb[NP][3], cc[NP][3];
#pragma unroll
                for (int k = 0; k < N; ++k) e[k] = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float tcv = rc[(1 + c)*32];
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        const int k0 = 2*q, k1 = (2*q + 1 < N) ? 2*q + 1 : 2*q;
                        if (p.w_ssim > 0.f) {
                            f2 eq;
                            fz_ssim2(S1[q][c], S2[q][c], S3[q][c], T1[c], T2[c], eq, ca[q][c], cb[q][c], cc[q][c]);
                            e[k0] = fmaf(ws3, lo2(eq), e[k0]);
                            if (k1 != k0) e[k1] = fmaf(ws3, hi2(eq), e[k1]);
                        } else ca[q][c] = cb[q][c] = cc[q][c] = splat2(0.f);
                        if (p.w_l1 > 0.f) {
                            e[k0] = fmaf(wl3, fabsf(rc[(5 + k0*9 + c)*32] - tcv), e[k0]);
                            if (k1 != k0) e[k1] = fmaf(wl3, fabsf(rc[(5 + k1*9 + c)*32] - tcv), e[k1]);
                        }
                    }
                }
                float emin = e[0];
                int ks = 0;
#pragma unroll
                for (int k = 1; k < N; ++k) if (e[k] < emin) { emin = e[k]; ks = k; }  // first index wins ties (torch.min)
                bool is_static = false;
                if (p.use_automask && centre) {
                    const size_t nidx = ((size_t)s*p.b + i)*HW + (size_t)yc*W + x;
                    float e0v = e0_row;
                    if (p.noise) e0v = fmaf(STV_EPS32, nz_row, e0v);
                    else if (seed) e0v = fmaf(STV_EPS32, hash_normal(seed, nidx), e0v);
                    if (!(emin <= e0v)) { emin = e0v; is_static = true; }  // torch.min(cat(err, static)): index 0 wins ties
                }
                if (centre && !is_static) kcode = ks;
                if (centre && col_inner && yc >= y0 && yc < y0 + rows_here) {   // this strip owns the pixel
                    selp[(size_t)yc*W + x] = (uint8_t)(is_static ? STV_SEL_STATIC : ks);
                    loss_acc += emin;
                }
                if (p.w_ssim > 0.f) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            const f2 a2 = ca[k/2][c], b2 = cb[k/2][c], c2 = cc[k/2][c];
                            if (kcode == k) {
                                csel[c*3 + 0] = (k & 1) ? hi2(a2) : lo2(a2);
                                csel[c*3 + 1] = (k & 1) ? hi2(b2) : lo2(b2);
                                csel[c*3 + 2] = (k & 1) ? hi2(c2) : lo2(c2);
                            }
                        }
                    }
                }
            }
            // the decision-ring slot held iteration it - (FZ_R - 2): wait until warp C has consumed it
            if (it >= FZ_R - 2) {
                fz_wait(bar_e + 8*slot_e, par_e);
                if (++slot_e == FZ_R) { slot_e = 0; par_e ^= 1u; }
            }
            float* const rb = ringb + slot_b*FZ_RB*32;
#pragma unroll
            for (int j = 0; j < 9; ++j) rb[j*32] = csel[j];
            rb[9*32] = __int_as_float(kcode);
            fz_signal(bar_fb + 8*slot_w, lane);   // FULL_B[slot]: the decision of row y-1 is in shared memory
            slot_c = slot_w;
            slot_b = slot_b == FZ_R - 3 ? 0 : slot_b + 1;
            if (++slot_w == FZ_R) { slot_w = 0; par_f ^= 1u; }
        }
        loss_acc = warp_sum(loss_acc);
        if (lane == 0) p.loss_partial[strip] = loss_acc;
    } else {
        // =============================== warp C: coefficient window sums and stage C ===============================
        const float mxl = (x == 1) ? 2.f : 1.f, mxr = (x == W - 2) ? 2.f : 1.f;
        const float gs = ws3*inv_cnt, gl = wl3*inv_cnt;
        float* __restrict__ gup = p.g_unit[s] + (size_t)i*HW;
        f2 vc[NP][9][2];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
#pragma unroll
            for (int j = 0; j < 9; ++j) vc[q][j][0] = vc[q][j][1] = splat2(0.f);
        }
        float accA[N][3], accB[N][3], accC[N][3], accK[FZ_NPART_K];
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) accA[k][r] = accB[k][r] = accC[k][r] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < FZ_NPART_K; ++q) accK[q] = 0.f;
        int kcode_prev = 255;
        int slot_w = 0, slot_p = FZ_R - 2;   // slot of iteration `it`, of row y-2
        int slot_b = 0;
        uint32_t par_f = 0;
#pragma unroll 1
        for (int it = 0; it < n_it; ++it) {
            const int y = y0 - 2 + it;
            fz_wait(bar_fb + 8*slot_w, par_f);
            const float* const rp = ring + slot_p*LY::RV*32;
            const float* const rb = ringb + slot_b*FZ_RB*32;
            const int kcode = __float_as_int(rb[9*32]);
            float csel[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) csel[j] = rb[j*32];

            const int kl = __shfl_up_sync(0xffffffffu, kcode, 1), kr = __shfl_down_sync(0xffffffffu, kcode, 1);
            // Window of pixel row yp = y-2: myu(yp) c[yp-1] + c[yp] + myd(yp) c[yp+1] (see the monolithic kernel)
            const int yp = y - 2;
            const float myd = (yp == H - 2) ? 2.f : 1.f, myu_next = (yp == 0) ? 2.f : 1.f;
            float cs[N][9];
            f2 ml[NP], mm[NP], mr[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int k0 = 2*q, k1 = 2*q + 1;
                ml[q] = mk2(kl == k0 ? mxl : 0.f, kl == k1 ? mxl : 0.f);
                mm[q] = mk2(kcode == k0 ? 1.f : 0.f, kcode == k1 ? 1.f : 0.f);
                mr[q] = mk2(kr == k0 ? mxr : 0.f, kr == k1 ? mxr : 0.f);
            }
            const f2 myd2 = splat2(myd), myu2 = splat2(myu_next);
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const float cl = __shfl_up_sync(0xffffffffu, csel[j], 1), cr = __shfl_down_sync(0xffffffffu, csel[j], 1);
                const f2 cl2 = splat2(cl), cr2 = splat2(cr), cm2 = splat2(csel[j]);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const f2 hk = fma2(ml[q], cl2, fma2(mr[q], cr2, mm[q]*cm2));
                    const f2 c2 = fma2(myd2, hk, vc[q][j][1]);
                    vc[q][j][1] = fma2(myu2, vc[q][j][0], hk);
                    vc[q][j][0] = hk;
                    cs[2*q][j] = lo2(c2);
                    if (2*q + 1 < N) cs[2*q + 1][j] = hi2(c2);
                }
            }
            // ---------------- stage C: pixel row yp = y-2 ----------------
            if (it >= 4 && col_inner) {
                const float dp = rp[0];
                float tp[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) tp[c] = rp[(1 + c)*32];
                const float vp = (float)yp;   // inside the image by construction (y0 <= yp < y0 + rows_here <= H)
                float rayp[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) { const float4 ki = cam4[r]; rayp[r] = fmaf(ki.x, u, fmaf(ki.y, vp, ki.z)); }
                const float4 k0p = cam4[3], k1p = cam4[4];
                const float cK0[3] = {k0p.x, k0p.y, k0p.z}, cK1[3] = {k1p.x, k1p.y, k1p.z};
                float gd = 0.f;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    float gqx = 0.f, gqy = 0.f;
                    bool any = false;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float w = rp[(5 + k*9 + c)*32];
                        float gw = gs*fmaf(2.f*w, cs[k][c*3 + 1], fmaf(tp[c], cs[k][c*3 + 2], cs[k][c*3 + 0]));
                        if (kcode_prev == k) {
                            const float df = w - tp[c];
                            gw += df != 0.f ? copysignf(gl, df) : 0.f;
                        }
                        any = any || gw != 0.f;
                        gqx = fmaf(gw, rp[(5 + k*9 + 3 + c)*32], gqx);
                        gqy = fmaf(gw, rp[(5 + k*9 + 6 + c)*32], gqy);
                    }
                    if (!any) continue;
                    float uu[3], Q[3], ct[3];   // closed-form depth gradient: see the monolithic kernel
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float4 rt = cam4[5 + k*3 + r];
                        uu[r] = fmaf(rt.x, rayp[0], fmaf(rt.y, rayp[1], rt.z*rayp[2]));
                        ct[r] = rt.w;
                        Q[r] = fmaf(dp, uu[r], ct[r]);
                    }
                    const float inv = rcp_fast(fmaxf(Q[2], STV_MIN_Z));
                    float gn[3], gQ[3];
                    float gz = 0.f, gnu = 0.f, gnt = 0.f;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        gn[r] = fmaf(cK0[r], gqx, cK1[r]*gqy); gQ[r] = gn[r]*inv;
                        gz = fmaf(gn[r], Q[r], gz); gnu = fmaf(gn[r], uu[r], gnu); gnt = fmaf(gn[r], ct[r], gnt);
                    }
                    const bool unclamped = Q[2] >= STV_MIN_Z;
                    if (unclamped) gQ[2] -= gz*inv*inv;
                    gd += unclamped ? inv*inv*fmaf(gnu, ct[2], -gnt*uu[2]) : inv*gnu;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float m = gQ[r]*dp;
                        accA[k][r] += m;
                        accB[k][r] = fmaf(m, vp, accB[k][r]);
                        accC[k][r] += gQ[r];
                        const float nr = Q[r]*inv;
                        accK[r] = fmaf(gqx, nr, accK[r]);
                        accK[3 + r] = fmaf(gqy, nr, accK[3 + r]);
                    }
                }
                gup[(size_t)yp*W + x] = gd*rp[4*32];
            }
            kcode_prev = kcode;
            // EMPTY[slot]: this iteration's decision record and row y-2 are consumed
            fz_signal(bar_e + 8*slot_w, lane);
            slot_p = slot_p == FZ_R - 1 ? 0 : slot_p + 1;
            slot_b = slot_b == FZ_R - 3 ? 0 : slot_b + 1;
            if (++slot_w == FZ_R) { slot_w = 0; par_f ^= 1u; }
        }
        float* __restrict__ gp = p.gpart + (size_t)strip*LY::NPART;
#pragma unroll
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float m0 = warp_sum(accA[k][r]*u), m1 = warp_sum(accB[k][r]), m2 = warp_sum(accA[k][r]), m3 = warp_sum(accC[k][r]);
                if (lane == 0) { gp[k*12 + r*4 + 0] = m0; gp[k*12 + r*4 + 1] = m1; gp[k*12 + r*4 + 2] = m2; gp[k*12 + r*4 + 3] = m3; }
            }
        }
#pragma unroll
        for (int q = 0; q < FZ_NPART_K; ++q) {
            const float m = warp_sum(accK[q]);
            if (lane == 0) gp[N*12 + q] = m;
        }
    }
}

}  // namespace stv

// ---------------------------------------------------------------------------------------------------------------------
// Backward side: nothing is recomputed. (1) pose / intrinsics: per-strip moment partials -> gT, gK, gKinv (fixed order,
// double accumulation), scaled by dL/dloss; (2) per-pixel maps: scale by dL/dloss, and — when the kernel consumed the
// low-resolution disparities — pull the full-resolution map back through the bilinear up-sampling (two separable gathers).
// ---------------------------------------------------------------------------------------------------------------------
namespace stv {

// One block per image. gpart rows of image i are contiguous: [i*spi, (i+1)*spi) x npart.
//   M_k[r][j] = sum gQ_r d (u,v,1)_j (j<3), M_k[r][3] = sum gQ_r;   gT[k,i][r][c<3] = sum_j Kinv[c][j] M_k[r][j], gT[r][3] = M_k[r][3]
//   gK[i][0][c] = part[n*12 + c], gK[i][1][c] = part[n*12 + 3 + c];     gKinv[i][c][j] = sum_k sum_r R_k[r][c] M_k[r][j]
__global__ void __launch_bounds__(256) fused_finalize_kernel(const float* __restrict__ gpart, int spi, int n, int b,
                                                             const float* __restrict__ T, const float* __restrict__ Kinv,
                                                             const float* __restrict__ grad_loss, float* __restrict__ gT,
                                                             float* __restrict__ gK, float* __restrict__ gKinv) {
    __shared__ double sums[STV_MAX_SUPPORT*12 + FZ_NPART_K];
    const int i = blockIdx.x, npart = n*12 + FZ_NPART_K;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float* __restrict__ base = gpart + (size_t)i*spi*npart;
    for (int q = wid; q < npart; q += nw) {
        double a = 0.0;
        for (int r = lane; r < spi; r += 32) a += (double)base[(size_t)r*npart + q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) sums[q] = a;
    }
    __syncthreads();
    const double g = (double)__ldg(grad_loss);
    const float* __restrict__ Ki = Kinv + (size_t)i*16;
    for (int e = threadIdx.x; e < n*16; e += blockDim.x) {   // gT[k,i]
        const int k = e/16, r = (e/4) & 3, c = e & 3;
        double v = 0.0;
        if (r < 3) {
            const double* M = sums + k*12 + r*4;
            v = c < 3 ? (double)Ki[c*4 + 0]*M[0] + (double)Ki[c*4 + 1]*M[1] + (double)Ki[c*4 + 2]*M[2] : M[3];
        }
        gT[((size_t)k*b + i)*16 + r*4 + c] = (float)(g*v);
    }
    if (gK != nullptr) {
        for (int e = threadIdx.x; e < 16; e += blockDim.x) {
            const int r = e/4, c = e & 3;
            gK[(size_t)i*16 + e] = (r < 2 && c < 3) ? (float)(g*sums[n*12 + r*3 + c]) : 0.f;
        }
    }
    if (gKinv != nullptr) {
        for (int e = threadIdx.x; e < 16; e += blockDim.x) {
            const int c = e/4, j = e & 3;
            double v = 0.0;
            if (c < 3 && j < 3) {
                for (int k = 0; k < n; ++k) {
                    const float* __restrict__ Tm = T + ((size_t)k*b + i)*16;
                    for (int r = 0; r < 3; ++r) v += (double)Tm[r*4 + c]*sums[k*12 + r*4 + j];
                }
            }
            gKinv[(size_t)i*16 + e] = (float)(g*v);
        }
    }
}

struct PullParams {
    int b, S, H, W;
    int h[STV_MAX_SCALES], w[STV_MAX_SCALES];
    const float* g_full[STV_MAX_SCALES];     // (b,1,H,W) gradient w.r.t. the up-sampled map
    const float* g_full2[STV_MAX_SCALES];    // optional second full-resolution gradient (added after the chain factor), or NULL
    const float* disp[STV_MAX_SCALES];       // low-resolution disparities (chain mode), or NULL
    float* tmp[STV_MAX_SCALES];              // (b,H,w) horizontally pulled rows (unused when w == W and h == H)
    float* out[STV_MAX_SCALES];              // (b,1,h,w)
    const float* scale;                      // device scalar multiplying everything (dL/dloss), or NULL
    float d_mul, d_add;
    int scaled, chain;                       // chain: g_full is d/d depth_up; multiply by d depth/d disp_up (recomputed from disp)
};

__device__ __forceinline__ void pull_tap(int dst, float scale, int n_in, int& i0, int& i1, float& lam) {
    const float src = fmaxf(scale*((float)dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)src, n_in - 1);
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    lam = src - (float)i0;
}

// Range of output indices whose taps can include input index `il`.
__device__ __forceinline__ void pull_range(int il, float r, int n_out, int& lo, int& hi) {
    lo = il == 0 ? 0 : max(0, (int)floorf(((float)il - 0.5f)/r - 0.5f) - 1);
    hi = min(n_out - 1, (int)ceilf(((float)il + 1.5f)/r - 0.5f) + 1);
}

// Value of the full-resolution gradient at (Y, X) of image i, scale s, including the optional depth chain.
__device__ __forceinline__ float pull_value(const PullParams& p, int s, int i, int Y, int X) {
    const int H = p.H, W = p.W, h = p.h[s], w = p.w[s];
    const size_t o = (size_t)i*H*W + (size_t)Y*W + X;
    float g = p.g_full[s] ? __ldg(p.g_full[s] + o) : 0.f;
    if (p.chain && p.g_full[s]) {
        int y0, y1, x0, x1;
        float ly, lx;
        pull_tap(Y, (float)h/(float)H, h, y0, y1, ly);
        pull_tap(X, (float)w/(float)W, w, x0, x1, lx);
        const float* q = p.disp[s] + (size_t)i*h*w;
        const float v = (1.f - ly)*((1.f - lx)*__ldg(q + y0*w + x0) + lx*__ldg(q + y0*w + x1)) +
                        ly*((1.f - lx)*__ldg(q + y1*w + x0) + lx*__ldg(q + y1*w + x1));
        const float dp = p.scaled ? __fadd_rn(__fmul_rn(p.d_mul, v), p.d_add) : v;
        if (dp > 0.f && dp >= STV_EPS32) { const float inv = 1.0f/dp; g = -g*inv*inv*(p.scaled ? p.d_mul : 1.f); }
        else g = 0.f;
    }
    if (p.g_full2[s]) g += __ldg(p.g_full2[s] + o);
    return g;
}

// Pass 1: tmp[i][Y][xl] = sum_X wx(X, xl) g(Y, X).  grid = (ceil(w_max/128), H, S*b); scales with w == W copy straight to `out`
// (h == H is implied for them by the callers: no resize at all).
__global__ void __launch_bounds__(128) pull_rows_kernel(const PullParams p) {
    const int s = blockIdx.z/p.b, i = blockIdx.z - s*p.b;
    const int w = p.w[s], W = p.W, H = p.H, Y = blockIdx.y;
    const int xl = blockIdx.x*blockDim.x + threadIdx.x;
    if (xl >= w) return;
    const float sc = p.scale ? __ldg(p.scale) : 1.f;
    if (w == W && p.h[s] == H) return;   // not resized: written by pull_scale_kernel
    const float rx = (float)w/(float)W;
    int lo, hi;
    pull_range(xl, rx, W, lo, hi);
    float acc = 0.f;
    for (int X = lo; X <= hi; ++X) {
        int x0, x1;
        float lx;
        pull_tap(X, rx, w, x0, x1, lx);
        const float wx = (x0 == xl ? 1.f - lx : 0.f) + (x1 == xl ? lx : 0.f);
        if (wx != 0.f) acc = fmaf(wx, pull_value(p, s, i, Y, X), acc);
    }
    p.tmp[s][((size_t)i*H + Y)*w + xl] = sc*acc;
}

// Pass 2: out[i][yl][xl] = sum_Y wy(Y, yl) tmp[i][Y][xl].  grid = (ceil(w_max/128), h_max, S*b)
__global__ void __launch_bounds__(128) pull_cols_kernel(const PullParams p) {
    const int s = blockIdx.z/p.b, i = blockIdx.z - s*p.b;
    const int w = p.w[s], h = p.h[s], H = p.H, yl = blockIdx.y;
    const int xl = blockIdx.x*blockDim.x + threadIdx.x;
    if (xl >= w || yl >= h || (w == p.W && h == H)) return;
    const float ry = (float)h/(float)H;
    int lo, hi;
    pull_range(yl, ry, H, lo, hi);
    float acc = 0.f;
    for (int Y = lo; Y <= hi; ++Y) {
        int y0, y1;
        float ly;
        pull_tap(Y, ry, h, y0, y1, ly);
        const float wy = (y0 == yl ? 1.f - ly : 0.f) + (y1 == yl ? ly : 0.f);
        if (wy != 0.f) acc = fmaf(wy, __ldg(p.tmp[s] + ((size_t)i*H + Y)*w + xl), acc);
    }
    p.out[s][((size_t)i*h + yl)*w + xl] = acc;
}

// Tiled pull-back: one block owns a px x py tile of the low-resolution output of one (scale, image), stages the
// full-resolution footprint of the tile in shared memory with coalesced row loads (chain factor and dL/dloss applied on the way
// in), pulls it horizontally into px columns and then vertically into py rows — every full-resolution value is read from
// HBM once (plus the footprint overlap between neighbouring tiles), fixed summation order (deterministic).
// The bilinear tap weights of the tile's columns / rows are tabulated once per block (the same float arithmetic as the forward
// up-sampling, pull_tap), so the two accumulation loops are one shared-memory weight, one shared-memory value and one FMA per tap.
constexpr int PT_X = 32, PT_Y = 16, PT_NT = 128;
constexpr int PT_MAX_FOOT = 9216;    // floats of footprint + row-pulled scratch + tap tables a block may stage (36 KB: 6 blocks per SM)

struct PullTiles {   // block ranges + tile shape + taps per output (table pitch) per scale
    int first[STV_MAX_SCALES + 1];
    int tx[STV_MAX_SCALES], ty[STV_MAX_SCALES], px[STV_MAX_SCALES], py[STV_MAX_SCALES], kw[STV_MAX_SCALES], kh[STV_MAX_SCALES];
};

// Taps of output index `il` along one axis: inputs [lo, lo + cnt) with weights wt[0..cnt) (zero where an input does not touch il).
__device__ __forceinline__ void pull_taps(int il, float r, int n_in_lowres, int n_out, int pitch, int& lo, int& cnt, float* wt) {
    int hi;
    pull_range(il, r, n_out, lo, hi);
    cnt = min(hi - lo + 1, pitch);
    for (int j = 0; j < cnt; ++j) {
        int i0, i1;
        float lam;
        pull_tap(lo + j, r, n_in_lowres, i0, i1, lam);
        wt[j] = (i0 == il ? 1.f - lam : 0.f) + (i1 == il ? lam : 0.f);
    }
}

__global__ void __launch_bounds__(PT_NT) pull_tile_kernel(const PullParams p, const PullTiles t) {
    extern __shared__ float pt_smem[];
    int s = 0;
    while (s + 1 < p.S && (int)blockIdx.x >= t.first[s + 1]) ++s;
    int rem = blockIdx.x - t.first[s];
    const int per_img = t.tx[s]*t.ty[s];
    const int i = rem/per_img; rem -= i*per_img;
    const int tyi = rem/t.tx[s], txi = rem - tyi*t.tx[s];
    const int H = p.H, W = p.W, h = p.h[s], w = p.w[s];
    const float sc = p.scale ? __ldg(p.scale) : 1.f;
    const int xl0 = txi*t.px[s], yl0 = tyi*t.py[s];
    const int nx = min(t.px[s], w - xl0), ny = min(t.py[s], h - yl0);
    if (w == W && h == H) {   // no resize: scale (and chain) only
        for (int q = threadIdx.x; q < nx*ny; q += PT_NT) {
            const int yy = yl0 + q/nx, xx = xl0 + q % nx;
            p.out[s][(size_t)i*H*W + (size_t)yy*W + xx] = sc*pull_value(p, s, i, yy, xx);
        }
        return;
    }
    const float rx = (float)w/(float)W, ry = (float)h/(float)H;
    const int KW = t.kw[s], KH = t.kh[s];
    int Xlo, Xhi, Ylo, Yhi, tmp_;
    pull_range(xl0, rx, W, Xlo, tmp_); pull_range(xl0 + nx - 1, rx, W, tmp_, Xhi);
    pull_range(yl0, ry, H, Ylo, tmp_); pull_range(yl0 + ny - 1, ry, H, tmp_, Yhi);
    const int fw = Xhi - Xlo + 1, fh = Yhi - Ylo + 1;
    float* foot = pt_smem;                  // [fh][fw]
    float* rows = foot + fh*fw;             // [fh][nx]
    float* wxt = rows + fh*nx;              // [nx][KW]
    float* wyt = wxt + nx*KW;               // [ny][KH]
    int* xlo = (int*)(wyt + ny*KH);         // [nx] first footprint column, [nx] taps
    int* xcn = xlo + nx;
    int* ylo = xcn + nx;                    // [ny] first footprint row, [ny] taps
    int* ycn = ylo + ny;
    for (int q = threadIdx.x; q < nx + ny; q += PT_NT) {
        int lo, cnt;
        if (q < nx) { pull_taps(xl0 + q, rx, w, W, KW, lo, cnt, wxt + q*KW); xlo[q] = lo - Xlo; xcn[q] = cnt; }
        else { const int yy = q - nx; pull_taps(yl0 + yy, ry, h, H, KH, lo, cnt, wyt + yy*KH); ylo[yy] = lo - Ylo; ycn[yy] = cnt; }
    }
    {   // footprint rows, coalesced; the plain case (no chain, no second map) is a bare scaled copy
        const bool plain = !p.chain && p.g_full2[s] == nullptr && p.g_full[s] != nullptr;
        const float* __restrict__ g = p.g_full[s] + (size_t)i*H*W + (size_t)Ylo*W + Xlo;
        for (int fy = threadIdx.x/32; fy < fh; fy += PT_NT/32) {
            for (int fx = threadIdx.x & 31; fx < fw; fx += 32)
                foot[fy*fw + fx] = plain ? sc*__ldg(g + (size_t)fy*W + fx) : sc*pull_value(p, s, i, Ylo + fy, Xlo + fx);
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < fh*nx; q += PT_NT) {
        const int fy = q/nx, xx = q - fy*nx;
        const float* __restrict__ f = foot + fy*fw + xlo[xx];
        const float* __restrict__ wt = wxt + xx*KW;
        const int cnt = xcn[xx];
        float acc = 0.f;
        for (int j = 0; j < cnt; ++j) acc = fmaf(wt[j], f[j], acc);
        rows[q] = acc;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < ny*nx; q += PT_NT) {
        const int yy = q/nx, xx = q - yy*nx;
        const float* __restrict__ f = rows + ylo[yy]*nx + xx;
        const float* __restrict__ wt = wyt + yy*KH;
        const int cnt = ycn[yy];
        float acc = 0.f;
        for (int j = 0; j < cnt; ++j) acc = fmaf(wt[j], f[j*nx], acc);
        p.out[s][((size_t)i*h + yl0 + yy)*w + xl0 + xx] = acc;
    }
}

// Scales that are not resized: out = dL/dloss * g (float4 streams when possible).
__global__ void __launch_bounds__(256) pull_scale_kernel(const PullParams p, int s, long long n) {
    const float sc = p.scale ? __ldg(p.scale) : 1.f;
    const long long stride = (long long)gridDim.x*blockDim.x, t0 = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    const bool plain = !p.chain && p.g_full2[s] == nullptr && p.g_full[s] != nullptr;
    if (plain && (n & 3) == 0 && (((uintptr_t)p.g_full[s] | (uintptr_t)p.out[s]) & 15) == 0) {
        const float4* g4 = reinterpret_cast<const float4*>(p.g_full[s]);
        float4* o4 = reinterpret_cast<float4*>(p.out[s]);
        for (long long q = t0; q < n/4; q += stride) { float4 v = __ldg(g4 + q); v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; o4[q] = v; }
        return;
    }
    const long long hw = (long long)p.H*p.W;
    for (long long q = t0; q < n; q += stride) {
        const int i = (int)(q/hw);
        const int rem = (int)(q - (long long)i*hw), Y = rem/p.W, X = rem - Y*p.W;
        p.out[s][q] = sc*pull_value(p, s, i, Y, X);
    }
}

int launch_pull(const PullParams& p, cudaStream_t st) {
    // Tiled single pass when every scale's tile footprint fits the staging budget (always for the 2^s pyramids of the networks);
    // otherwise the two separable passes through `tmp`.
    PullTiles t{};
    bool tiled = true;
    int blocks = 0;
    for (int s = 0; s < p.S; ++s) {
        t.first[s] = blocks;
        if (p.w[s] == p.W && p.h[s] == p.H) {   // not resized: one streaming kernel, no tiles
            const long long n = (long long)p.b*p.H*p.W;
            long long nb = (n/4 + 255)/256;
            if (nb > 148*8) nb = 148*8;
            pull_scale_kernel<<<(unsigned)(nb < 1 ? 1 : nb), 256, 0, st>>>(p, s, n);
            count_launch();
            if (int rc = check_launch("pull_scale_kernel")) return rc;
            t.px[s] = PT_X; t.py[s] = PT_Y; t.tx[s] = t.ty[s] = 0;
            continue;
        }
        int px = PT_X, py = PT_Y;
        {   // shrink the tile until its full-resolution footprint + scratch + tap tables fit the staging budget
            const double fx = (double)p.W/p.w[s], fy = (double)p.H/p.h[s];
            t.kw[s] = (int)(2*fx) + 6; t.kh[s] = (int)(2*fy) + 6;   // pull_range spans at most 2f + 5 inputs
            auto need = [&](int ax, int ay) {
                const double fw = (ax + 2)*fx + 6, fh = (ay + 2)*fy + 6;
                return fw*fh + fh*ax + ax*t.kw[s] + ay*t.kh[s] + 2*(ax + ay) + 8;
            };
            while (need(px, py) > PT_MAX_FOOT && (px > 8 || py > 1)) { if (py > 1 && (py >= 4 || px <= 8)) py /= 2; else px /= 2; }
            if (need(px, py) > PT_MAX_FOOT) tiled = false;
        }
        t.px[s] = px; t.py[s] = py;
        t.tx[s] = (p.w[s] + px - 1)/px; t.ty[s] = (p.h[s] + py - 1)/py;
        blocks += t.tx[s]*t.ty[s]*p.b;
    }
    t.first[p.S] = blocks;
    if (blocks == 0) return STV_OK;
    if (tiled) {
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(pull_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_MAX_FOOT*(int)sizeof(float)); attr = true; }
        pull_tile_kernel<<<blocks, PT_NT, PT_MAX_FOOT*sizeof(float), st>>>(p, t);
        count_launch();
        return check_launch("pull_tile_kernel");
    }
    int wmax = 0, hmax = 0;
    bool any_resize = false;
    for (int s = 0; s < p.S; ++s) {
        wmax = max(wmax, p.w[s]); hmax = max(hmax, p.h[s]);
        any_resize = any_resize || p.w[s] != p.W || p.h[s] != p.H;
    }
    pull_rows_kernel<<<dim3((wmax + 127)/128, p.H, p.S*p.b), 128, 0, st>>>(p);
    count_launch();
    if (int rc = check_launch("pull_rows_kernel")) return rc;
    if (any_resize) {
        pull_cols_kernel<<<dim3((wmax + 127)/128, hmax, p.S*p.b), 128, 0, st>>>(p);
        count_launch();
        if (int rc = check_launch("pull_cols_kernel")) return rc;
    }
    return STV_OK;
}

__global__ void fused_loss_reduce_kernel(const float* __restrict__ partial, long long n, double inv_count, float* __restrict__ out,
                                         unsigned long long* step) {
    __shared__ double sh[256];
    double a = 0.0;
    for (long long q = threadIdx.x; q < n; q += blockDim.x) a += (double)partial[q];
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

}  // namespace stv

// ---------------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------------
using namespace stv;

static size_t fz_align(size_t v) { return (v + 255) & ~(size_t)255; }

static bool fz_use_split(int n);

// `split`: the launch goes to the three-warp kernel (the forward and the backward entry points must agree on it).
static int fz_rows(const stv_photo_cfg* c, bool split) {
    // Strip height: tall strips amortise the 4 warm-up rows; enough strips to fill 148 SMs x ~12 resident warps several times.
    static const int env = getenv("STV_FUSED_ROWS") ? atoi(getenv("STV_FUSED_ROWS")) : 0;  // developer sweep
    if (env > 0) return env < 8 ? 8 : env;   // the workspace / partial buffers are sized for >= 8 rows per strip
    const long long per_row_strips = (long long)c->b*c->S*((c->W + FZ_COLS - 1)/FZ_COLS);
    if (split) {
        // The three-warp kernel also pays a G -> S -> C pipeline fill per strip: it wants taller strips, as long as ~3 waves of the
        // 148 x 5 resident blocks remain (config 3: 96 rows, 0.69 -> 0.65 ms; 192 / 384 rows: no further gain; profiles/r2_photo_split.txt).
        for (int rows : {96, 64, 48})
            if (per_row_strips*((c->H + rows - 1)/rows) >= 2200) return rows;
    }
    int rows = 32;
    while (rows > 8 && per_row_strips*((c->H + rows - 1)/rows) < 148*12*3) rows /= 2;
    return rows;
}

static int fz_check(const stv_photo_cfg* c, const stv_photo_src* src) {
    STV_REQUIRE(c != nullptr && src != nullptr, "stv_photo_fused: cfg / src is NULL");
    STV_REQUIRE(c->b > 0 && c->n > 0 && c->S > 0, "stv_photo_fused: b, n, S must be positive (b=%d n=%d S=%d)", c->b, c->n, c->S);
    STV_REQUIRE(c->S <= STV_MAX_SCALES, "stv_photo_fused: S=%d exceeds STV_MAX_SCALES=%d", c->S, STV_MAX_SCALES);
    STV_REQUIRE(c->n <= STV_FUSED_MAX_SUPPORT, "stv_photo_fused: n=%d support frames (at most %d; use stv_photo_fwd/bwd)", c->n, STV_FUSED_MAX_SUPPORT);
    STV_REQUIRE(c->use_min, "stv_photo_fused: needs min-reprojection (use_min); the mean reduction goes through stv_photo_fwd/bwd");
    STV_REQUIRE(c->H >= 3 && c->W >= 3, "stv_photo_fused: H, W must be >= 3 for reflection padding (H=%d W=%d)", c->H, c->W);
    STV_REQUIRE((long long)c->H*c->W*3 < (1ll << 31), "stv_photo_fused: image too large for 32-bit plane offsets");
    STV_REQUIRE(c->w_ssim >= 0.f && c->w_l1 >= 0.f, "stv_photo_fused: negative loss weights");
    STV_REQUIRE(src->mode == STV_PHOTO_SRC_DEPTH || src->mode == STV_PHOTO_SRC_DISP, "stv_photo_fused: bad source mode %d", src->mode);
    if (src->mode == STV_PHOTO_SRC_DISP) {
        for (int s = 0; s < c->S; ++s)
            STV_REQUIRE(src->h[s] > 0 && src->w[s] > 0, "stv_photo_fused: bad disparity size at scale %d (%d x %d)", s, src->h[s], src->w[s]);
        if (src->min_depth > 0.f || src->max_depth > 0.f) {
            STV_REQUIRE(src->min_depth > 0.f, "Min depth must be greater than 0. (%g)", src->min_depth);
            STV_REQUIRE(!(src->max_depth > 0.f) || src->max_depth >= src->min_depth, "Max depth must be greater than min. (%g vs. %g)",
                        src->max_depth, src->min_depth);
        }
    }
    return STV_OK;
}

static long long fz_strips(const stv_photo_cfg* c, int rows) {
    return (long long)c->b*c->S*((c->W + FZ_COLS - 1)/FZ_COLS)*((c->H + rows - 1)/rows);
}

extern "C" size_t stv_photo_fused_workspace_bytes(const stv_photo_cfg* c) {
    if (c == nullptr || c->b <= 0 || c->H < 3 || c->W < 3 || c->S <= 0) return 0;
    return fz_align((size_t)c->b*c->H*c->W*sizeof(float)) + fz_align((size_t)fz_strips(c, 8)*sizeof(float));
}

extern "C" size_t stv_photo_fused_partial_bytes(const stv_photo_cfg* c) {
    if (c == nullptr || c->b <= 0 || c->H < 3 || c->W < 3 || c->S <= 0 || c->n <= 0) return 0;
    return (size_t)fz_strips(c, 8)*(c->n*12 + FZ_NPART_K)*sizeof(float);
}

extern "C" int stv_tex_create(const float* ptr, long long rows, int W, unsigned long long* handle) {
    STV_REQUIRE(ptr != nullptr && handle != nullptr && rows > 0 && W > 0, "stv_tex_create: bad arguments");
    *handle = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("stv_tex_create: no CUDA device"); return STV_E_CUDA; }
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, dev) != cudaSuccess) { set_error("stv_tex_create: cudaGetDeviceProperties failed"); return STV_E_CUDA; }
    const size_t pitch = (size_t)W*sizeof(float);
    if (((uintptr_t)ptr % pr.textureAlignment) != 0 || (pitch % pr.texturePitchAlignment) != 0 || W > pr.maxTexture2DLinear[0] ||
        rows > pr.maxTexture2DLinear[1] || rows >= (1 << 23))
        return STV_OK;  // not bindable: handle stays 0 and the kernels use plain loads
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypePitch2D;
    rd.res.pitch2D.devPtr = const_cast<float*>(ptr);
    rd.res.pitch2D.desc = cudaCreateChannelDesc<float>();
    rd.res.pitch2D.width = W; rd.res.pitch2D.height = (size_t)rows; rd.res.pitch2D.pitchInBytes = pitch;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t tex = 0;
    if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) { cudaGetLastError(); return STV_OK; }
    *handle = (unsigned long long)tex;
    return STV_OK;
}

extern "C" int stv_tex_destroy(unsigned long long handle) {
    if (handle) cudaDestroyTextureObject((cudaTextureObject_t)handle);
    return STV_OK;
}

// Three-warp variant (gradients on, n <= 2): STV_FUSED_SPLIT=0 falls back to the monolithic sweep (developer A/B switch).
// Blocks per SM are chosen so that no role spills (ptxas: 128 registers at 5 blocks for n = 2; forcing 6-8 blocks spills and is
// 20-70 % slower; n = 3, 4 need more state per warp than 15 resident warps leave and stay on the monolithic kernel:
// profiles/r2_photo_split.txt).
static bool fz_use_split(int n) {
    static const int env = getenv("STV_FUSED_SPLIT") ? atoi(getenv("STV_FUSED_SPLIT")) : 1;
    return env != 0 && n <= FZ_SPLIT_MAX_N;
}

template <int N>
static int fz_launch_split(const FusedParams& p, long long strips, cudaStream_t st) {
    const size_t smem = (size_t)FzSplitLayout<N>::PER_BLOCK*sizeof(float);
    constexpr int MB = 5;
#define FZ_GO(TEX)                                                                                                           \
    do {                                                                                                                     \
        static bool attr = false;                                                                                            \
        if (!attr) { cudaFuncSetAttribute(photo_fused_split_kernel<N, TEX, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; } \
        photo_fused_split_kernel<N, TEX, MB><<<(unsigned)strips, FZ_SPLIT_NT, smem, st>>>(p);                                 \
    } while (0)
    if (p.tex) FZ_GO(true); else FZ_GO(false);
#undef FZ_GO
    count_launch();
    return check_launch("photo_fused_split_kernel");
}

template <int N>
static int fz_launch(const FusedParams& p, long long strips, bool grad, cudaStream_t st) {
    if (N <= FZ_SPLIT_MAX_N && grad && fz_use_split(N)) return fz_launch_split<(N <= FZ_SPLIT_MAX_N ? N : 1)>(p, strips, st);
    const size_t smem = (size_t)FZ_WARPS*FzLayout<N>::PER_WARP*sizeof(float);
    const unsigned blocks = (unsigned)((strips + FZ_WARPS - 1)/FZ_WARPS);
#define FZ_GO(TEX, GRAD)                                                                                                     \
    do {                                                                                                                     \
        static bool attr = false;                                                                                            \
        if (!attr) { cudaFuncSetAttribute(photo_fused_kernel<N, TEX, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; } \
        photo_fused_kernel<N, TEX, GRAD><<<blocks, FZ_NT, smem, st>>>(p);                                                     \
    } while (0)
    if (p.tex) { if (grad) FZ_GO(true, true); else FZ_GO(true, false); }
    else { if (grad) FZ_GO(false, true); else FZ_GO(false, false); }
#undef FZ_GO
    count_launch();
    return check_launch("photo_fused_kernel");
}

extern "C" int stv_photo_fused_fwd(const stv_photo_cfg* c, const stv_photo_src* src, const float* const* maps, const float* tgt,
                                   const float* supp, unsigned long long supp_tex, const float* T, const float* K, const float* Kinv,
                                   const float* noise, unsigned long long* noise_step, float* loss, uint8_t* sel, float* warp0,
                                   float* const* g_unit, float* gpart, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = fz_check(c, src)) return rc;
    STV_REQUIRE(maps && tgt && supp && T && K && Kinv && loss && sel, "stv_photo_fused_fwd: NULL pointer");
    const bool grad = g_unit != nullptr;
    STV_REQUIRE(!grad || gpart != nullptr, "stv_photo_fused_fwd: gradients requested without a partial-sum buffer");
    for (int s = 0; s < c->S; ++s) {
        STV_REQUIRE(maps[s] != nullptr, "stv_photo_fused_fwd: maps[%d] is NULL", s);
        STV_REQUIRE(!grad || g_unit[s] != nullptr, "stv_photo_fused_fwd: g_unit[%d] is NULL", s);
    }
    if (ws == nullptr || ws_bytes < stv_photo_fused_workspace_bytes(c)) {
        set_error("stv_photo_fused_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, stv_photo_fused_workspace_bytes(c));
        return STV_E_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    FusedParams p{};
    p.b = c->b; p.n = c->n; p.S = c->S; p.H = c->H; p.W = c->W; p.w_ssim = c->w_ssim; p.w_l1 = c->w_l1;
    p.use_automask = c->use_automask; p.seed = c->noise_seed; p.step = noise_step;
    p.mode = src->mode;
    for (int s = 0; s < c->S; ++s) {
        p.h[s] = src->mode == STV_PHOTO_SRC_DISP ? src->h[s] : c->H;
        p.w[s] = src->mode == STV_PHOTO_SRC_DISP ? src->w[s] : c->W;
        p.src[s] = maps[s];
        p.g_unit[s] = grad ? g_unit[s] : nullptr;
    }
    p.scaled = (src->min_depth > 0.f || src->max_depth > 0.f) ? 1 : 0;
    if (p.scaled) {
        const float i_max = 1.f/src->min_depth, i_min = src->max_depth > 0.f ? 1.f/src->max_depth : 0.f;
        p.d_mul = i_max - i_min; p.d_add = i_min;
    }
    p.rows = fz_rows(c, grad && c->n <= FZ_SPLIT_MAX_N && fz_use_split(c->n));
    p.nsx = (c->W + FZ_COLS - 1)/FZ_COLS; p.nsy = (c->H + p.rows - 1)/p.rows;
    p.tgt = tgt; p.supp = supp; p.T = T; p.K = K; p.Kinv = Kinv; p.noise = noise;
    float* e0 = (float*)ws;
    p.e0 = e0;
    p.loss_partial = (float*)((char*)ws + fz_align((size_t)c->b*c->H*c->W*sizeof(float)));
    p.gpart = gpart; p.sel = sel; p.warp0 = warp0; p.tex = supp_tex;
    if (c->use_automask) {
        if (int rc = photo_identity_error(c, tgt, supp, e0, st)) return rc;
    }
    const long long strips = fz_strips(c, p.rows);
    int rc;
    switch (c->n) {
        case 1: rc = fz_launch<1>(p, strips, grad, st); break;
        case 2: rc = fz_launch<2>(p, strips, grad, st); break;
        case 3: rc = fz_launch<3>(p, strips, grad, st); break;
        default: rc = fz_launch<4>(p, strips, grad, st); break;
    }
    if (rc) return rc;
    fused_loss_reduce_kernel<<<1, 256, 0, st>>>(p.loss_partial, strips, 1.0/((double)c->S*c->b*c->H*c->W), loss,
                                                (c->use_automask && noise == nullptr && c->noise_seed) ? noise_step : nullptr);
    count_launch();
    return check_launch("fused_loss_reduce_kernel");
}

extern "C" size_t stv_photo_fused_bwd_workspace_bytes(const stv_photo_cfg* c, const stv_photo_src* src) {
    if (c == nullptr || src == nullptr || src->mode != STV_PHOTO_SRC_DISP) return 0;
    size_t total = 0;
    for (int s = 0; s < c->S && s < STV_MAX_SCALES; ++s)
        if (src->h[s] != c->H || src->w[s] != c->W) total += fz_align((size_t)c->b*c->H*src->w[s]*sizeof(float));
    return total;
}

extern "C" int stv_photo_fused_bwd(const stv_photo_cfg* c, const stv_photo_src* src, const float* grad_loss, const float* const* g_unit,
                                   const float* gpart, const float* T, const float* Kinv, float* const* g_maps, float* gT, float* gK,
                                   float* gKinv, void* ws, size_t ws_bytes, void* stream) {
    if (int rc = fz_check(c, src)) return rc;
    STV_REQUIRE(grad_loss && g_unit && gpart && T && Kinv, "stv_photo_fused_bwd: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (gT != nullptr) {
        const int rows = fz_rows(c, c->n <= FZ_SPLIT_MAX_N && fz_use_split(c->n));   // the forward ran with gradients on
        const int spi = (int)(fz_strips(c, rows)/c->b);
        fused_finalize_kernel<<<c->b, 256, 0, st>>>(gpart, spi, c->n, c->b, T, Kinv, grad_loss, gT, gK, gKinv);
        count_launch();
        if (int rc = check_launch("fused_finalize_kernel")) return rc;
    }
    if (g_maps != nullptr) {
        const size_t need = stv_photo_fused_bwd_workspace_bytes(c, src);
        if (need && (ws == nullptr || ws_bytes < need)) { set_error("stv_photo_fused_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need); return STV_E_WORKSPACE; }
        PullParams q{};
        q.b = c->b; q.S = c->S; q.H = c->H; q.W = c->W; q.scale = grad_loss;
        char* wp = (char*)ws;
        for (int s = 0; s < c->S; ++s) {
            STV_REQUIRE(g_unit[s] && g_maps[s], "stv_photo_fused_bwd: g_unit / g_maps[%d] is NULL", s);
            q.h[s] = src->mode == STV_PHOTO_SRC_DISP ? src->h[s] : c->H;
            q.w[s] = src->mode == STV_PHOTO_SRC_DISP ? src->w[s] : c->W;
            q.g_full[s] = g_unit[s]; q.out[s] = g_maps[s];
            if (q.h[s] != c->H || q.w[s] != c->W) { q.tmp[s] = (float*)wp; wp += fz_align((size_t)c->b*c->H*q.w[s]*sizeof(float)); }
        }
        if (int rc = launch_pull(q, st)) return rc;
    }
    return STV_OK;
}

// Backward of stv_disp_to_depth_fwd (the stand-alone up-sampling + to_scaled used when a caller wants the up-sampled maps
// themselves, e.g. the reference's forward_postprocess, src/core/trainer.py:316-321): the same two separable gathers, with the
// depth chain d depth/d disp_up recomputed from the low-resolution disparities inside the first one.
extern "C" size_t stv_disp_to_depth_bwd_workspace_bytes(int b, int h, int w, int H, int W) {
    if (b <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || (h == H && w == W)) return 0;
    return fz_align((size_t)b*H*w*sizeof(float));
}

extern "C" int stv_disp_to_depth_bwd(int b, int h, int w, int H, int W, float min_depth, float max_depth, const float* disp,
                                     const float* g_depth_up, const float* g_disp_up, float* g_disp, void* ws, size_t ws_bytes,
                                     void* stream) {
    STV_REQUIRE(b > 0 && h > 0 && w > 0 && H > 0 && W > 0, "stv_disp_to_depth_bwd: bad shape");
    STV_REQUIRE(b <= 65535 && H <= 65535, "stv_disp_to_depth_bwd: b/H exceed grid limits");
    STV_REQUIRE(disp && g_disp && (g_depth_up || g_disp_up), "stv_disp_to_depth_bwd: NULL pointer");
    const size_t need = stv_disp_to_depth_bwd_workspace_bytes(b, h, w, H, W);
    if (need && (ws == nullptr || ws_bytes < need)) { set_error("stv_disp_to_depth_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need); return STV_E_WORKSPACE; }
    PullParams q{};
    q.b = b; q.S = 1; q.H = H; q.W = W; q.h[0] = h; q.w[0] = w;
    q.g_full[0] = g_depth_up; q.g_full2[0] = g_disp_up; q.disp[0] = disp; q.tmp[0] = (float*)ws; q.out[0] = g_disp;
    q.chain = 1;
    q.scaled = (min_depth > 0.f || max_depth > 0.f) ? 1 : 0;
    if (q.scaled) {
        const float i_max = 1.f/min_depth, i_min = max_depth > 0.f ? 1.f/max_depth : 0.f;
        q.d_mul = i_max - i_min; q.d_add = i_min;
    }
    return launch_pull(q, (cudaStream_t)stream);
}
