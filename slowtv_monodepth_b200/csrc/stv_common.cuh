// Shared device helpers for libstv (sm_100a). See include/stv.h for the ABI and DESIGN.md for the kernel inventory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/stv.h"

#define STV_EPS32 1.1920928955078125e-07f  // torch.finfo(float32).eps == reference ops.eps() (src/tools/ops.py:63-66)
#define STV_MIN_Z 0.1f                     // ProjectPoints clamp (src/tools/geometry.py:341)
#define STV_C1 1e-4f                       // SSIM eps1 = 0.01**2 (src/losses/photometric.py:30)
#define STV_C2 9e-4f                       // SSIM eps2 = 0.03**2 (src/losses/photometric.py:31)

#include <cstdlib>

// Reproducible mode (STV_DETERMINISTIC=1, read once per process): every floating-point accumulation has ONE contributor per address
// per launch — no split-K, column reductions on a single row block, the head weight gradient one x-tile at a time — so that two runs
// of the same step give bit-identical gradients. Several times slower; a debugging aid like cuDNN's deterministic switch.
inline bool stv_deterministic() {
    static const bool on = std::getenv("STV_DETERMINISTIC") && std::getenv("STV_DETERMINISTIC")[0] == '1';
    return on;
}

namespace stv {

// ---- host-side error plumbing ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_launch(const char* what);

#define STV_REQUIRE(cond, ...)                     \
    do {                                           \
        if (!(cond)) {                             \
            stv::set_error(__VA_ARGS__);           \
            return STV_E_ARG;                      \
        }                                          \
    } while (0)

// Identity (static) photometric error min_k photo(supp_k, tgt) for the auto-mask -> e0 (b,H,W); defined in stv_photo.cu.
int photo_identity_error(const stv_photo_cfg* c, const float* tgt, const float* supp, float* e0, cudaStream_t st);

// ---- small device helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect_idx(int i, int n) {  // nn.ReflectionPad2d(1): -1 -> 1, n -> n-2
    i = i < 0 ? -i : i;
    return i >= n ? 2*n - 2 - i : i;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of `v`; result valid in thread 0. `red` must hold >= 32 floats. All threads must call.
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = lane < nw ? red[lane] : 0.f;
        v = warp_sum(v);
    }
    return v;
}

// Per-(support frame, image) camera constants. Row-major 4x4 inputs (T: (n,b,4,4), K/Kinv: (b,4,4)).
struct Cam {
    float R[9], t[3];   // top 3x4 of T                                   (src/tools/geometry.py:386)
    float Ki[9];        // Kinv[:3,:3]                                     (src/tools/geometry.py:313)
    float K0[3], K1[3]; // K[0,:3], K[1,:3]                                (src/tools/geometry.py:341)
};

__device__ __forceinline__ void load_cam(Cam& c, const float* __restrict__ T, const float* __restrict__ K,
                                         const float* __restrict__ Kinv) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            c.R[r*3 + j] = __ldg(T + r*4 + j);
            c.Ki[r*3 + j] = __ldg(Kinv + r*4 + j);
        }
        c.t[r] = __ldg(T + r*4 + 3);
        c.K0[r] = __ldg(K + r);
        c.K1[r] = __ldg(K + 4 + r);
    }
}

// Everything the projection of one pixel produces (forward values kept for the backward chain rule).
struct Proj {
    float ray[3], P[3], Q[3], zc, inv, nrm[3], ix, iy;  // ix, iy: un-clamped sample position in pixels
};

// Row 9-11 of SURVEY 8a. sx = W/(W-1), sy = H/(H-1):  ix = qx*sx - 0.5  (geometry.py:347-349 + ATen unnormalize).
__device__ __forceinline__ void project(const Cam& c, float u, float v, float d, float sx, float sy, Proj& p) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        p.ray[r] = fmaf(c.Ki[r*3], u, fmaf(c.Ki[r*3 + 1], v, c.Ki[r*3 + 2]));
        p.P[r] = p.ray[r]*d;
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
        p.Q[r] = fmaf(c.R[r*3], p.P[0], fmaf(c.R[r*3 + 1], p.P[1], fmaf(c.R[r*3 + 2], p.P[2], c.t[r])));
    p.zc = fmaxf(fmaxf(p.Q[2], STV_EPS32), STV_MIN_Z);
    p.inv = 1.0f/p.zc;
#pragma unroll
    for (int r = 0; r < 3; ++r) p.nrm[r] = p.Q[r]*p.inv;
    const float qx = fmaf(c.K0[0], p.nrm[0], fmaf(c.K0[1], p.nrm[1], c.K0[2]*p.nrm[2]));
    const float qy = fmaf(c.K1[0], p.nrm[0], fmaf(c.K1[1], p.nrm[1], c.K1[2]*p.nrm[2]));
    p.ix = fmaf(qx, sx, -0.5f);
    p.iy = fmaf(qy, sy, -0.5f);
}

// Bilinear taps for F.grid_sample(bilinear, border, align_corners=False) at pixel-unit position (ix, iy).
struct Taps {
    int o00, o01, o10, o11;  // offsets into one channel plane
    float wx, wy;            // weights of the +1 taps
    float gx, gy;            // d(clamped)/d(raw): 0 at/beyond the border (ATen clip_coordinates_set_grad), else 1
};

__device__ __forceinline__ void make_taps(float ix, float iy, int H, int W, Taps& t) {
    const float mx = (float)(W - 1), my = (float)(H - 1);
    t.gx = (ix > 0.f && ix < mx) ? 1.f : 0.f;
    t.gy = (iy > 0.f && iy < my) ? 1.f : 0.f;
    ix = fminf(fmaxf(ix, 0.f), mx);  // NaN -> 0
    iy = fminf(fmaxf(iy, 0.f), my);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    t.wx = ix - x0f;
    t.wy = iy - y0f;
    const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);  // an out-of-image +1 tap always has weight 0
    t.o00 = y0*W + x0; t.o01 = y0*W + x1; t.o10 = y1*W + x0; t.o11 = y1*W + x1;
}

__device__ __forceinline__ float sample_plane(const float* __restrict__ p, const Taps& t) {
    const float a = __ldg(p + t.o00), b = __ldg(p + t.o01), c = __ldg(p + t.o10), d = __ldg(p + t.o11);
    const float top = fmaf(t.wx, b - a, a), bot = fmaf(t.wx, d - c, c);
    return fmaf(t.wy, bot - top, top);
}

// Value and d/dix, d/diy (before the border mask) of one plane.
__device__ __forceinline__ float sample_plane_grad(const float* __restrict__ p, const Taps& t, float& dx, float& dy) {
    const float a = __ldg(p + t.o00), b = __ldg(p + t.o01), c = __ldg(p + t.o10), d = __ldg(p + t.o11);
    const float top = fmaf(t.wx, b - a, a), bot = fmaf(t.wx, d - c, c);
    dx = fmaf(t.wy, (d - c) - (b - a), b - a);
    dy = bot - top;
    return fmaf(t.wy, bot - top, top);
}

// SSIM error of one channel from 3x3 window SUMS (not means). (src/losses/photometric.py:40-50)
//   S1 = sum x, S2 = sum x^2, S3 = sum x*y, T1 = sum y, T2 = sum y^2
__device__ __forceinline__ float ssim_err(float S1, float S2, float S3, float T1, float T2) {
    const float k = 1.f/9.f;
    const float mx = S1*k, my = T1*k;
    const float sxx = fmaf(S2, k, -mx*mx), syy = fmaf(T2, k, -my*my), sxy = fmaf(S3, k, -mx*my);
    const float num = (2.f*mx*my + STV_C1)*(2.f*sxy + STV_C2);
    const float den = (mx*mx + my*my + STV_C1)*(sxx + syy + STV_C2);
    const float e = 0.5f*(1.f - num/den);
    return fminf(fmaxf(e, 0.f), 1.f);
}

// d(ssim_err)/d(S1,S2,S3); zero where the clamp is active (torch clamp passes the gradient at equality).
__device__ __forceinline__ void ssim_err_grad(float S1, float S2, float S3, float T1, float T2, float& a, float& b,
                                              float& c) {
    const float k = 1.f/9.f;
    const float mx = S1*k, my = T1*k;
    const float sxx = fmaf(S2, k, -mx*mx), syy = fmaf(T2, k, -my*my), sxy = fmaf(S3, k, -mx*my);
    const float A1 = 2.f*mx*my + STV_C1, A2 = 2.f*sxy + STV_C2;
    const float B1 = mx*mx + my*my + STV_C1, B2 = sxx + syy + STV_C2;
    const float iB1 = 1.f/B1, iB2 = 1.f/B2;
    const float r = A1*A2*iB1*iB2;
    const float e = 0.5f*(1.f - r);
    if (!(e >= 0.f && e <= 1.f)) { a = b = c = 0.f; return; }
    // r = A1 A2 / (B1 B2); with sxx = S2/9 - mx^2 and sxy = S3/9 - mx my:
    //   dA1/dmx = 2 my, dA2/dmx = -2 my, dB1/dmx = 2 mx, dB2/dmx = -2 mx
    const float dr_dmx = (2.f*my*(A2 - A1))*iB1*iB2 - r*(2.f*mx*(B2 - B1))*iB1*iB2;
    const float dr_dS2 = -r*iB2;         // via B2 (d sxx / d(S2/9) = 1)
    const float dr_dS3 = 2.f*A1*iB1*iB2; // via A2 (d sxy / d(S3/9) = 1, factor 2)
    a = -0.5f*k*dr_dmx;
    b = -0.5f*k*dr_dS2;
    c = -0.5f*k*dr_dS3;
}

// Counter-based tie-break noise for the auto-mask (the reference adds eps * randn, src/losses/reconstruction.py:72, only so
// that exact ties between the warped and the static error do not always resolve one way). One 32-bit hash per pixel; the sum
// of its four bytes, centred and scaled to unit variance, is a bell-shaped (Irwin-Hall, n=4) stand-in for N(0,1).
__device__ __forceinline__ float hash_normal(uint64_t seed, uint64_t idx) {
    uint32_t h = (uint32_t)idx*0x9E3779B1u + (uint32_t)(idx >> 32)*0x85EBCA77u + (uint32_t)seed;
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    const int sum = (int)(h & 255u) + (int)((h >> 8) & 255u) + (int)((h >> 16) & 255u) + (int)(h >> 24);
    return ((float)sum - 510.0f)*(1.0f/147.80f);  // var of one byte = (256^2 - 1)/12; four of them -> sigma = 147.80
}

}  // namespace stv
