// Disparity post-processing (bilinear upsample fused with disparity -> depth) and the stand-alone ViewSynth kernels.
#include "stv_common.cuh"

namespace stv {

// ATen upsample_bilinear2d(align_corners=False) source taps along one axis (rows 6 of SURVEY 8a).
__device__ __forceinline__ void lin_tap(int dst, float scale, int n_in, int& i0, int& i1, float& lam) {
    const float src = fmaxf(scale*((float)dst + 0.5f) - 0.5f, 0.f);
    i0 = min((int)src, n_in - 1);
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    lam = src - (float)i0;
}

struct DepthScale { float mul, add; int scaled; };

__device__ __forceinline__ float to_depth(float disp, const DepthScale& ds, float& dprime) {
    // to_scaled: disp' = (1/min - 1/max) disp + 1/max  (geometry.py:73-75); to_inv: (x > 0) / clamp(x, eps)  (:89)
    dprime = ds.scaled ? __fadd_rn(__fmul_rn(ds.mul, disp), ds.add) : disp;
    return dprime > 0.f ? 1.0f/fmaxf(dprime, STV_EPS32) : 0.f;
}

__global__ void __launch_bounds__(256) disp_to_depth_fwd_kernel(int h, int w, int H, int W, DepthScale ds,
                                                                const float* __restrict__ disp,
                                                                float* __restrict__ disp_up, float* __restrict__ depth_up) {
    const int X = blockIdx.x*blockDim.x + threadIdx.x, Y = blockIdx.y, i = blockIdx.z;
    if (X >= W) return;
    int y0, y1, x0, x1;
    float ly, lx;
    lin_tap(Y, (float)h/(float)H, h, y0, y1, ly);
    lin_tap(X, (float)w/(float)W, w, x0, x1, lx);
    const float* p = disp + (size_t)i*h*w;
    const float v = (1.f - ly)*((1.f - lx)*__ldg(p + y0*w + x0) + lx*__ldg(p + y0*w + x1)) +
                    ly*((1.f - lx)*__ldg(p + y1*w + x0) + lx*__ldg(p + y1*w + x1));
    const size_t o = (size_t)i*H*W + (size_t)Y*W + X;
    if (disp_up) disp_up[o] = v;
    float dprime;
    depth_up[o] = to_depth(v, ds, dprime);
}

// ---------------------------------------------------------------------------------------------------------------------
// Stand-alone ViewSynth (src/tools/geometry.py:366-391) for arbitrary channel counts. grid = (ceil(HW/256), B)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) view_synth_fwd_kernel(int C, int H, int W, const float* __restrict__ input,
                                                             const float* __restrict__ depth, const float* __restrict__ T,
                                                             const float* __restrict__ K, const float* __restrict__ Kinv,
                                                             float* __restrict__ warp, float* __restrict__ depth_warp,
                                                             uint8_t* __restrict__ mask_valid) {
    const int HW = H*W, pix = blockIdx.x*blockDim.x + threadIdx.x, i = blockIdx.y;
    if (pix >= HW) return;
    const int y = pix/W, x = pix - y*W;
    Cam cam;
    load_cam(cam, T + (size_t)i*16, K + (size_t)i*16, Kinv + (size_t)i*16);
    Proj pr;
    project(cam, (float)x, (float)y, __ldg(depth + (size_t)i*HW + pix), (float)W/(float)(W - 1), (float)H/(float)(H - 1), pr);
    Taps t;
    make_taps(pr.ix, pr.iy, H, W, t);
    for (int c = 0; c < C; ++c) warp[((size_t)i*C + c)*HW + pix] = sample_plane(input + ((size_t)i*C + c)*HW, t);
    if (depth_warp) depth_warp[(size_t)i*HW + pix] = fmaxf(pr.Q[2], STV_EPS32);
    if (mask_valid) {
        // grid = (q/(size-1) - .5)*2 in (-1, 1)  <=>  pixel position in (-0.5, size-0.5)   (geometry.py:388)
        const float gx = (pr.ix + 0.5f)*2.f/(float)W - 1.f, gy = (pr.iy + 0.5f)*2.f/(float)H - 1.f;
        mask_valid[(size_t)i*HW + pix] = (fabsf(gx) < 1.f && fabsf(gy) < 1.f) ? 1 : 0;
    }
}

constexpr int VS_NACC = 27;

__global__ void __launch_bounds__(256) view_synth_bwd_kernel(int C, int H, int W, const float* __restrict__ input,
                                                             const float* __restrict__ depth, const float* __restrict__ T,
                                                             const float* __restrict__ K, const float* __restrict__ Kinv,
                                                             const float* __restrict__ g_warp,
                                                             const float* __restrict__ g_depth_warp,
                                                             float* __restrict__ g_depth, float* __restrict__ partial,
                                                             float* __restrict__ g_input) {
    __shared__ float red[8][VS_NACC];
    const int HW = H*W, pix = blockIdx.x*blockDim.x + threadIdx.x, i = blockIdx.y;
    float acc[VS_NACC];
#pragma unroll
    for (int q = 0; q < VS_NACC; ++q) acc[q] = 0.f;
    if (pix < HW) {
        const int y = pix/W, x = pix - y*W;
        const float sx = (float)W/(float)(W - 1), sy = (float)H/(float)(H - 1);
        Cam cam;
        load_cam(cam, T + (size_t)i*16, K + (size_t)i*16, Kinv + (size_t)i*16);
        Proj pr;
        const float d = __ldg(depth + (size_t)i*HW + pix);
        project(cam, (float)x, (float)y, d, sx, sy, pr);
        Taps t;
        make_taps(pr.ix, pr.iy, H, W, t);
        float gix = 0.f, giy = 0.f;
        if (g_warp) {
            for (int c = 0; c < C; ++c) {
                const float gv = __ldg(g_warp + ((size_t)i*C + c)*HW + pix);
                float dx, dy;
                sample_plane_grad(input + ((size_t)i*C + c)*HW, t, dx, dy);
                gix = fmaf(gv, dx, gix);
                giy = fmaf(gv, dy, giy);
                if (g_input && gv != 0.f) {
                    float* gi = g_input + ((size_t)i*C + c)*HW;
                    atomicAdd(gi + t.o00, gv*(1.f - t.wx)*(1.f - t.wy));
                    atomicAdd(gi + t.o01, gv*t.wx*(1.f - t.wy));
                    atomicAdd(gi + t.o10, gv*(1.f - t.wx)*t.wy);
                    atomicAdd(gi + t.o11, gv*t.wx*t.wy);
                }
            }
        }
        const float gqx = gix*t.gx*sx, gqy = giy*t.gy*sy;
        float gn[3], gQ[3], gP[3], gz = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) gn[r] = fmaf(cam.K0[r], gqx, cam.K1[r]*gqy);
#pragma unroll
        for (int r = 0; r < 3; ++r) { gQ[r] = gn[r]*pr.inv; gz = fmaf(gn[r], pr.Q[r], gz); }
        if (pr.Q[2] >= STV_MIN_Z) gQ[2] -= gz*pr.inv*pr.inv;
        if (g_depth_warp && pr.Q[2] >= STV_EPS32) gQ[2] += __ldg(g_depth_warp + (size_t)i*HW + pix);
#pragma unroll
        for (int r = 0; r < 3; ++r) gP[r] = fmaf(cam.R[r], gQ[0], fmaf(cam.R[3 + r], gQ[1], cam.R[6 + r]*gQ[2]));
        g_depth[(size_t)i*HW + pix] = fmaf(gP[0], pr.ray[0], fmaf(gP[1], pr.ray[1], gP[2]*pr.ray[2]));
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            acc[r*4 + 0] = gQ[r]*pr.P[0]; acc[r*4 + 1] = gQ[r]*pr.P[1]; acc[r*4 + 2] = gQ[r]*pr.P[2]; acc[r*4 + 3] = gQ[r];
            acc[12 + r] = gqx*pr.nrm[r];
            acc[15 + r] = gqy*pr.nrm[r];
            const float gr = gP[r]*d;
            acc[18 + r*3 + 0] = gr*(float)x; acc[18 + r*3 + 1] = gr*(float)y; acc[18 + r*3 + 2] = gr;
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < VS_NACC; ++q) {
        const float v = warp_sum(acc[q]);
        if (lane == 0) red[wid][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < VS_NACC) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        partial[((size_t)i*gridDim.x + blockIdx.x)*VS_NACC + threadIdx.x] = v;
    }
}

__global__ void view_synth_finalize_kernel(const float* __restrict__ partial, int B, int nblk, float* __restrict__ gT,
                                           float* __restrict__ gK, float* __restrict__ gKinv) {
    const int idx = blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= B*48) return;
    const int i = idx/48, e = idx - i*48, which = e/16, m = e & 15, r = m >> 2, c = m & 3;
    float* out = which == 0 ? gT : (which == 1 ? gK : gKinv);
    if (!out) return;
    int off = -1;
    if (which == 0 && r < 3) off = r*4 + c;
    if (which == 1 && r < 2 && c < 3) off = 12 + r*3 + c;
    if (which == 2 && r < 3 && c < 3) off = 18 + r*3 + c;
    double a = 0.0;
    if (off >= 0) for (int t = 0; t < nblk; ++t) a += (double)partial[((size_t)i*nblk + t)*VS_NACC + off];
    out[(size_t)i*16 + m] = (float)a;
}

}  // namespace stv

using namespace stv;

static DepthScale make_scale(float min_depth, float max_depth) {
    DepthScale ds{1.f, 0.f, 0};
    if (min_depth > 0.f || max_depth > 0.f) {
        // trainer.py:48-49: `should_scale = min_depth or max_depth`; to_scaled(min, max) (geometry.py:70-75)
        const double i_max = 1.0/(double)min_depth, i_min = max_depth > 0.f ? 1.0/(double)max_depth : 0.0;
        ds.mul = (float)(i_max - i_min); ds.add = (float)i_min; ds.scaled = 1;
    }
    return ds;
}

extern "C" int stv_disp_to_depth_fwd(int b, int h, int w, int H, int W, float min_depth, float max_depth, const float* disp,
                                     float* disp_up, float* depth_up, void* stream) {
    STV_REQUIRE(b > 0 && h > 0 && w > 0 && H > 0 && W > 0, "stv_disp_to_depth_fwd: bad shape");
    STV_REQUIRE(b <= 65535 && H <= 65535, "stv_disp_to_depth_fwd: b/H exceed grid limits");
    STV_REQUIRE(disp && depth_up, "stv_disp_to_depth_fwd: NULL pointer");
    STV_REQUIRE(!(max_depth > 0.f && !(min_depth > 0.f)), "stv_disp_to_depth_fwd: min_depth must be > 0 when scaling (%g)", min_depth);
    STV_REQUIRE(!(max_depth > 0.f && max_depth < min_depth), "stv_disp_to_depth_fwd: max_depth < min_depth (%g vs %g)", max_depth, min_depth);
    disp_to_depth_fwd_kernel<<<dim3((W + 255)/256, H, b), 256, 0, (cudaStream_t)stream>>>(h, w, H, W, make_scale(min_depth, max_depth),
                                                                                         disp, disp_up, depth_up);
    count_launch();
    return check_launch("disp_to_depth_fwd_kernel");
}

extern "C" int stv_view_synth_fwd(int B, int C, int H, int W, const float* input, const float* depth, const float* T,
                                  const float* K, const float* Kinv, float* warp, float* depth_warp, uint8_t* mask_valid,
                                  void* stream) {
    STV_REQUIRE(B > 0 && C > 0 && H >= 2 && W >= 2, "stv_view_synth_fwd: bad shape");
    STV_REQUIRE(B <= 65535, "stv_view_synth_fwd: B exceeds grid limits");
    STV_REQUIRE(input && depth && T && K && Kinv && warp, "stv_view_synth_fwd: NULL pointer");
    view_synth_fwd_kernel<<<dim3((H*W + 255)/256, B), 256, 0, (cudaStream_t)stream>>>(C, H, W, input, depth, T, K, Kinv, warp,
                                                                                     depth_warp, mask_valid);
    count_launch();
    return check_launch("view_synth_fwd_kernel");
}

extern "C" size_t stv_view_synth_workspace_bytes(int B, int C, int H, int W) {
    (void)C;
    return (size_t)B*((H*W + 255)/256)*VS_NACC*sizeof(float);
}

extern "C" int stv_view_synth_bwd(int B, int C, int H, int W, const float* input, const float* depth, const float* T,
                                  const float* K, const float* Kinv, const float* g_warp, const float* g_depth_warp,
                                  float* g_depth, float* gT, float* gK, float* gKinv, float* g_input, void* ws,
                                  size_t ws_bytes, void* stream) {
    STV_REQUIRE(B > 0 && C > 0 && H >= 2 && W >= 2, "stv_view_synth_bwd: bad shape");
    STV_REQUIRE(B <= 65535, "stv_view_synth_bwd: B exceeds grid limits");
    STV_REQUIRE(input && depth && T && K && Kinv && g_depth && gT, "stv_view_synth_bwd: NULL pointer");
    const size_t need = stv_view_synth_workspace_bytes(B, C, H, W);
    if (!ws || ws_bytes < need) {
        set_error("stv_view_synth_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, need);
        return STV_E_WORKSPACE;
    }
    const int nblk = (H*W + 255)/256;
    view_synth_bwd_kernel<<<dim3(nblk, B), 256, 0, (cudaStream_t)stream>>>(C, H, W, input, depth, T, K, Kinv, g_warp, g_depth_warp,
                                                                          g_depth, (float*)ws, g_input);
    count_launch();
    if (int rc = check_launch("view_synth_bwd_kernel")) return rc;
    view_synth_finalize_kernel<<<(B*48 + 127)/128, 128, 0, (cudaStream_t)stream>>>((const float*)ws, B, nblk, gT, gK, gKinv);
    count_launch();
    return check_launch("view_synth_finalize_kernel");
}
