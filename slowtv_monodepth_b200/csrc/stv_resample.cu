// Separable-position bilinear resampling of (P, H, W) fp32 planes -> (P, oh, ow): the two image operations of the reference's
// aspect-ratio augmentation (src/core/aspect_ratio.py:36-186), which runs on the GPU inside training_step (trainer.py:106)
// over 2*(1+n)*b images per step:
//   mode STV_RESAMPLE_GRID   sample position  ix = ax*j + bx  (pixel units), zero padding outside the image — what
//                            kornia.center_crop -> warp_affine -> F.affine_grid + F.grid_sample(bilinear, zeros,
//                            align_corners=False) evaluates (crop_aug, aspect_ratio.py:69-97); ax/bx are host-computed.
//   mode STV_RESAMPLE_INTERP F.interpolate(bilinear, align_corners=False) (resize_aug, aspect_ratio.py:129-167):
//                            ix = max(ax*(j + 0.5) - 0.5, 0), ax = in/out; the +1 tap is clamped to the last pixel.
// HBM-bound by construction: every output is written once, every input texel is read once from HBM (the <= 4 taps of
// neighbouring outputs hit L1/L2); a thread produces 4 horizontally adjacent outputs and stores them as one 16-byte word.
#include "stv_common.cuh"

namespace stv {

struct Tap { int i0, i1; float w0, w1; };  // value = w0*src[i0] + w1*src[i1]

__device__ __forceinline__ Tap make_tap(int j, float a, float b, int n, int mode) {
    Tap t;
    if (mode == STV_RESAMPLE_INTERP) {
        const float s = fmaxf(a*((float)j + 0.5f) - 0.5f, 0.f);   // ATen area_pixel_compute_source_index
        const int i0 = min((int)s, n - 1);
        t.i0 = i0; t.i1 = min(i0 + 1, n - 1);
        t.w1 = s - (float)i0; t.w0 = 1.f - t.w1;
    } else {
        const float s = fmaf(a, (float)j, b);
        const float f = floorf(s);
        const int i0 = (int)f;
        const float w1 = s - f, w0 = 1.f - w1;
        const bool in0 = i0 >= 0 && i0 < n, in1 = i0 + 1 >= 0 && i0 + 1 < n;
        t.i0 = in0 ? i0 : 0; t.i1 = in1 ? i0 + 1 : 0;
        t.w0 = in0 ? w0 : 0.f; t.w1 = in1 ? w1 : 0.f;            // zeros padding
    }
    return t;
}

__global__ void __launch_bounds__(256) resample_kernel(int H, int W, int oh, int ow, float ax, float bx, float ay, float by, int mode,
                                                       const float* __restrict__ src, float* __restrict__ dst) {
    const int xq = blockIdx.x*blockDim.x + threadIdx.x;   // group of 4 output columns
    const int y = blockIdx.y;
    const size_t plane = blockIdx.z;
    const int x0 = xq*4;
    if (x0 >= ow) return;
    const Tap ty = make_tap(y, ay, by, H, mode);
    const float* r0 = src + (plane*H + ty.i0)*W;
    const float* r1 = src + (plane*H + ty.i1)*W;
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const Tap tx = make_tap(min(x0 + u, ow - 1), ax, bx, W, mode);
        const float top = tx.w0*__ldg(r0 + tx.i0) + tx.w1*__ldg(r0 + tx.i1);
        const float bot = tx.w0*__ldg(r1 + tx.i0) + tx.w1*__ldg(r1 + tx.i1);
        v[u] = ty.w0*top + ty.w1*bot;
    }
    float* o = dst + (plane*oh + y)*ow + x0;
    if (x0 + 3 < ow && (((uintptr_t)o) & 15) == 0) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    else
        for (int u = 0; u < 4 && x0 + u < ow; ++u) o[u] = v[u];
}

}  // namespace stv

using namespace stv;

extern "C" int stv_resample_bilinear(long long P, int H, int W, int oh, int ow, float ax, float bx, float ay, float by, int mode,
                                     const float* src, float* dst, void* stream) {
    STV_REQUIRE(P > 0 && H > 0 && W > 0 && oh > 0 && ow > 0, "stv_resample_bilinear: bad shape (P=%lld H=%d W=%d oh=%d ow=%d)", P, H, W, oh, ow);
    STV_REQUIRE(P <= 65535 && oh <= 65535, "stv_resample_bilinear: too many planes / rows for one launch (P=%lld oh=%d)", P, oh);
    STV_REQUIRE(mode == STV_RESAMPLE_GRID || mode == STV_RESAMPLE_INTERP, "stv_resample_bilinear: unknown mode %d", mode);
    STV_REQUIRE(src && dst, "stv_resample_bilinear: NULL pointer");
    const int groups = (ow + 3)/4;
    dim3 grid((groups + 255)/256, oh, (unsigned)P);
    resample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(H, W, oh, ow, ax, bx, ay, by, mode, src, dst);
    count_launch();
    return check_launch("resample_kernel");
}
