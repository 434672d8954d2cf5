// Packed fp32x2 arithmetic for sm_100a (PTX add/mul/fma .f32x2 -> SASS FADD2/FMUL2/FFMA2): two fp32 lanes per issue slot.
// The fused loss kernels are issue-bound, not HBM-bound, once the warp is fused in, so halving the FP instruction count is
// what moves them toward the HBM roofline.
#pragma once
#include <cuda_runtime.h>

namespace stv {

struct f2 {
    unsigned long long v;
};

__device__ __forceinline__ f2 mk2(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f2 splat2(float a) { return mk2(a, a); }
__device__ __forceinline__ float lo2(f2 a) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
    return x;
}
__device__ __forceinline__ float hi2(f2 a) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
    return y;
}
__device__ __forceinline__ f2 operator+(f2 a, f2 b) {
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 operator*(f2 a, f2 b) {
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ f2 ld2(const float* p) {  // 8-byte aligned shared/global load as a packed pair
    const float2 t = *reinterpret_cast<const float2*>(p);
    return mk2(t.x, t.y);
}
__device__ __forceinline__ float rcp_fast(float x) {  // MUFU.RCP, <= 1 ulp; callers guarantee |x| is far from denormal
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

}  // namespace stv
