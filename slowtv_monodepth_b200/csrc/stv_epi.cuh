// Accumulator read-back and fused epilogue shared by the tcgen05 GEMM and implicit-GEMM convolution kernels.
#pragma once
#include "stv_common.cuh"
#include "stv_tc.cuh"

namespace stv {

__device__ __forceinline__ float act_fwd(int act, float x) {
    switch (act) {
        case STV_ACT_RELU: return fmaxf(x, 0.f);
        case STV_ACT_GELU: return 0.5f*x*(1.f + erff(x*0.70710678118654752f));
        case STV_ACT_ELU: return x > 0.f ? x : expm1f(x);
        case STV_ACT_SIGMOID: return 1.f/(1.f + __expf(-x));
        default: return x;
    }
}
// Derivative of the activation. `s` is the saved tensor: the pre-activation for GELU, the OUTPUT for the others.
__device__ __forceinline__ float act_bwd(int act, float s) {
    switch (act) {
        case STV_ACT_RELU: return s > 0.f ? 1.f : 0.f;
        case STV_ACT_GELU: return 0.5f*(1.f + erff(s*0.70710678118654752f)) + s*0.3989422804014327f*__expf(-0.5f*s*s);
        case STV_ACT_ELU: return s > 0.f ? 1.f : s + 1.f;
        case STV_ACT_SIGMOID: return s*(1.f - s);
        default: return 1.f;
    }
}

// One warp drains its 32 TMEM lanes (= 32 output rows) of a 128 x bn fp32 accumulator tile, 32 columns at a time:
//   v = acc + bias[n];  aux = v;  v = act(v);  v *= gamma[n];  v += res;  v *= act'(dact_src);  C = v  or  C += v (red.add).
// q = TMEM lane quarter of the calling warp (warp index & 3). Row r of the tile is output row m0 + r at C + (m0 + r)*ldc.
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_base, int q, int lane, int m0, int n0, int bn, int M, int N, float* C,
                                              long long ldc, const stv_gemm_epi& e) {
    const int row = m0 + q*32 + lane;
    const bool row_ok = row < M;
    const size_t roff = (size_t)row*ldc;
    const bool vec = (N & 3) == 0;
    for (int c = 0; c < bn; c += 32) {
        if (n0 + c >= N) break;  // warp-uniform
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q*32) << 16) + (uint32_t)c, v);
        tc::tmem_ld_wait();
        if (!row_ok) continue;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = n0 + c + 4*j;
                if (n >= N) break;
                float4 r = make_float4(__uint_as_float(v[4*j]), __uint_as_float(v[4*j + 1]), __uint_as_float(v[4*j + 2]),
                                       __uint_as_float(v[4*j + 3]));
                if (e.bias) {
                    const float4 bb = __ldg((const float4*)(e.bias + n));
                    r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
                }
                if (e.aux) *(float4*)(e.aux + roff + n) = r;
                if (e.act) { r.x = act_fwd(e.act, r.x); r.y = act_fwd(e.act, r.y); r.z = act_fwd(e.act, r.z); r.w = act_fwd(e.act, r.w); }
                if (e.gamma) {
                    const float4 g = __ldg((const float4*)(e.gamma + n));
                    r.x *= g.x; r.y *= g.y; r.z *= g.z; r.w *= g.w;
                }
                if (e.res) {
                    const float4 s = __ldg((const float4*)(e.res + roff + n));
                    r.x += s.x; r.y += s.y; r.z += s.z; r.w += s.w;
                }
                if (e.dact_src) {
                    const float4 s = __ldg((const float4*)(e.dact_src + roff + n));
                    r.x *= act_bwd(e.dact, s.x); r.y *= act_bwd(e.dact, s.y); r.z *= act_bwd(e.dact, s.z); r.w *= act_bwd(e.dact, s.w);
                }
                if (e.accumulate) tc::red_add_v4(C + roff + n, r.x, r.y, r.z, r.w);
                else *(float4*)(C + roff + n) = r;
            }
        } else {  // narrow outputs (e.g. the 1-channel disparity heads): scalar columns
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = n0 + c + j;
                if (n >= N) break;
                float r = __uint_as_float(v[j]);
                if (e.bias) r += __ldg(e.bias + n);
                if (e.aux) e.aux[roff + n] = r;
                if (e.act) r = act_fwd(e.act, r);
                if (e.gamma) r *= __ldg(e.gamma + n);
                if (e.res) r += __ldg(e.res + roff + n);
                if (e.dact_src) r *= act_bwd(e.dact, __ldg(e.dact_src + roff + n));
                if (e.accumulate) atomicAdd(C + roff + n, r);
                else C[roff + n] = r;
            }
        }
    }
}

}  // namespace stv
