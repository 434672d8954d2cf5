// Accumulator read-back and fused epilogue shared by the tcgen05 GEMM and implicit-GEMM convolution kernels.
#pragma once
#include "stv_common.cuh"
#include "stv_tc.cuh"

namespace stv {

__device__ __forceinline__ float act_fwd(int act, float x) {
    switch (act) {
        case STV_ACT_RELU: return fmaxf(x, 0.f);
        case STV_ACT_GELU: return 0.5f*x*(1.f + erff(x*0.70710678118654752f));
        case STV_ACT_ELU: return x > 0.f ? x : __expf(x) - 1.f;  // |abs err| < 2e-7; expm1f costs ~4x the instructions
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
// tcgen05.ld hands every thread ONE ROW (32 consecutive columns); written out like that, each store instruction would touch
// 32 different 128-byte lines. The chunk is therefore transposed through `xs` (this warp's 32 x EPI_LD float staging tile in
// shared memory) so that 8 lanes cover the 128 contiguous bytes of a row and every global access (C, aux, res, dact_src) is
// a full-line, coalesced 128-bit access.
constexpr int EPI_LD = 36;                       // floats per staged row: 16-byte aligned, conflict-free for 128-bit accesses
constexpr int EPI_WARP_FLOATS = 32*EPI_LD;       // staging floats per epilogue warp

__device__ __forceinline__ void epilogue_tile(uint32_t tmem_base, int q, int lane, int m0, int n0, int bn, int M, int N, float* C,
                                              long long ldc, const stv_gemm_epi& e, float* xs) {
    const bool vec = (N & 3) == 0;
    const int rsub = lane >> 3, cq = (lane & 7)*4;
    for (int c = 0; c < bn; c += 32) {
        if (n0 + c >= N) break;  // warp-uniform
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q*32) << 16) + (uint32_t)c, v);
        tc::tmem_ld_wait();
        if (vec) {
            float4* stage = (float4*)(xs + lane*EPI_LD);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                stage[j] = make_float4(__uint_as_float(v[4*j]), __uint_as_float(v[4*j + 1]), __uint_as_float(v[4*j + 2]), __uint_as_float(v[4*j + 3]));
            __syncwarp();
            const int n = n0 + c + cq;
            if (n < N) {
                const float4 bb = e.bias ? __ldg((const float4*)(e.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 gg = e.gamma ? __ldg((const float4*)(e.gamma + n)) : make_float4(1.f, 1.f, 1.f, 1.f);
                const int row0 = m0 + q*32 + rsub;
                // The epilogue's own global reads (residual OR activation-derivative source) are issued for the whole chunk
                // up front: with one resident warp per scheduler a load placed next to its use costs a full memory latency.
                const float* __restrict__ pre_src = e.res ? e.res : e.dact_src;
                float4 pre[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = row0 + 4*i;
                    pre[i] = (pre_src && row < M) ? __ldg((const float4*)(pre_src + (size_t)row*ldc + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = row0 + 4*i;
                    if (row >= M) break;
                    const size_t o = (size_t)row*ldc + n;
                    float4 x = *(const float4*)(xs + (4*i + rsub)*EPI_LD + cq);
                    x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
                    if (e.aux) *(float4*)(e.aux + o) = x;
                    if (e.act) { x.x = act_fwd(e.act, x.x); x.y = act_fwd(e.act, x.y); x.z = act_fwd(e.act, x.z); x.w = act_fwd(e.act, x.w); }
                    x.x *= gg.x; x.y *= gg.y; x.z *= gg.z; x.w *= gg.w;
                    if (e.res) { x.x += pre[i].x; x.y += pre[i].y; x.z += pre[i].z; x.w += pre[i].w; }
                    if (e.dact_src) {
                        const float4 s = e.res ? __ldg((const float4*)(e.dact_src + o)) : pre[i];
                        x.x *= act_bwd(e.dact, s.x); x.y *= act_bwd(e.dact, s.y); x.z *= act_bwd(e.dact, s.z); x.w *= act_bwd(e.dact, s.w);
                    }
                    if (e.accumulate) tc::red_add_v4(C + o, x.x, x.y, x.z, x.w);
                    else *(float4*)(C + o) = x;
                }
            }
            __syncwarp();
        } else {  // narrow outputs (e.g. the 1-channel disparity heads): one row per thread, scalar columns
            const int row = m0 + q*32 + lane;
            if (row >= M) continue;
            const size_t roff = (size_t)row*ldc;
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
