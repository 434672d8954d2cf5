// Accumulator read-back and fused epilogue shared by the tcgen05 GEMM and implicit-GEMM convolution kernels.
#pragma once
#include "stv_common.cuh"
#include "stv_tc.cuh"
#include "stv_f2.cuh"

namespace stv {

// Branch-free erf (Abramowitz & Stegun 7.1.26, |abs err| <= 1.5e-7) given e = exp(-u*u): one bare MUFU.RCP (rcp.approx, <= 1 ulp;
// __frcp_rn adds a Newton fix-up and a special-case branch, ~7 more instructions per element for accuracy the polynomial lacks) + 5 FMA. erff() costs ~3x
// the instructions and branches on |u|, which serialises the four elements of a float4 in the epilogue warps.
__device__ __forceinline__ float erf_fast(float u, float e) {
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, fabsf(u), 1.f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    return copysignf(fmaf(-p*t, e, 1.f), u);
}

__device__ __forceinline__ float act_fwd(int act, float x) {
    switch (act) {
        case STV_ACT_RELU: return fmaxf(x, 0.f);
        case STV_ACT_GELU: {  // exact (erf) GELU, torch.nn.functional.gelu default
            const float u = x*0.70710678118654752f;
            return 0.5f*x*(1.f + erf_fast(u, __expf(-u*u)));
        }
        case STV_ACT_ELU: return x > 0.f ? x : __expf(x) - 1.f;  // |abs err| < 2e-7; expm1f costs ~4x the instructions
        case STV_ACT_SIGMOID: return 1.f/(1.f + __expf(-x));
        default: return x;
    }
}
// Derivative of the activation. `s` is the saved tensor: the pre-activation for GELU, the OUTPUT for the others.
__device__ __forceinline__ float act_bwd(int act, float s) {
    switch (act) {
        case STV_ACT_RELU: return s > 0.f ? 1.f : 0.f;
        case STV_ACT_GELU: {  // Phi(s) + s*phi(s); exp(-s^2/2) is shared by the erf and the density term
            const float u = s*0.70710678118654752f, e = __expf(-u*u);
            return fmaf(s*0.3989422804014327f, e, 0.5f*(1.f + erf_fast(u, e)));
        }
        case STV_ACT_ELU: return s > 0.f ? 1.f : s + 1.f;
        case STV_ACT_SIGMOID: return s*(1.f - s);
        default: return 1.f;
    }
}

// ---- packed (two elements per instruction) GELU / GELU' for the float4 epilogue ------------------------------------------------
// The GELU epilogues are issue-bound (fc1 / GELU'-dgrad of the ConvNeXt MLP: ~2x the time of their plain siblings at equal
// bytes and flops), so the polynomial, the products and the final blend run as FFMA2 / FMUL2 on element pairs; only the two
// MUFU ops (rcp, ex2) and the sign/abs bit operations stay per element. Same A&S 7.1.26 formulation as erf_fast.
__device__ __forceinline__ f2 abs2(f2 a) { f2 r; r.v = a.v & 0x7FFFFFFF7FFFFFFFull; return r; }
__device__ __forceinline__ f2 copysign2(f2 mag, f2 sgn) { f2 r; r.v = mag.v | (sgn.v & 0x8000000080000000ull); return r; }  // mag >= 0
__device__ __forceinline__ float ex2_fast(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// erf(u) for u = x/sqrt(2) and e = exp(-u^2), both packed.
__device__ __forceinline__ void erf_exp2(f2 x, f2& erf, f2& e) {
    const f2 u = x*splat2(0.70710678118654752f);
    const f2 d = fma2(splat2(0.3275911f), abs2(u), splat2(1.f));
    const f2 t = mk2(rcp_fast(lo2(d)), rcp_fast(hi2(d)));
    const f2 a = (u*u)*splat2(-1.4426950408889634f);                      // -u^2 * log2(e)
    e = mk2(ex2_fast(lo2(a)), ex2_fast(hi2(a)));
    f2 p = fma2(splat2(-1.061405429f), t, splat2(1.453152027f));          // negated polynomial: r = 1 - poly(t) * e
    p = fma2(p, t, splat2(-1.421413741f));
    p = fma2(p, t, splat2(0.284496736f));
    p = fma2(p, t, splat2(-0.254829592f));
    erf = copysign2(fma2(p*t, e, splat2(1.f)), u);
}
__device__ __forceinline__ f2 gelu2(f2 x) {
    f2 erf, e;
    erf_exp2(x, erf, e);
    const f2 hx = x*splat2(0.5f);
    return fma2(hx, erf, hx);
}
__device__ __forceinline__ f2 gelu_grad2(f2 s) {  // Phi(s) + s*phi(s)
    f2 erf, e;
    erf_exp2(s, erf, e);
    return fma2(s*splat2(0.3989422804014327f), e, fma2(splat2(0.5f), erf, splat2(0.5f)));
}
__device__ __forceinline__ void gelu4(float4& x) {
    const f2 a = gelu2(mk2(x.x, x.y)), b = gelu2(mk2(x.z, x.w));
    x = make_float4(lo2(a), hi2(a), lo2(b), hi2(b));
}
__device__ __forceinline__ void mul_gelu_grad4(float4& x, const float4 s) {
    const f2 a = mk2(x.x, x.y)*gelu_grad2(mk2(s.x, s.y)), b = mk2(x.z, x.w)*gelu_grad2(mk2(s.z, s.w));
    x = make_float4(lo2(a), hi2(a), lo2(b), hi2(b));
}

// Output row -> element offset. Plain GEMM / convolution outputs are row-major (row*ldc). A stride-s data gradient is computed
// as s*s stride-1 sub-problems, one per output parity (a, b): row (n, y', x') of a sub-problem lands at pixel (n, s*y'+a, s*x'+b).
// remap == 2 (row-segment convolution, stv_conv3.cu): GEMM rows enumerate (image row n*H + y, segment, 128 pixels); pixels past
// the end of the image row do not exist (`ROW_NONE`).
constexpr size_t ROW_NONE = ~(size_t)0;
struct RowMap {
    long long ldc;
    int remap, gH, gW, oH, oW, ost, oa, ob;
    __device__ __forceinline__ size_t off(int row) const {
        if (!remap) return (size_t)row*ldc;
        if (remap == 2) {   // gW = segments per image row, oW = pixels per image row
            const int tile = row >> 7, i = row & 127, line = tile/gW, x = (tile - line*gW)*128 + i;
            return x < oW ? ((size_t)line*oW + x)*ldc : ROW_NONE;
        }
        const int hw = gH*gW, n = row/hw, rem = row - n*hw, y = rem/gW, x = rem - y*gW;
        return ((size_t)(n*oH + y*ost + oa)*oW + (x*ost + ob))*ldc;
    }
    // every one of the 128 rows starting at m0 exists
    __device__ __forceinline__ bool tile_full(int m0, int M) const {
        if (m0 + 128 > M) return false;
        if (remap != 2) return true;
        const int tile = m0 >> 7, line = tile/gW;
        return (tile - line*gW)*128 + 128 <= oW;
    }
};

constexpr int EPI_LD = 36;                       // floats per staged row: 16-byte aligned, conflict-free for 128-bit accesses
constexpr int EPI_WARP_FLOATS = 32*EPI_LD;       // staging floats per epilogue warp

// The 8 rows a lane owns in one transposed 32 x 32 chunk: epilogue chain + store, returns the lane's column sums.
// Every template argument is a compile-time copy of a kernel-uniform descriptor field (-1 = read it at run time): the epilogue
// warps are issue-bound on the narrow-K layers, and per-element branches on descriptor fields were ~25 of their ~54
// instructions per output element. FULL = all 128 rows of the tile exist (no per-row bound check).
template <int ACT, int DACT, int AUX, int GAM, int RES, int ACC, int CS, bool FULL>
__device__ __forceinline__ float4 epilogue_rows(const stv_gemm_epi& e, float* __restrict__ C, const float* xs, const size_t (&roffs)[8],
                                                const float4 (&pre)[8], const float4 bb, const float4 gg, int row0, int M, int n, int rsub,
                                                int cq) {
    const int act = ACT >= 0 ? ACT : e.act, dact = DACT >= 0 ? DACT : e.dact;
    const bool aux = AUX >= 0 ? AUX != 0 : e.aux != nullptr, gam = GAM >= 0 ? GAM != 0 : e.gamma != nullptr;
    const bool res = RES >= 0 ? RES != 0 : e.res != nullptr, accm = ACC >= 0 ? ACC != 0 : e.accumulate != 0;
    const bool cs_on = CS >= 0 ? CS != 0 : e.colsum != nullptr;
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (!FULL && roffs[i] == ROW_NONE) continue;
        const size_t o = roffs[i] + n;
        float4 x = *(const float4*)(xs + (4*i + rsub)*EPI_LD + cq);
        x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
        if (aux) *(float4*)(e.aux + o) = x;
        if (ACT == STV_ACT_GELU) gelu4(x);
        else if (act) { x.x = act_fwd(act, x.x); x.y = act_fwd(act, x.y); x.z = act_fwd(act, x.z); x.w = act_fwd(act, x.w); }
        if (gam) { x.x *= gg.x; x.y *= gg.y; x.z *= gg.z; x.w *= gg.w; }
        if (res) { x.x += pre[i].x; x.y += pre[i].y; x.z += pre[i].z; x.w += pre[i].w; }
        if (DACT != STV_ACT_NONE && e.dact_src) {
            const float4 s = res ? __ldg((const float4*)(e.dact_src + o)) : pre[i];
            if (DACT == STV_ACT_GELU) mul_gelu_grad4(x, s);
            else { x.x *= act_bwd(dact, s.x); x.y *= act_bwd(dact, s.y); x.z *= act_bwd(dact, s.z); x.w *= act_bwd(dact, s.w); }
        }
        if (accm) tc::red_add_v4(C + o, x.x, x.y, x.z, x.w);
        else *(float4*)(C + o) = x;
        if (cs_on) { cs.x += x.x; cs.y += x.y; cs.z += x.z; cs.w += x.w; }
    }
    return cs;
}

// Chunks c_first, c_first + c_step, ... of the tile are handled by this warp (two warps per lane quarter split the columns).
__device__ __forceinline__ void epilogue_tile(uint32_t tmem_base, int q, int lane, int m0, int n0, int bn, int M, int N, float* C,
                                              const RowMap& rm, const stv_gemm_epi& e, float* xs, int c_first = 0, int c_step = 32) {
    const bool vec = (N & 3) == 0;
    const int rsub = lane >> 3, cq = (lane & 7)*4;
    for (int c = c_first; c < bn; c += c_step) {
        if (n0 + c >= N) break;  // warp-uniform
        if (vec) {
            // The chunk's own global reads (bias, layer scale, residual OR activation-derivative source) are requested BEFORE the
            // accumulator is read back: with one or two resident warps per scheduler nothing else hides their latency, and they
            // then overlap with the TMEM read + the transposition through shared memory instead of following them.
            const int n = n0 + c + cq;
            const bool col_ok = n < N;
            const int row0 = m0 + q*32 + rsub;
            size_t roffs[8];
            float4 pre[8];
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f), gg = make_float4(1.f, 1.f, 1.f, 1.f);
            if (col_ok) {
                if (e.bias) bb = __ldg((const float4*)(e.bias + n));
                if (e.gamma) gg = __ldg((const float4*)(e.gamma + n));
#pragma unroll
                for (int i = 0; i < 8; ++i) roffs[i] = row0 + 4*i < M ? rm.off(row0 + 4*i) : ROW_NONE;
                const float* __restrict__ pre_src = e.res ? e.res : e.dact_src;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    pre[i] = (pre_src && roffs[i] != ROW_NONE) ? __ldg((const float4*)(pre_src + roffs[i] + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            uint32_t v[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(q*32) << 16) + (uint32_t)c, v);
            tc::tmem_ld_wait();
            float4* stage = (float4*)(xs + lane*EPI_LD);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                stage[j] = make_float4(__uint_as_float(v[4*j]), __uint_as_float(v[4*j + 1]), __uint_as_float(v[4*j + 2]), __uint_as_float(v[4*j + 3]));
            __syncwarp();
            if (col_ok) {
                // Descriptor fields are kernel-uniform: dispatch once per chunk to a loop specialised on the combinations the networks
                // use (fc1, fc2, GELU' dgrad, plain store, split-K accumulate, conv + ELU / ReLU); anything else takes the generic loop.
                float4 cs;
                const bool full = rm.tile_full(m0, M);
                const bool plain = !e.aux && !e.gamma && !e.res && !e.dact_src && !e.accumulate && !e.colsum;
#define STV_EPI_ROWS(...) cs = full ? epilogue_rows<__VA_ARGS__, true>(e, C, xs, roffs, pre, bb, gg, row0, M, n, rsub, cq) \
                               : epilogue_rows<__VA_ARGS__, false>(e, C, xs, roffs, pre, bb, gg, row0, M, n, rsub, cq)
                if (plain && e.act == STV_ACT_NONE) STV_EPI_ROWS(STV_ACT_NONE, STV_ACT_NONE, 0, 0, 0, 0, 0);
                else if (plain && e.act == STV_ACT_ELU) STV_EPI_ROWS(STV_ACT_ELU, STV_ACT_NONE, 0, 0, 0, 0, 0);
                else if (plain && e.act == STV_ACT_RELU) STV_EPI_ROWS(STV_ACT_RELU, STV_ACT_NONE, 0, 0, 0, 0, 0);
                else if (e.act == STV_ACT_GELU && e.aux && !e.gamma && !e.res && !e.dact_src && !e.accumulate && !e.colsum)
                    STV_EPI_ROWS(STV_ACT_GELU, STV_ACT_NONE, 1, 0, 0, 0, 0);                                   // fc1 + bias + GELU, z saved
                else if (e.act == STV_ACT_NONE && !e.aux && e.gamma && e.res && !e.dact_src && !e.accumulate && !e.colsum)
                    STV_EPI_ROWS(STV_ACT_NONE, STV_ACT_NONE, 0, 1, 1, 0, 0);                                   // fc2 + bias + layer-scale + residual
                else if (e.act == STV_ACT_NONE && !e.aux && !e.gamma && !e.res && e.dact_src && e.dact == STV_ACT_GELU && !e.accumulate)
                    STV_EPI_ROWS(STV_ACT_NONE, STV_ACT_GELU, 0, 0, 0, 0, -1);                                  // (g W2) * GELU'(z) [+ column sums]
                else if (e.act == STV_ACT_NONE && !e.aux && !e.gamma && !e.res && !e.dact_src && e.accumulate && !e.colsum)
                    STV_EPI_ROWS(STV_ACT_NONE, STV_ACT_NONE, 0, 0, 0, 1, 0);                                   // split-K weight gradients
                else STV_EPI_ROWS(-1, -1, -1, -1, -1, -1, -1);
#undef STV_EPI_ROWS
                if (e.colsum) {  // lanes l, l^8, l^16, l^24 hold the same 4 columns (different rows): combine, then one red per column group
                    // (all 32 lanes of the warp reach this point together when n < N for the whole warp; guard with the active mask)
                    const unsigned am = __activemask();
#pragma unroll
                    for (int o = 8; o <= 16; o <<= 1) {
                        cs.x += __shfl_xor_sync(am, cs.x, o); cs.y += __shfl_xor_sync(am, cs.y, o);
                        cs.z += __shfl_xor_sync(am, cs.z, o); cs.w += __shfl_xor_sync(am, cs.w, o);
                    }
                    if (rsub == 0) tc::red_add_v4(e.colsum + n, cs.x, cs.y, cs.z, cs.w);
                }
            }
            __syncwarp();
        } else {  // narrow outputs (e.g. the 1-channel disparity heads): one row per thread, scalar columns
            uint32_t v[32];
            tc::tmem_ld32(tmem_base + ((uint32_t)(q*32) << 16) + (uint32_t)c, v);
            tc::tmem_ld_wait();
            const int row = m0 + q*32 + lane;
            if (row >= M) continue;
            const size_t roff = rm.off(row);
            if (roff == ROW_NONE) continue;
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
