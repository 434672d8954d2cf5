// Implicit-GEMM convolutions on the tcgen05 tensor cores (kind::tf32, fp32 NHWC activations, filters stored (Cout, R, S, Cin)):
// forward, data gradient and weight gradient of every spatial convolution of the networks — the Monodepth decoder's
// reflect-padded 3x3 convolutions with nearest-x2 upsampling and skip concatenation fused into the operand gather
// (src/networks/decoders/monodepth.py:71-89, utils.py:44-54), the ResNet-18 pose encoder (src/networks/pose.py:40), the
// ConvNeXt stem / down-sampling convolutions (src/networks/depth.py:97) and the pose heads (src/networks/pose.py:75-106).
//
//   fprop   Y[m, k]        = sum_{r,s,c} V[n, p*st + r - pad, q*st + s - pad, c] * Wt[k, (r,s,c)]      m = (n, p, q)
//   dgrad   dV[m, c]       = sum_{r,s,k} dY[n, (y + pad - r)/st, (x + pad - s)/st, k] * Wt[k, (r,s,c)]  m = (n, y, x)
//   wgrad   dWt[k, (r,s,c)] += sum_m dY[m, k] * V[n, p*st + r - pad, q*st + s - pad, c]                 split over m
//
// V is a *virtual* input: cat(up2(src1), src2) along channels, reflect- or zero-padded — never materialised. The im2col
// operand is gathered by four producer warps with 16-byte cp.async copies straight into the swizzled slab layout the tensor
// core reads (K-major slabs for fprop/dgrad, MN-major slabs for wgrad; see stv_tc.cuh); the filter / output-gradient
// operand arrives by TMA; accumulators live in TMEM; the producer warps double as the epilogue warps.
// With reflection padding the data gradient is produced on the PADDED grid (H+2p, W+2p): consumers fold the border back
// (stv_grad_pull), so no atomics are needed.
#include <mutex>

#include "stv_common.cuh"
#include "stv_epi.cuh"
#include "stv_tc.cuh"

namespace stv {

int make_tmap_2d(CUtensorMap* tm, const float* base, long long rows, long long cols, long long ld, int box_rows, int mn_major);

constexpr int CV_BM = 128, CV_BK = 32, CV_THREADS = 192, CV_PRODUCERS = 128, CV_MAX_STAGES = 8, CV_LAG = 2;
constexpr int CV_A_BYTES = CV_BM*CV_BK*4;
constexpr int CV_SLAB_MN = 32*128;

struct Gather {
    const float* p1;
    const float* p2;
    int C, C1, C2, up1;  // virtual tensor channels = C1 + C2; up1: src1 stored at half resolution
    int H, W;            // virtual tensor height / width
    int R, S, stride, pad, reflect;
    int dgrad;           // 0: grid pixels are conv outputs, tensor = conv input; 1: grid pixels are conv inputs, tensor = dY
    int Ktot;            // R*S*C
};

struct ConvParams {
    Gather g;
    int M, N;            // GEMM rows / columns of the output tile grid
    int gridH, gridW;    // pixel grid the gathered rows enumerate (fprop: P,Q; dgrad: input (padded) H,W; wgrad: P,Q)
    long long npix;      // number of grid pixels (N*gridH*gridW)
    int bn, stages, b_mn;
    int kb_total, kb_per_split;
    int taps_kb;         // dgrad: k-blocks per filter tap (ceil(Cout/32)); B row coordinate restarts at every tap
    float* C;
    long long ldc;
    stv_gemm_epi e;
};

struct Tap { int r, s, c; bool ok; };

__device__ __forceinline__ Tap decode_tap(const Gather& g, int kcol) {
    Tap t;
    t.ok = kcol < g.Ktot;
    const int rs = kcol/g.C;
    t.c = kcol - rs*g.C;
    t.r = rs/g.S;
    t.s = rs - t.r*g.S;
    return t;
}

__device__ __forceinline__ int reflect_any(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2*n - 2 - i : i;
}

// Address of the 16-byte chunk (4 channels from t.c) that grid pixel (n, py, px) reads through filter tap t; nullptr = zeros.
__device__ __forceinline__ const float* gather_src(const Gather& g, int n, int py, int px, const Tap& t) {
    if (!t.ok || n < 0) return nullptr;
    int yy, xx;
    if (!g.dgrad) {
        yy = py*g.stride + t.r - g.pad;
        xx = px*g.stride + t.s - g.pad;
        if (g.reflect) {
            yy = reflect_any(yy, g.H);
            xx = reflect_any(xx, g.W);
        } else if ((unsigned)yy >= (unsigned)g.H || (unsigned)xx >= (unsigned)g.W) return nullptr;
    } else {
        const int ty = py + g.pad - t.r, tx = px + g.pad - t.s;
        if (ty < 0 || tx < 0) return nullptr;
        yy = ty/g.stride;
        xx = tx/g.stride;
        if (yy*g.stride != ty || xx*g.stride != tx || yy >= g.H || xx >= g.W) return nullptr;
    }
    if (t.c < g.C1) {
        if (g.up1) return g.p1 + ((size_t)(n*(g.H >> 1) + (yy >> 1))*(g.W >> 1) + (xx >> 1))*g.C1 + t.c;
        return g.p1 + ((size_t)(n*g.H + yy)*g.W + xx)*g.C1 + t.c;
    }
    return g.p2 + ((size_t)(n*g.H + yy)*g.W + xx)*g.C2 + (t.c - g.C1);
}

__device__ __forceinline__ void decode_pixel(long long m, long long npix, int gridH, int gridW, int& n, int& py, int& px) {
    if (m >= npix) { n = -1; py = px = 0; return; }
    const int hw = gridH*gridW;
    n = (int)(m/hw);
    const int rem = (int)(m - (long long)n*hw);
    py = rem/gridW;
    px = rem - py*gridW;
}

struct ConvSmem {
    uint8_t* tiles;
    uint64_t *full, *empty, *tmem_full;
    uint32_t* tmem_slot;
};

__device__ __forceinline__ ConvSmem carve(uint8_t* smem_raw, int stages, int stage_bytes) {
    const uint32_t raw = tc::smem_u32(smem_raw);
    ConvSmem s;
    s.tiles = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    s.full = (uint64_t*)(s.tiles + (size_t)stages*stage_bytes);
    s.empty = s.full + CV_MAX_STAGES;
    s.tmem_full = s.empty + CV_MAX_STAGES;
    s.tmem_slot = (uint32_t*)(s.tmem_full + 1);
    return s;
}

// Producer-side completion: cp.async groups retire in order; once group (it - LAG) has landed, make it visible to the
// async proxy (tcgen05.mma reads shared memory through it) and arrive on that stage's full barrier.
#define CV_PRODUCER_PUBLISH(it_done)                         \
    do {                                                     \
        tc::fence_proxy_async_smem();                        \
        tc::mbar_arrive(&sm.full[(it_done) % p.stages]);     \
    } while (0)

// ---- fprop / dgrad: A = gathered activations (K-major slabs), B = filters by TMA -------------------------------------
__global__ void __launch_bounds__(CV_THREADS) conv_igemm_kernel(const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_bytes = p.bn*CV_BK*4, stage_bytes = CV_A_BYTES + b_bytes;
    const ConvSmem sm = carve(smem_raw, p.stages, stage_bytes);
    const int m0 = blockIdx.x*CV_BM, n0 = blockIdx.y*p.bn;
    const int kb0 = blockIdx.z*p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const int nk = kb1 - kb0;
    const uint32_t tmem_cols = p.bn <= 32 ? 32u : p.bn <= 64 ? 64u : p.bn <= 128 ? 128u : 256u;

    if (warp == 4 && lane == 0) {
        tc::tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) {
            tc::mbar_init(&sm.full[s], CV_PRODUCERS + 1);
            tc::mbar_init(&sm.empty[s], 1);
        }
        tc::mbar_init(sm.tmem_full, 1);
        tc::fence_barrier_init();
    } else if (warp == 5) {
        tc::tmem_alloc(sm.tmem_slot, tmem_cols);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp < 4) {
        // ---- gather producers: thread t copies chunk j = t & 7 of rows (t >> 3) + 16*i, i = 0..7 -------------------------
        const int t = threadIdx.x, j = t & 7, r0 = t >> 3;
        int pn[8], pyx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int n, py, px;
            decode_pixel((long long)m0 + r0 + 16*i, p.npix, p.gridH, p.gridW, n, py, px);
            pn[i] = n;
            pyx[i] = py | (px << 16);
        }
        for (int it = 0; it < nk; ++it) {
            const int s = it % p.stages;
            tc::mbar_wait(&sm.empty[s], ((uint32_t)(it/p.stages) & 1u) ^ 1u);
            int kcol;
            if (!p.g.dgrad) kcol = (kb0 + it)*CV_BK + j*4;
            else {  // reduction index = (tap, k): every tap spans taps_kb blocks of 32 output channels (zero-filled tail)
                const int kb = kb0 + it, tap = kb/p.taps_kb, k = (kb - tap*p.taps_kb)*CV_BK + j*4;
                kcol = k < p.g.C ? tap*p.g.C + k : p.g.Ktot;
            }
            const Tap tp = decode_tap(p.g, kcol);
            const uint32_t a = tc::smem_u32(sm.tiles + (size_t)s*stage_bytes);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + 16*i;
                const float* src = gather_src(p.g, pn[i], pyx[i] & 0xFFFF, pyx[i] >> 16, tp);
                tc::cp_async16(a + row*128 + ((j ^ (row & 7)) << 4), src ? src : p.g.p1, src ? 16u : 0u);
            }
            tc::cp_async_commit();
            if (it >= CV_LAG) {
                tc::cp_async_wait<CV_LAG>();
                CV_PRODUCER_PUBLISH(it - CV_LAG);
            }
        }
        tc::cp_async_wait<0>();
        for (int it = max(nk - CV_LAG, 0); it < nk; ++it) CV_PRODUCER_PUBLISH(it);
        // ---- epilogue ------------------------------------------------------------------------------------------------
        tc::mbar_wait(sm.tmem_full, 0);
        tc::tcgen05_fence_after();
        epilogue_tile(tmem_base, warp, lane, m0, n0, p.bn, p.M, p.N, p.C, p.ldc, p.e);
    } else if (warp == 4) {
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = it % p.stages;
                tc::mbar_wait(&sm.empty[s], ((uint32_t)(it/p.stages) & 1u) ^ 1u);
                tc::mbar_arrive_expect_tx(&sm.full[s], (uint32_t)b_bytes);
                uint8_t* b = sm.tiles + (size_t)s*stage_bytes + CV_A_BYTES;
                const int kb = kb0 + it;
                if (!p.b_mn) tc::tma_load_2d(b, &tmB, &sm.full[s], kb*CV_BK, n0);  // fprop: Wt[n0.., kb*32..] K-major
                else {  // dgrad: rows = output channel k (reduction), columns = tap*C_in + n0 .. (MN-major slabs of 32 columns)
                    const int tap = kb/p.taps_kb, k = (kb - tap*p.taps_kb)*CV_BK;
                    for (int jj = 0; jj < p.bn/32; ++jj) tc::tma_load_2d(b + jj*CV_SLAB_MN, &tmB, &sm.full[s], tap*p.N + n0 + 32*jj, k);
                }
            }
        }
        __syncwarp();
    } else {
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_tf32(CV_BM, p.bn, false, p.b_mn != 0);
            for (int it = 0; it < nk; ++it) {
                const int s = it % p.stages;
                tc::mbar_wait(&sm.full[s], (uint32_t)(it/p.stages) & 1u);
                tc::tcgen05_fence_after();
                const uint32_t a = tc::smem_u32(sm.tiles + (size_t)s*stage_bytes), b = a + CV_A_BYTES;
#pragma unroll
                for (int k8 = 0; k8 < CV_BK/8; ++k8) {
                    const uint64_t da = tc::umma_desc_kmajor(a, k8);
                    const uint64_t db = p.b_mn ? tc::umma_desc_mnmajor(b, k8, CV_SLAB_MN) : tc::umma_desc_kmajor(b, k8);
                    tc::umma_tf32(tmem_base, da, db, idesc, (it > 0 || k8 > 0) ? 1u : 0u);
                }
                tc::umma_commit(&sm.empty[s]);
            }
            tc::umma_commit(sm.tmem_full);
        }
        __syncwarp();
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// ---- wgrad: A = dY by TMA (MN-major: rows = pixels), B = gathered im2col rows (MN-major slabs), reduction over pixels -----
__global__ void __launch_bounds__(CV_THREADS) conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_bytes = p.bn*CV_BK*4, stage_bytes = CV_A_BYTES + b_bytes;
    const ConvSmem sm = carve(smem_raw, p.stages, stage_bytes);
    const int m0 = blockIdx.x*CV_BM, n0 = blockIdx.y*p.bn;  // m0: first output channel, n0: first (r,s,c) column
    const int kb0 = blockIdx.z*p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const int nk = kb1 - kb0;
    const uint32_t tmem_cols = p.bn <= 32 ? 32u : p.bn <= 64 ? 64u : p.bn <= 128 ? 128u : 256u;

    if (warp == 4 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        for (int s = 0; s < p.stages; ++s) {
            tc::mbar_init(&sm.full[s], CV_PRODUCERS + 1);
            tc::mbar_init(&sm.empty[s], 1);
        }
        tc::mbar_init(sm.tmem_full, 1);
        tc::fence_barrier_init();
    } else if (warp == 5) {
        tc::tmem_alloc(sm.tmem_slot, tmem_cols);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    if (warp < 4) {
        // thread t: chunk j = t & 7 of pixel rows (t >> 3) and (t >> 3) + 16, in every 32-column slab of the B tile
        const int t = threadIdx.x, j = t & 7, r0 = t >> 3;
        const int nslab = p.bn/32;
        Tap taps[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) taps[jj] = decode_tap(p.g, jj < nslab ? n0 + jj*32 + j*4 : p.g.Ktot);
        for (int it = 0; it < nk; ++it) {
            const int s = it % p.stages;
            tc::mbar_wait(&sm.empty[s], ((uint32_t)(it/p.stages) & 1u) ^ 1u);
            const uint32_t b = tc::smem_u32(sm.tiles + (size_t)s*stage_bytes + CV_A_BYTES);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = r0 + 16*i;
                int n, py, px;
                decode_pixel((long long)(kb0 + it)*CV_BK + row, p.npix, p.gridH, p.gridW, n, py, px);
                const uint32_t dst = b + row*128 + ((((j >> 1) ^ (row & 3)) << 5) | ((j & 1) << 4));
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    if (jj < nslab) {
                        const float* src = gather_src(p.g, n, py, px, taps[jj]);
                        tc::cp_async16(dst + jj*CV_SLAB_MN, src ? src : p.g.p1, src ? 16u : 0u);
                    }
                }
            }
            tc::cp_async_commit();
            if (it >= CV_LAG) {
                tc::cp_async_wait<CV_LAG>();
                CV_PRODUCER_PUBLISH(it - CV_LAG);
            }
        }
        tc::cp_async_wait<0>();
        for (int it = max(nk - CV_LAG, 0); it < nk; ++it) CV_PRODUCER_PUBLISH(it);
        tc::mbar_wait(sm.tmem_full, 0);
        tc::tcgen05_fence_after();
        epilogue_tile(tmem_base, warp, lane, m0, n0, p.bn, p.M, p.N, p.C, p.ldc, p.e);
    } else if (warp == 4) {
        if (lane == 0) {
            for (int it = 0; it < nk; ++it) {
                const int s = it % p.stages;
                tc::mbar_wait(&sm.empty[s], ((uint32_t)(it/p.stages) & 1u) ^ 1u);
                tc::mbar_arrive_expect_tx(&sm.full[s], (uint32_t)CV_A_BYTES);
                uint8_t* a = sm.tiles + (size_t)s*stage_bytes;
                for (int jj = 0; jj < CV_BM/32; ++jj) tc::tma_load_2d(a + jj*CV_SLAB_MN, &tmA, &sm.full[s], m0 + 32*jj, (kb0 + it)*CV_BK);
            }
        }
        __syncwarp();
    } else {
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_tf32(CV_BM, p.bn, true, true);
            for (int it = 0; it < nk; ++it) {
                const int s = it % p.stages;
                tc::mbar_wait(&sm.full[s], (uint32_t)(it/p.stages) & 1u);
                tc::tcgen05_fence_after();
                const uint32_t a = tc::smem_u32(sm.tiles + (size_t)s*stage_bytes), b = a + CV_A_BYTES;
#pragma unroll
                for (int k8 = 0; k8 < CV_BK/8; ++k8)
                    tc::umma_tf32(tmem_base, tc::umma_desc_mnmajor(a, k8, CV_SLAB_MN), tc::umma_desc_mnmajor(b, k8, CV_SLAB_MN), idesc,
                                  (it > 0 || k8 > 0) ? 1u : 0u);
                tc::umma_commit(&sm.empty[s]);
            }
            tc::umma_commit(sm.tmem_full);
        }
        __syncwarp();
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 5) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// ---- host side ------------------------------------------------------------------------------------------------------
static int pick_bn_conv(int N) {
    int best = 32, best_cost = 1 << 30;
    for (int bn = 256; bn >= 32; bn -= 32) {
        const int tiles = (N + bn - 1)/bn, cost = tiles*bn;
        if (cost < best_cost) { best = bn; best_cost = cost; }
    }
    return best;
}

static int conv_stages(int bn, int nk) {
    const int stage_bytes = CV_A_BYTES + bn*CV_BK*4;
    const int budget = bn <= 128 ? 100*1024 : 200*1024;
    int st = budget/stage_bytes;
    st = st > CV_MAX_STAGES ? CV_MAX_STAGES : st;
    if (st > nk) st = nk;
    return st < CV_LAG + 1 ? CV_LAG + 1 : st;
}

static size_t conv_smem(int bn, int stages) {
    return (size_t)stages*(CV_A_BYTES + bn*CV_BK*4) + 1024 + (2*CV_MAX_STAGES + 1)*8 + 16;
}

static int set_smem_attr(const void* fn, const char* what) {
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024);
    if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute failed (%s)", what, cudaGetErrorString(e)); return STV_E_CUDA; }
    return STV_OK;
}

static int check_geom(const stv_conv_geom* g, const char* what, int& P, int& Q) {
    STV_REQUIRE(g, "%s: null geometry", what);
    STV_REQUIRE(g->N > 0 && g->H > 0 && g->W > 0 && g->C1 > 0 && g->C2 >= 0 && g->Cout > 0, "%s: empty tensor", what);
    STV_REQUIRE(g->C1 % 4 == 0 && g->C2 % 4 == 0, "%s: input channels must be multiples of 4 (C1=%d, C2=%d)", what, g->C1, g->C2);
    STV_REQUIRE(g->R > 0 && g->S > 0 && g->stride > 0 && g->pad >= 0, "%s: bad filter geometry", what);
    STV_REQUIRE(!g->up1 || (g->H % 2 == 0 && g->W % 2 == 0), "%s: x2 nearest upsampling needs even H, W", what);
    STV_REQUIRE(!g->reflect || (g->pad < g->H && g->pad < g->W), "%s: reflection padding larger than the image", what);
    STV_REQUIRE(g->H < 32768 && g->W < 32768, "%s: image too large", what);
    P = (g->H + 2*g->pad - g->R)/g->stride + 1;
    Q = (g->W + 2*g->pad - g->S)/g->stride + 1;
    STV_REQUIRE(P > 0 && Q > 0, "%s: empty output", what);
    return STV_OK;
}

static Gather make_gather(const stv_conv_geom* g, const float* s1, const float* s2) {
    Gather G = {};
    G.p1 = s1; G.p2 = s2 ? s2 : s1;
    G.C1 = g->C1; G.C2 = g->C2; G.C = g->C1 + g->C2; G.up1 = g->up1;
    G.H = g->H; G.W = g->W; G.R = g->R; G.S = g->S; G.stride = g->stride; G.pad = g->pad; G.reflect = g->reflect;
    G.dgrad = 0; G.Ktot = g->R*g->S*G.C;
    return G;
}

}  // namespace stv

using namespace stv;

extern "C" int stv_conv_fprop(const stv_conv_geom* g, const float* src1, const float* src2, const float* w, float* y,
                              const stv_gemm_epi* epi, void* stream) {
    int P, Q;
    if (int rc = check_geom(g, "stv_conv_fprop", P, Q)) return rc;
    STV_REQUIRE(src1 && w && y && (g->C2 == 0 || src2), "stv_conv_fprop: null pointer");
    ConvParams p = {};
    p.g = make_gather(g, src1, src2);
    p.npix = (long long)g->N*P*Q;
    STV_REQUIRE(p.npix < (1ll << 31), "stv_conv_fprop: too many output pixels");
    p.M = (int)p.npix; p.N = g->Cout; p.gridH = P; p.gridW = Q;
    p.bn = pick_bn_conv(p.N); p.b_mn = 0;
    p.kb_total = (p.g.Ktot + CV_BK - 1)/CV_BK; p.kb_per_split = p.kb_total; p.taps_kb = 1;
    p.stages = conv_stages(p.bn, p.kb_total);
    p.C = y; p.ldc = g->Cout;
    if (epi) p.e = *epi;
    STV_REQUIRE(!p.e.accumulate, "stv_conv_fprop: accumulate is not supported");
    CUtensorMap tmB;
    if (int rc = make_tmap_2d(&tmB, w, g->Cout, p.g.Ktot, p.g.Ktot, p.bn, 0)) return rc;
    static std::once_flag once; static int attr_rc = 0;
    std::call_once(once, [] { attr_rc = set_smem_attr((const void*)conv_igemm_kernel, "stv_conv_fprop"); });
    if (attr_rc) return attr_rc;
    const dim3 grid((p.M + CV_BM - 1)/CV_BM, (p.N + p.bn - 1)/p.bn, 1);
    conv_igemm_kernel<<<grid, CV_THREADS, conv_smem(p.bn, p.stages), (cudaStream_t)stream>>>(tmB, p);
    count_launch();
    return check_launch("stv_conv_fprop");
}

extern "C" int stv_conv_dgrad(const stv_conv_geom* g, const float* dy, const float* w, float* dv, const stv_gemm_epi* epi, void* stream) {
    int P, Q;
    if (int rc = check_geom(g, "stv_conv_dgrad", P, Q)) return rc;
    STV_REQUIRE(dy && w && dv, "stv_conv_dgrad: null pointer");
    STV_REQUIRE(g->Cout % 4 == 0, "stv_conv_dgrad: Cout must be a multiple of 4 (got %d)", g->Cout);
    const int Cin = g->C1 + g->C2;
    // Gathered tensor = dY (N, P, Q, Cout); grid pixels = (padded, when reflecting) input pixels.
    ConvParams p = {};
    Gather& G = p.g;
    G.p1 = G.p2 = dy; G.C = G.C1 = g->Cout; G.C2 = 0; G.up1 = 0;
    G.H = P; G.W = Q; G.R = g->R; G.S = g->S; G.stride = g->stride; G.reflect = 0; G.dgrad = 1;
    G.pad = g->reflect ? 0 : g->pad;
    G.Ktot = g->R*g->S*g->Cout;
    p.gridH = g->reflect ? g->H + 2*g->pad : g->H;
    p.gridW = g->reflect ? g->W + 2*g->pad : g->W;
    p.npix = (long long)g->N*p.gridH*p.gridW;
    STV_REQUIRE(p.npix < (1ll << 31), "stv_conv_dgrad: too many pixels");
    p.M = (int)p.npix; p.N = Cin;
    p.bn = pick_bn_conv(p.N); p.b_mn = 1;
    p.taps_kb = (g->Cout + CV_BK - 1)/CV_BK;
    p.kb_total = g->R*g->S*p.taps_kb; p.kb_per_split = p.kb_total;
    p.stages = conv_stages(p.bn, p.kb_total);
    p.C = dv; p.ldc = Cin;
    if (epi) p.e = *epi;
    STV_REQUIRE(!p.e.accumulate, "stv_conv_dgrad: accumulate is not supported");
    CUtensorMap tmB;  // filters as a (Cout) x (R*S*Cin) matrix, read in MN-major boxes {32 columns, 32 rows}
    if (int rc = make_tmap_2d(&tmB, w, g->Cout, (long long)g->R*g->S*Cin, (long long)g->R*g->S*Cin, 32, 1)) return rc;
    static std::once_flag once; static int attr_rc = 0;
    std::call_once(once, [] { attr_rc = set_smem_attr((const void*)conv_igemm_kernel, "stv_conv_dgrad"); });
    if (attr_rc) return attr_rc;
    const dim3 grid((p.M + CV_BM - 1)/CV_BM, (p.N + p.bn - 1)/p.bn, 1);
    conv_igemm_kernel<<<grid, CV_THREADS, conv_smem(p.bn, p.stages), (cudaStream_t)stream>>>(tmB, p);
    count_launch();
    return check_launch("stv_conv_dgrad");
}

extern "C" int stv_conv_wgrad(const stv_conv_geom* g, const float* src1, const float* src2, const float* dy, float* dw, int split_k,
                              void* stream) {
    int P, Q;
    if (int rc = check_geom(g, "stv_conv_wgrad", P, Q)) return rc;
    STV_REQUIRE(src1 && dy && dw && (g->C2 == 0 || src2), "stv_conv_wgrad: null pointer");
    STV_REQUIRE(g->Cout % 4 == 0, "stv_conv_wgrad: Cout must be a multiple of 4 (got %d)", g->Cout);
    ConvParams p = {};
    p.g = make_gather(g, src1, src2);
    p.npix = (long long)g->N*P*Q;
    STV_REQUIRE(p.npix < (1ll << 31), "stv_conv_wgrad: too many output pixels");
    p.M = g->Cout; p.N = p.g.Ktot; p.gridH = P; p.gridW = Q;
    p.bn = pick_bn_conv(p.N); p.b_mn = 1; p.taps_kb = 1;
    p.kb_total = (int)((p.npix + CV_BK - 1)/CV_BK);
    if (split_k <= 0) {  // ~2 waves of CTAs, at least 8 k-blocks each
        const int tiles = ((p.M + CV_BM - 1)/CV_BM)*((p.N + p.bn - 1)/p.bn);
        split_k = (2*148 + tiles - 1)/tiles;
        if (split_k > p.kb_total/8) split_k = p.kb_total/8;
        if (split_k < 1) split_k = 1;
    }
    split_k = split_k < p.kb_total ? split_k : p.kb_total;
    p.kb_per_split = (p.kb_total + split_k - 1)/split_k;
    split_k = (p.kb_total + p.kb_per_split - 1)/p.kb_per_split;
    p.stages = conv_stages(p.bn, p.kb_per_split);
    p.C = dw; p.ldc = p.g.Ktot;
    p.e.accumulate = 1;
    CUtensorMap tmA;  // dY as a (pixels) x (Cout) matrix, MN-major boxes {32 channels, 32 pixels}
    if (int rc = make_tmap_2d(&tmA, dy, p.npix, g->Cout, g->Cout, 32, 1)) return rc;
    static std::once_flag once; static int attr_rc = 0;
    std::call_once(once, [] { attr_rc = set_smem_attr((const void*)conv_wgrad_kernel, "stv_conv_wgrad"); });
    if (attr_rc) return attr_rc;
    const dim3 grid((p.M + CV_BM - 1)/CV_BM, (p.N + p.bn - 1)/p.bn, split_k);
    conv_wgrad_kernel<<<grid, CV_THREADS, conv_smem(p.bn, p.stages), (cudaStream_t)stream>>>(tmA, p);
    count_launch();
    return check_launch("stv_conv_wgrad");
}
