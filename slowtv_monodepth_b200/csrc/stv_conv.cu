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
#include "stv_gemm.cuh"
#include "stv_tc.cuh"

namespace stv {

constexpr int CV_BM = 128, CV_BK = 32, CV_THREADS = 192, CV_PRODUCERS = 128, CV_MAX_STAGES = 8;
constexpr int CV_A_BYTES = CV_BM*CV_BK*4;
constexpr int CV_SLAB_MN = 32*128;

struct Gather {
    const float* p1;
    const float* p2;
    int C, C1, C2, up1;  // virtual tensor channels = C1 + C2; up1: src1 stored at half resolution
    int H, W;            // virtual tensor height / width
    int R, S, stride, pad, reflect;
    int dgrad;           // 0: grid pixels are conv outputs, tensor = conv input; 1: grid pixels are conv inputs, tensor = dY
    int sshift, smask;   // log2(stride), stride - 1 (strides are powers of two)
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

// Filter tap + first channel of one 16-byte chunk of the reduction / column axis. ok = false: beyond the axis (zero fill).
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
// The same chunk one k-block (32 columns) further along the (r, s, c) axis — no divisions in the main loop.
__device__ __forceinline__ void advance_tap(const Gather& g, Tap& t) {
    t.c += 32;
    while (t.c >= g.C) {
        t.c -= g.C;
        if (++t.s == g.S) { t.s = 0; ++t.r; }
    }
    t.ok = t.r < g.R;
}

__device__ __forceinline__ int reflect_any(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2*n - 2 - i : i;
}

// Address of the 16-byte chunk (4 channels from t.c) that one grid pixel reads through filter tap t; nullptr = zeros.
// (n, y0, x0) describe the pixel: n = image index (-1: row beyond the tensor); fprop / wgrad: y0 = py*stride - pad, the
// input row of tap 0; dgrad: y0 = py + pad, so that the output row of tap r is (y0 - r)/stride when that division is exact.
__device__ __forceinline__ const float* gather_src(const Gather& g, int n, int y0, int x0, const Tap& t) {
    if (!t.ok || n < 0) return nullptr;
    int yy, xx;
    if (!g.dgrad) {
        yy = y0 + t.r;
        xx = x0 + t.s;
        if (g.reflect) {
            yy = reflect_any(yy, g.H);
            xx = reflect_any(xx, g.W);
        } else if ((unsigned)yy >= (unsigned)g.H || (unsigned)xx >= (unsigned)g.W) return nullptr;
    } else {
        const int ty = y0 - t.r, tx = x0 - t.s;
        if ((ty | tx) < 0 || ((ty | tx) & g.smask)) return nullptr;
        yy = ty >> g.sshift;
        xx = tx >> g.sshift;
        if (yy >= g.H || xx >= g.W) return nullptr;
    }
    if (t.c < g.C1) {
        const int pix = g.up1 ? (n*(g.H >> 1) + (yy >> 1))*(g.W >> 1) + (xx >> 1) : (n*g.H + yy)*g.W + xx;
        return g.p1 + (size_t)pix*g.C1 + t.c;
    }
    return g.p2 + (size_t)((n*g.H + yy)*g.W + xx)*g.C2 + (t.c - g.C1);
}

__device__ __forceinline__ void decode_pixel(long long m, long long npix, int gridH, int gridW, int& n, int& py, int& px) {
    if (m >= npix) { n = -1; py = px = 0; return; }
    const unsigned hw = (unsigned)(gridH*gridW), mm = (unsigned)m;  // npix < 2^31 (checked on the host): 32-bit divisions
    n = (int)(mm/hw);
    const unsigned rem = mm - (unsigned)n*hw;
    py = (int)(rem/(unsigned)gridW);
    px = (int)(rem - (unsigned)py*(unsigned)gridW);
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
        const int lag = p.stages - 1;  // cp.async groups kept in flight per thread (a stage is published `lag` iterations later)
        int pn[8], py0[8], px0[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int n, py, px;
            decode_pixel((long long)m0 + r0 + 16*i, p.npix, p.gridH, p.gridW, n, py, px);
            pn[i] = n;
            py0[i] = p.g.dgrad ? py + p.g.pad : py*p.g.stride - p.g.pad;
            px0[i] = p.g.dgrad ? px + p.g.pad : px*p.g.stride - p.g.pad;
        }
        // This thread's chunk of the reduction axis, advanced incrementally (no divisions in the loop).
        Tap tp;
        int kc = 0;  // dgrad: first output channel of the current k-block within its tap
        if (!p.g.dgrad) tp = decode_tap(p.g, kb0*CV_BK + j*4);
        else {
            const int tap = kb0/p.taps_kb;
            kc = (kb0 - tap*p.taps_kb)*CV_BK;
            tp.r = tap/p.g.S; tp.s = tap - tp.r*p.g.S; tp.c = kc + j*4; tp.ok = tp.c < p.g.C && tp.r < p.g.R;
        }
        for (int it = 0; it < nk; ++it) {
            const int s = it % p.stages;
            tc::mbar_wait(&sm.empty[s], ((uint32_t)(it/p.stages) & 1u) ^ 1u);
            const uint32_t a = tc::smem_u32(sm.tiles + (size_t)s*stage_bytes);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = r0 + 16*i;
                const float* src = gather_src(p.g, pn[i], py0[i], px0[i], tp);
                tc::cp_async16(a + row*128 + ((j ^ (row & 7)) << 4), src ? src : p.g.p1, src ? 16u : 0u);
            }
            tc::cp_async_commit();
            if (!p.g.dgrad) advance_tap(p.g, tp);
            else {  // reduction index = (tap, k): every tap spans taps_kb blocks of 32 output channels (zero-filled tail)
                kc += CV_BK;
                if (kc >= p.taps_kb*CV_BK) { kc = 0; if (++tp.s == p.g.S) { tp.s = 0; ++tp.r; } }
                tp.c = kc + j*4; tp.ok = tp.c < p.g.C && tp.r < p.g.R;
            }
            if (it >= lag) {
                tc::cp_async_wait_dyn(lag);
                CV_PRODUCER_PUBLISH(it - lag);
            }
        }
        tc::cp_async_wait<0>();
        for (int it = max(nk - lag, 0); it < nk; ++it) CV_PRODUCER_PUBLISH(it);
        // ---- epilogue ------------------------------------------------------------------------------------------------
        tc::mbar_wait(sm.tmem_full, 0);
        tc::tcgen05_fence_after();
        epilogue_tile(tmem_base, warp, lane, m0, n0, p.bn, p.M, p.N, p.C, RowMap{p.ldc, 0, 0, 0, 0, 0, 0, 0, 0}, p.e, (float*)sm.tiles + warp*EPI_WARP_FLOATS);
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
        const int lag = p.stages - 1;
        const int nslab = p.bn/32;
        Tap taps[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) taps[jj] = decode_tap(p.g, jj < nslab ? n0 + jj*32 + j*4 : p.g.Ktot);
        // The two pixel rows of this thread, advanced by 32 pixels per k-block (no divisions in the loop).
        int pn[2], py[2], px[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) decode_pixel((long long)kb0*CV_BK + r0 + 16*i, p.npix, p.gridH, p.gridW, pn[i], py[i], px[i]);
        const int nimg = (int)(p.npix/(p.gridH*p.gridW));
        for (int it = 0; it < nk; ++it) {
            const int s = it % p.stages;
            tc::mbar_wait(&sm.empty[s], ((uint32_t)(it/p.stages) & 1u) ^ 1u);
            const uint32_t b = tc::smem_u32(sm.tiles + (size_t)s*stage_bytes + CV_A_BYTES);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = r0 + 16*i;
                const int y0 = py[i]*p.g.stride - p.g.pad, x0 = px[i]*p.g.stride - p.g.pad;
                const uint32_t dst = b + row*128 + ((((j >> 1) ^ (row & 3)) << 5) | ((j & 1) << 4));
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    if (jj < nslab) {
                        const float* src = gather_src(p.g, pn[i], y0, x0, taps[jj]);
                        tc::cp_async16(dst + jj*CV_SLAB_MN, src ? src : p.g.p1, src ? 16u : 0u);
                    }
                }
                if (pn[i] >= 0) {
                    px[i] += CV_BK;
                    while (px[i] >= p.gridW) {
                        px[i] -= p.gridW;
                        if (++py[i] == p.gridH) { py[i] = 0; ++pn[i]; }
                    }
                    if (pn[i] >= nimg) pn[i] = -1;
                }
            }
            tc::cp_async_commit();
            if (it >= lag) {
                tc::cp_async_wait_dyn(lag);
                CV_PRODUCER_PUBLISH(it - lag);
            }
        }
        tc::cp_async_wait<0>();
        for (int it = max(nk - lag, 0); it < nk; ++it) CV_PRODUCER_PUBLISH(it);
        tc::mbar_wait(sm.tmem_full, 0);
        tc::tcgen05_fence_after();
        epilogue_tile(tmem_base, warp, lane, m0, n0, p.bn, p.M, p.N, p.C, RowMap{p.ldc, 0, 0, 0, 0, 0, 0, 0, 0}, p.e, (float*)sm.tiles + warp*EPI_WARP_FLOATS);
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
// Ring depth: two resident CTAs per SM (110 KB each) when that still leaves >= 4 stages, otherwise one CTA with up to 200 KB.
// The gather producers keep stages-1 cp.async groups in flight per thread, so depth is what hides the L2/HBM latency.
static int conv_stages(int bn, int nk) {
    const int stage_bytes = CV_A_BYTES + bn*CV_BK*4;
    int st = (110*1024)/stage_bytes;
    if (st < 4) st = (200*1024)/stage_bytes;
    st = st > CV_MAX_STAGES ? CV_MAX_STAGES : st;
    if (st > nk) st = nk;
    return st < 2 ? 2 : st;
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
    STV_REQUIRE((g->stride & (g->stride - 1)) == 0, "%s: stride must be a power of two (got %d)", what, g->stride);
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
    G.smask = g->stride - 1; G.sshift = 0;
    while ((1 << G.sshift) < g->stride) ++G.sshift;
    return G;
}

// ---- TMA im2col front-ends: the whole operand gather is done by the TMA unit (stv_gemm.cu, ConvOperand) ---------------------------
// Eligible: one real source tensor, zero padding, channels a multiple of 32 (a k-block is then one filter tap x 32 channels).
// Everything else (the 3/6-channel stems, 16-channel decoder level 0, strided data gradients, and — when the caller does not
// materialise it with stv_vpad — the virtual upsample/concat/reflect input) goes through the cp.async gather kernels above.
static bool tma_eligible(const stv_conv_geom* g) {
    return g->C2 == 0 && !g->up1 && !g->reflect && g->C1 % 32 == 0 && g->stride <= 8 && g->pad <= 120 && g->R <= 120 && g->S <= 120;
}

static int fprop_tma(const stv_conv_geom* g, int P, int Q, const float* x, const float* w, float* y, const stv_gemm_epi* epi, cudaStream_t st) {
    GemmParams p = {};
    const int Ktot = g->R*g->S*g->C1;
    if (conv3_eligible(g->N, P, Q, g->C1, g->Cout, g->R, g->S, g->stride))
        return conv3_launch(x, g->N, g->H, g->W, g->C1, P, Q, -g->pad, -g->pad, g->R, g->S, w, g->Cout, 0, 0, 0, g->Cout, Ktot, y, g->Cout, epi, st,
                            "stv_conv_fprop(row segments)");
    p.M = g->N*P*Q; p.N = g->Cout; p.K = Ktot;
    p.bn = pick_bn(p.N, (p.M + GEMM_BM - 1)/GEMM_BM);
    p.kb_total = Ktot/GEMM_BK; p.kb_per_split = p.kb_total;
    p.C = y; p.ldc = g->Cout;
    if (epi) p.e = *epi;
    ConvOperand& cv = p.cv;
    cv.mode = 1; cv.gridH = P; cv.gridW = Q; cv.stride = g->stride; cv.lw = cv.lh = -g->pad;
    cv.R = g->R; cv.S = g->S; cv.C = g->C1; cv.cblocks = g->C1/GEMM_BK; cv.flip = 0;
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap_im2col(&tmA, x, g->N, g->H, g->W, g->C1, -g->pad, -g->pad, g->pad - (g->S - 1), g->pad - (g->R - 1), g->stride, GEMM_BM, 0)) return rc;
    if (int rc = make_tmap_2d(&tmB, w, g->Cout, Ktot, Ktot, p.bn, 0)) return rc;
    p.pair_B = w; p.pair_ldb = Ktot;   // the CTA-pair kernel re-encodes the filters with half-tile boxes
    return launch_gemm(tmA, tmB, p, 1, st, "stv_conv_fprop(tma)");
}

static int dgrad_tma(const stv_conv_geom* g, int P, int Q, const float* dy, const float* w, float* dx, const stv_gemm_epi* epi, cudaStream_t st) {
    GemmParams p = {};
    const int Cin = g->C1 + g->C2;
    if (conv3_eligible(g->N, g->H, g->W, g->Cout, Cin, g->R, g->S, 1))
        return conv3_launch(dy, g->N, P, Q, g->Cout, g->H, g->W, g->pad - (g->S - 1), g->pad - (g->R - 1), g->R, g->S, w, Cin, 1, 1, Cin, g->Cout,
                            (long long)g->R*g->S*Cin, dx, Cin, epi, st, "stv_conv_dgrad(row segments)");
    p.M = g->N*g->H*g->W; p.N = Cin; p.K = g->R*g->S*g->Cout;
    p.bn = pick_bn(p.N, (p.M + GEMM_BM - 1)/GEMM_BM);
    p.b_mn = 1;
    ConvOperand& cv = p.cv;
    cv.mode = 1; cv.gridH = g->H; cv.gridW = g->W; cv.stride = 1; cv.lw = g->pad - (g->S - 1); cv.lh = g->pad - (g->R - 1);
    cv.R = g->R; cv.S = g->S; cv.C = g->Cout; cv.cblocks = (g->Cout + GEMM_BK - 1)/GEMM_BK; cv.flip = 1; cv.b_tap_cols = Cin;
    cv.r0 = cv.s0 = 0; cv.tstep = 1; cv.Sfull = g->S;
    p.kb_total = g->R*g->S*cv.cblocks; p.kb_per_split = p.kb_total;
    p.C = dx; p.ldc = Cin;
    if (epi) p.e = *epi;
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap_im2col(&tmA, dy, g->N, P, Q, g->Cout, cv.lw, cv.lh, cv.lw + (g->W - Q), cv.lh + (g->H - P), 1, GEMM_BM, 0)) return rc;
    if (int rc = make_tmap_2d(&tmB, w, g->Cout, (long long)g->R*g->S*Cin, (long long)g->R*g->S*Cin, 32, 1)) return rc;
    p.pair_B = w;   // MN-major filters: the 32-column slab boxes serve the CTA-pair kernel as they are
    return launch_gemm(tmA, tmB, p, 1, st, "stv_conv_dgrad(tma)");
}

// Data gradient of a stride-s convolution as s*s stride-1 problems: output pixels of parity (a, b) only see the filter taps
// r = (a + pad) mod s + s*i, s = (b + pad) mod s + s*j, read from dY at (y' + q_a - i, x' + q_b - j), q_a = (a + pad - r_a)/s.
static int dgrad_tma_strided(const stv_conv_geom* g, int P, int Q, const float* dy, const float* w, float* dx, const stv_gemm_epi* epi,
                             cudaStream_t st) {
    const int Cin = g->C1 + g->C2, s = g->stride, Hs = g->H/s, Ws = g->W/s;
    bool all = true;
    for (int a = 0; a < s; ++a) all = all && ((a + g->pad) % s < g->R) && ((a + g->pad) % s < g->S);
    if (!all && cudaMemsetAsync(dx, 0, (size_t)g->N*g->H*g->W*Cin*sizeof(float), st) != cudaSuccess) return check_launch("stv_conv_dgrad(memset)");
    CUtensorMap tmB;
    if (int rc = make_tmap_2d(&tmB, w, g->Cout, (long long)g->R*g->S*Cin, (long long)g->R*g->S*Cin, 32, 1)) return rc;
    for (int a = 0; a < s; ++a)
        for (int b = 0; b < s; ++b) {
            const int ra = (a + g->pad) % s, sb = (b + g->pad) % s;
            if (ra >= g->R || sb >= g->S) continue;
            const int na = (g->R - ra + s - 1)/s, nb = (g->S - sb + s - 1)/s, qa = (a + g->pad - ra)/s, qb = (b + g->pad - sb)/s;
            GemmParams p = {};
            p.M = g->N*Hs*Ws; p.N = Cin; p.K = na*nb*g->Cout;
            p.bn = pick_bn(p.N, (p.M + GEMM_BM - 1)/GEMM_BM);
            p.b_mn = 1;
            ConvOperand& cv = p.cv;
            cv.mode = 1; cv.gridH = Hs; cv.gridW = Ws; cv.stride = 1; cv.lw = qb - (nb - 1); cv.lh = qa - (na - 1);
            cv.R = na; cv.S = nb; cv.C = g->Cout; cv.cblocks = (g->Cout + GEMM_BK - 1)/GEMM_BK; cv.flip = 1; cv.b_tap_cols = Cin;
            cv.r0 = ra; cv.s0 = sb; cv.tstep = s; cv.Sfull = g->S;
            p.kb_total = na*nb*cv.cblocks; p.kb_per_split = p.kb_total;
            p.C = dx; p.ldc = Cin;
            p.remap = 1; p.oH = g->H; p.oW = g->W; p.ost = s; p.oa = a; p.ob = b;
            if (epi) p.e = *epi;
            CUtensorMap tmA;
            if (int rc = make_tmap_im2col(&tmA, dy, g->N, P, Q, g->Cout, cv.lw, cv.lh, cv.lw + (Ws - Q), cv.lh + (Hs - P), 1, GEMM_BM, 0)) return rc;
            p.pair_B = w;
            if (int rc = launch_gemm(tmA, tmB, p, 1, st, "stv_conv_dgrad(tma, strided)")) return rc;
        }
    return STV_OK;
}

static int wgrad_tma(const stv_conv_geom* g, int P, int Q, const float* x, const float* dy, float* dw, int split_k, cudaStream_t st) {
    GemmParams p = {};
    const int Ktot = g->R*g->S*g->C1;
    const long long npix = (long long)g->N*P*Q;
    p.M = g->Cout; p.N = Ktot; p.K = (int)npix;
    p.bn = pick_bn(p.N, 1 << 20);
    p.a_mn = 1; p.b_mn = 1;
    p.kb_total = (int)((npix + GEMM_BK - 1)/GEMM_BK);
    if (stv_deterministic()) split_k = 1;
    if (split_k <= 0) {  // at most two full waves of CTAs, at least 8 k-blocks each
        const int tiles = ((p.M + GEMM_BM - 1)/GEMM_BM)*((p.N + p.bn - 1)/p.bn);
        split_k = (2*148)/tiles;
        if (split_k > p.kb_total/8) split_k = p.kb_total/8;
        if (split_k < 1) split_k = 1;
    }
    split_k = split_k < p.kb_total ? split_k : p.kb_total;
    p.kb_per_split = (p.kb_total + split_k - 1)/split_k;
    split_k = (p.kb_total + p.kb_per_split - 1)/p.kb_per_split;
    p.C = dw; p.ldc = Ktot; p.e.accumulate = 1;
    ConvOperand& cv = p.cv;
    cv.mode = 2; cv.gridH = P; cv.gridW = Q; cv.stride = g->stride; cv.lw = cv.lh = -g->pad;
    cv.R = g->R; cv.S = g->S; cv.C = g->C1;
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap_2d(&tmA, dy, npix, g->Cout, g->Cout, 32, 1)) return rc;
    if (int rc = make_tmap_im2col(&tmB, x, g->N, g->H, g->W, g->C1, -g->pad, -g->pad, g->pad - (g->S - 1), g->pad - (g->R - 1), g->stride, GEMM_BK, 1)) return rc;
    return launch_gemm(tmA, tmB, p, split_k, st, "stv_conv_wgrad(tma)");
}

// out (N, H+2p, W+2p, Cp) = reflection-pad_p(cat(up2(src1) | src1, src2)), channels [C1+C2, Cp) zero; p = g.pad when g.reflect else 0.
__global__ void vpad_kernel(int N, int H, int W, int C1, int C2, int Cp, int up1, int pad, const float* __restrict__ s1,
                            const float* __restrict__ s2, float* __restrict__ out) {
    const int Cc = C1 + C2, c4n = Cp >> 2, Hp = H + 2*pad, Wp = W + 2*pad;
    const long long total = (long long)N*Hp*Wp*c4n;
    for (unsigned idx = blockIdx.x*blockDim.x + threadIdx.x; idx < (unsigned)total; idx += gridDim.x*blockDim.x) {  // 32-bit index math (count < 2^31, checked on the host)
        const int c = (int)(idx % c4n)*4;
        unsigned r = idx/c4n;
        const int xp = (int)(r % Wp); r /= Wp;
        const int yp = (int)(r % Hp);
        const int n = (int)(r/Hp);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < Cc) {
            const int y = reflect_any(yp - pad, H), x = reflect_any(xp - pad, W);
            const float* src = c < C1 ? (up1 ? s1 + ((size_t)(n*(H >> 1) + (y >> 1))*(W >> 1) + (x >> 1))*C1 + c : s1 + ((size_t)(n*H + y)*W + x)*C1 + c)
                                      : s2 + ((size_t)(n*H + y)*W + x)*C2 + (c - C1);
            v = __ldg((const float4*)src);
        }
        ((float4*)out)[idx] = v;
    }
}

}  // namespace stv

using namespace stv;

extern "C" int stv_vpad(const stv_conv_geom* g, const float* src1, const float* src2, int Cp, float* out, void* stream) {
    int P, Q;
    if (int rc = check_geom(g, "stv_vpad", P, Q)) return rc;
    STV_REQUIRE(src1 && out && (g->C2 == 0 || src2), "stv_vpad: null pointer");
    STV_REQUIRE(Cp >= g->C1 + g->C2 && Cp % 4 == 0, "stv_vpad: padded channel count %d must be a multiple of 4 and >= %d", Cp, g->C1 + g->C2);
    const int pad = g->reflect ? g->pad : 0;
    const long long total = (long long)g->N*(g->H + 2*pad)*(g->W + 2*pad)*(Cp/4);
    STV_REQUIRE(total < (1ll << 31), "stv_vpad: tensor too large");
    const int blocks = (int)((total + 255)/256 < 148ll*32 ? (total + 255)/256 : 148ll*32);
    vpad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g->N, g->H, g->W, g->C1, g->C2, Cp, g->up1, pad, src1, src2 ? src2 : src1, out);
    count_launch();
    return check_launch("stv_vpad");
}

extern "C" int stv_conv_fprop(const stv_conv_geom* g, const float* src1, const float* src2, const float* w, float* y,
                              const stv_gemm_epi* epi, void* stream) {
    int P, Q;
    if (int rc = check_geom(g, "stv_conv_fprop", P, Q)) return rc;
    STV_REQUIRE(src1 && w && y && (g->C2 == 0 || src2), "stv_conv_fprop: null pointer");
    STV_REQUIRE(!(epi && epi->accumulate), "stv_conv_fprop: accumulate is not supported");
    if (tma_eligible(g) && (long long)g->N*P*Q < (1ll << 31)) return fprop_tma(g, P, Q, src1, w, y, epi, (cudaStream_t)stream);
    ConvParams p = {};
    p.g = make_gather(g, src1, src2);
    p.npix = (long long)g->N*P*Q;
    STV_REQUIRE(p.npix < (1ll << 31), "stv_conv_fprop: too many output pixels");
    p.M = (int)p.npix; p.N = g->Cout; p.gridH = P; p.gridW = Q;
    p.bn = pick_bn(p.N, (p.M + CV_BM - 1)/CV_BM); p.b_mn = 0;
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
    STV_REQUIRE(!(epi && epi->accumulate), "stv_conv_dgrad: accumulate is not supported");
    const int Cin = g->C1 + g->C2;
    if (g->stride == 1 && !g->reflect && Cin % 4 == 0 && g->pad <= 120 && g->R <= 120 && g->S <= 120 && (long long)g->N*g->H*g->W < (1ll << 31))
        return dgrad_tma(g, P, Q, dy, w, dv, epi, (cudaStream_t)stream);
    if (g->stride > 1 && !g->reflect && g->H % g->stride == 0 && g->W % g->stride == 0 && g->pad <= 120 && g->R <= 120 && g->S <= 120 &&
        !(epi && (epi->aux || epi->res || epi->dact_src)) && (long long)g->N*g->H*g->W < (1ll << 31))
        return dgrad_tma_strided(g, P, Q, dy, w, dv, epi, (cudaStream_t)stream);
    // Gathered tensor = dY (N, P, Q, Cout); grid pixels = (padded, when reflecting) input pixels.
    ConvParams p = {};
    Gather& G = p.g;
    G.p1 = G.p2 = dy; G.C = G.C1 = g->Cout; G.C2 = 0; G.up1 = 0;
    G.H = P; G.W = Q; G.R = g->R; G.S = g->S; G.stride = g->stride; G.reflect = 0; G.dgrad = 1;
    G.pad = g->reflect ? 0 : g->pad;
    G.Ktot = g->R*g->S*g->Cout;
    G.smask = g->stride - 1; G.sshift = 0;
    while ((1 << G.sshift) < g->stride) ++G.sshift;
    p.gridH = g->reflect ? g->H + 2*g->pad : g->H;
    p.gridW = g->reflect ? g->W + 2*g->pad : g->W;
    p.npix = (long long)g->N*p.gridH*p.gridW;
    STV_REQUIRE(p.npix < (1ll << 31), "stv_conv_dgrad: too many pixels");
    p.M = (int)p.npix; p.N = Cin;
    p.bn = pick_bn(p.N, (p.M + CV_BM - 1)/CV_BM); p.b_mn = 1;
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
    if (tma_eligible(g) && (long long)g->N*P*Q < (1ll << 31)) return wgrad_tma(g, P, Q, src1, dy, dw, split_k, (cudaStream_t)stream);
    ConvParams p = {};
    p.g = make_gather(g, src1, src2);
    p.npix = (long long)g->N*P*Q;
    STV_REQUIRE(p.npix < (1ll << 31), "stv_conv_wgrad: too many output pixels");
    p.M = g->Cout; p.N = p.g.Ktot; p.gridH = P; p.gridW = Q;
    p.bn = pick_bn(p.N, 1 << 20); p.b_mn = 1; p.taps_kb = 1;
    p.kb_total = (int)((p.npix + CV_BK - 1)/CV_BK);
    if (stv_deterministic()) split_k = 1;
    if (split_k <= 0) {  // ~2 waves of CTAs, at least 8 k-blocks each
        const int tiles = ((p.M + CV_BM - 1)/CV_BM)*((p.N + p.bn - 1)/p.bn);
        split_k = (2*148)/tiles;  // floor: at most two full waves of CTAs
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
