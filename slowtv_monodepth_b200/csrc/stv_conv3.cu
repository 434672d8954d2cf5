// Row-segment convolution (forward and stride-1 data gradient of the 3-tap-wide layers): the filter taps of a filter row share
// ONE staged input slab, read through row-shifted UMMA descriptors; input rows are shared by the output rows above and below.
//
// STATUS: correct (bit-exact, tests/test_conv_gpu.py::test_rowseg_kernel_exact) but SLOWER than the im2col path on every layer
// of the step, so it is OFF by default (STV_CONV_ROWSEG=1 enables its heuristic, =2 forces it where the geometry allows). It is
// kept because it settles what bounds the small-channel convolutions (profiles/r2_conv3_timeline.txt):
//   * the idea: the im2col TMA path (stv_gemm.cu, ConvOperand mode 1) fetches a 128-pixel x 32-channel box per filter tap — a
//     3x3 convolution pulls every input element through L2 -> shared memory nine times (~10 TB/s aggregate on `res_l1` /
//     `upconv_0_1`, tools/bench_conv.py). Here a tile is NR output rows x 128 consecutive pixels; per (input row, 32-channel block)
//     ONE tiled TMA box brings 128 + S - 1 input pixels into shared memory (out-of-image pixels zero-filled by the copy engine), the
//     S taps are MMAs whose A descriptor START ADDRESS is shifted by s x 128 bytes (`tools/probe/umma_rowshift_probe.cu`: tcgen05.mma
//     honours row-shifted K-major slabs; the 128-byte swizzle is a function of the absolute shared-memory address), the slab feeds
//     the accumulators of the up-to-R output rows it belongs to, and the filters stay resident in shared memory: input traffic
//     drops from 9 to (NR + 2)/NR slabs per output row and a stage is ONE TMA instruction;
//   * what the in-kernel timeline shows: with the operand stream out of the way a stage's period is exactly 81-83 clocks x its
//     number of MMAs, for N = 16, 32 and 64 alike, shifted descriptors or not, chained on one accumulator or interleaved over
//     three: a tcgen05.mma kind::tf32 instruction (M = 128, K = 8) has a ~80-clock floor, so these layers (Cout = 16..64: 36 narrow
//     MMAs per 128 pixels and channel block) are bound by the MMA INSTRUCTION RATE, not by L2 or HBM. Two CTAs per SM (the im2col
//     kernel) overlap two instruction streams; this kernel (one CTA per SM, its ring and resident filters fill shared memory) cannot,
//     and the partly empty last segment of a row (W = 160: 128 + 32) costs it another 20-40 %.
//   What would actually help those layers: fewer, fatter MMAs — a 64-byte-swizzle (16-channel) K-block for the 16-channel decoder
//   level (half the instructions), or pixels on the N side with M = 64 — not a smaller operand stream.
//
// Structure and epilogue are those of the persistent GEMM (one CTA per SM: producer warp, MMA warp, 16 epilogue warps); GEMM rows
// enumerate (image row, segment, pixel) and the epilogue's RowMap (remap = 2) drops the pixels past the end of the row.
//
// Reference call sites: the 3x3 `nn.Conv2d` layers of `src/networks/decoders/monodepth.py:15-89` (conv_block / conv3x3,
// `decoders/utils.py:44-54`) and of the timm ResNet-18 pose encoder built at `src/networks/pose.py:40`.
#include <cstdlib>
#include <mutex>

#include "stv_common.cuh"
#include "stv_tc.cuh"
#include "stv_epi.cuh"
#include "stv_gemm.cuh"

namespace stv {

constexpr int C3_SEG = 128;                       // output pixels per tile
constexpr int C3_MAX_S = 3;                       // filter taps along x sharing a slab
constexpr int C3_A_ROWS = C3_SEG + C3_MAX_S - 1;  // 130 input pixels per slab
constexpr int C3_A_BYTES = 17*1024;               // 130 rows x 128 B, rounded up to the 1024-byte slab alignment

struct Conv3Params {
    int N, H, W, segs;       // output grid; segments per image row
    int NR, nrb;             // output rows per tile, row blocks per image
    int Cout, R, S, cblocks; // filter taps, 32-channel blocks of the input
    int lw, lh;              // input pixel of tap (0, 0) for output pixel (y, x): (y + lh, x + lw)
    int flip, b_mn, b_tap_cols, Cin_k;  // dgrad: flipped taps, MN-major filters (column = tap*b_tap_cols + n); fprop: k = tap*Cin_k + c
    int bn, stages;
    float* C;
    long long ldc;
    stv_gemm_epi e;
};

constexpr int C3_MAX_ACC = 8;   // accumulators: NR output rows of a tile, double-buffered across tiles

__device__ __forceinline__ void tma_load_4d_s(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n) : "memory");
}

// Tile = NR output rows x one 128-pixel segment x one column tile of bn output channels. The K loop walks the NR + R - 1 INPUT
// rows the tile touches (x 32-channel blocks): each staged input slab feeds the accumulators of up to R output rows (filter row
// r = input row - output row), so the input is fetched (NR + R - 1)/NR times instead of R*S times, and consecutive MMAs go to
// DIFFERENT accumulators: narrow MMAs (N = 16..64) are latency-bound at ~80 clocks each when chained on one accumulator
// (profiles/r2_conv3_timeline.txt) and interleaving the rows' chains hides that. The filters of the CTA's column tile are resident
// in shared memory (loaded once); an output row's accumulator is committed to the epilogue warps as soon as its last input row
// has been consumed, and the 16 epilogue warps drain several rows side by side.
__global__ void __launch_bounds__(GEMM_THREADS_WIDE, 1)
conv3_rowseg_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Conv3Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nepi = (int)(blockDim.x >> 5) - 2;
    const int tap_bytes = p.bn*GEMM_BK*4;                    // filters of one tap: bn rows x 128 B (either major)
    const int res_bytes = p.R*p.cblocks*p.S*tap_bytes;       // resident filter block: [r][cb][sx] tap tiles
    const int stages = p.stages, NR = p.NR;
    uint8_t* ringp = smem + res_bytes;
    float* staging = (float*)(ringp + (size_t)stages*C3_A_BYTES);
    uint64_t* full = (uint64_t*)(staging + nepi*EPI_WARP_FLOATS);
    uint64_t* empty = full + GEMM_MAX_STAGES;
    uint64_t* tmem_full = empty + GEMM_MAX_STAGES;   // [C3_MAX_ACC]
    uint64_t* tmem_empty = tmem_full + C3_MAX_ACC;   // [C3_MAX_ACC]
    uint64_t* bfull = tmem_empty + C3_MAX_ACC;       // resident filters have landed
    uint32_t* tmem_slot = (uint32_t*)(bfull + 1);

    const int chunks = (p.bn + 31) >> 5;                       // 32-column chunks per output row
    const int G = max(1, (nepi >> 2)/chunks);                  // output rows drained concurrently (a warp reads its own lane quarter only)
    const uint32_t acc_cols = p.bn <= 32 ? 32u : p.bn <= 64 ? 64u : 128u;
    uint32_t tmem_cols = 32;
    while (tmem_cols < 2u*(uint32_t)NR*acc_cols) tmem_cols <<= 1;
    const int nt = (p.Cout + p.bn - 1)/p.bn;
    const int mtiles = p.N*p.nrb*p.segs;
    const int my_nt = (int)(blockIdx.x % nt);                  // a CTA keeps ONE column tile (its filters stay resident)
    const int t_first = (int)(blockIdx.x/nt), t_step = (int)(gridDim.x/nt);
    const int yin_rows = NR + p.R - 1;
    if (threadIdx.x == 0) STV_TRACE(0);
    [[maybe_unused]] int trace_i = 0;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < C3_MAX_ACC; ++a) {
            tc::mbar_init(&tmem_full[a], 1);
            tc::mbar_init(&tmem_empty[a], (uint32_t)(4*chunks));
        }
        tc::mbar_init(bfull, 1);
        tc::fence_barrier_init();
    } else if (warp == 1) {
        tc::tmem_alloc(tmem_slot, tmem_cols);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t res0 = tc::smem_u32(smem), ring = tc::smem_u32(ringp), full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
    // bytes one stage receives: the copy engine writes the whole box, zero-filled where the image ends
    const uint32_t a_bytes = (uint32_t)((C3_SEG + p.S - 1)*GEMM_BK*4);
    if (threadIdx.x == 0) STV_TRACE(1);

    // tile index -> (image n, first output row y0, first pixel x0)
    auto decode = [&](int t, int& n, int& y0, int& x0) {
        const int seg = t % p.segs, rest = t/p.segs, rb = rest % p.nrb;
        n = rest/p.nrb; y0 = rb*NR; x0 = seg*C3_SEG;
    };

    if (warp == 0) {
        if (tc::elect_one()) {   // resident filters of this CTA's column tile, in image-offset order (r, cb, sx)
            const uint32_t bb = tc::smem_u32(bfull);
            const int n0 = my_nt*p.bn;
            tc::mbar_arrive_expect_tx_s(bb, (uint32_t)res_bytes);
            for (int r = 0; r < p.R; ++r)
                for (int cb = 0; cb < p.cblocks; ++cb)
                    for (int sx = 0; sx < p.S; ++sx) {
                        const int tap = (p.flip ? p.R - 1 - r : r)*p.S + (p.flip ? p.S - 1 - sx : sx);
                        const uint32_t dst = res0 + (uint32_t)(((r*p.cblocks + cb)*p.S + sx)*tap_bytes);
                        if (!p.b_mn) tc::tma_load_2d_s(dst, &tmB, bb, tap*p.Cin_k + cb*GEMM_BK, n0);
                        else
                            for (int j = 0; j < (p.bn >> 5); ++j)
                                tc::tma_load_2d_s(dst + j*SLAB_MN_BYTES, &tmB, bb, tap*p.b_tap_cols + n0 + 32*j, cb*GEMM_BK);
                    }
        }
        int s = 0;
        uint32_t ph = 0;
        for (int t = t_first; t < mtiles; t += t_step) {
            int n, y0, x0;
            decode(t, n, y0, x0);
            for (int yi = 0; yi < yin_rows; ++yi) {
                if (y0 + yi - (p.R - 1) >= p.H) break;     // no output row of the image uses this input row (last, partial row block)
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    tc::mbar_wait_spin_s(empty0 + 8*s, ph ^ 1u);
                    if (lane == 0) { STV_TRACE_KB(16, trace_i); ++trace_i; }
                    if (tc::elect_one()) {
                        const uint32_t fb = full0 + 8*s;
                        tc::mbar_arrive_expect_tx_s(fb, a_bytes);
                        tma_load_4d_s(ring + (uint32_t)(s*C3_A_BYTES), &tmA, fb, cb*GEMM_BK, x0 + p.lw, y0 + yi + p.lh, n);
                    }
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::umma_idesc_tf32(GEMM_BM, p.bn, false, p.b_mn != 0);
        const uint64_t da0 = tc::umma_desc_template(false, SLAB_MN_BYTES), db0 = tc::umma_desc_template(p.b_mn != 0, SLAB_MN_BYTES);
        const uint32_t kbs = tc::umma_desc_kstep(p.b_mn != 0);
        const uint32_t tmem_full0 = tc::smem_u32(tmem_full), tmem_empty0 = tc::smem_u32(tmem_empty);
        const uint32_t tap_units = (uint32_t)tap_bytes >> 4;
        tc::mbar_wait_spin_s(tc::smem_u32(bfull), 0u);
        int s = 0, jt = 0;
        uint32_t ph = 0, used = 0;   // bit a: parity of the number of times accumulator a has been handed to the epilogue
        for (int t = t_first; t < mtiles; t += t_step, ++jt) {
            int n, y0, x0;
            decode(t, n, y0, x0);
            const int rows = min(NR, p.H - y0);              // output rows of this tile that exist (the last row block may be partial)
            const int acc0 = (jt & 1)*NR;
            for (int yi = 0; yi < yin_rows; ++yi) {
                if (y0 + yi - (p.R - 1) >= p.H) break;
                // output rows fed by input row yi: q = yi - r, r = 0..R-1
                const int q_lo = max(0, yi - (p.R - 1)), q_hi = min(rows - 1, yi);
                if (yi < rows) {   // row q = yi receives its first product now: its accumulator must have been drained
                    const int a = acc0 + yi;
                    tc::mbar_wait_spin_s(tmem_empty0 + 8*a, ((used >> a) & 1u) ^ 1u);
                    tc::tcgen05_fence_after();
                    used ^= 1u << a;
                }
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    tc::mbar_wait_spin_s(full0 + 8*s, ph);
                    if (lane == 0) { STV_TRACE_KB(528, trace_i); ++trace_i; }
                    tc::tcgen05_fence_after();
                    if (tc::elect_one()) {
                        const uint32_t a = (ring + (uint32_t)(s*C3_A_BYTES)) >> 4;
                        for (int sx = 0; sx < p.S; ++sx) {
#pragma unroll
                            for (int k8 = 0; k8 < GEMM_BK/8; ++k8) {
                                // tap sx reads slab rows sx .. sx + 127: start address + sx * 128 B (8 descriptor units)
                                const uint64_t da = da0 + (a + 8u*(uint32_t)sx + 2u*(uint32_t)k8);
                                for (int q = q_lo; q <= q_hi; ++q) {   // consecutive MMAs accumulate into different rows
                                    const int r = yi - q;
                                    const uint64_t db = db0 + ((res0 >> 4) + (uint32_t)((r*p.cblocks + cb)*p.S + sx)*tap_units + kbs*(uint32_t)k8);
                                    tc::umma_tf32(tmem_base + (uint32_t)(acc0 + q)*acc_cols, da, db, idesc, (r > 0 || cb > 0 || sx > 0 || k8 > 0) ? 1u : 0u);
                                }
                            }
                        }
                        tc::umma_commit_s(empty0 + 8*s);
                        // output row q = yi - (R - 1) has seen its last input row once the last channel block is in
                        if (cb == p.cblocks - 1 && yi >= p.R - 1 && yi - (p.R - 1) < rows) tc::umma_commit_s(tmem_full0 + 8*(acc0 + yi - (p.R - 1)));
                    }
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
            }
            if (lane == 0 && jt == 0) STV_TRACE(2);
        }
        __syncwarp();
    } else {
        const RowMap rm = {p.ldc, 2, 0, p.segs, 0, p.W, 0, 0, 0};
        const int Mv = p.N*p.H*p.segs*C3_SEG;
        const int ew = warp - 2;
        const int grp = (ew >> 2)/chunks, ci = (ew >> 2) - grp*chunks;   // row slot and column chunk of this warp
        if (grp < G) {
            int jt = 0, job = 0;
            uint32_t seen = 0;   // bit a: parity of the number of finished rows accumulator a has held (all groups count alike)
            for (int t = t_first; t < mtiles; t += t_step, ++jt) {
                int n, y0, x0;
                decode(t, n, y0, x0);
                const int rows = min(NR, p.H - y0);
                for (int q = 0; q < rows; ++q, ++job) {
                    const int acc = (jt & 1)*NR + q;
                    const uint32_t par = (seen >> acc) & 1u;
                    seen ^= 1u << acc;
                    if (job % G != grp) continue;
                    tc::mbar_wait(&tmem_full[acc], par);
                    tc::tcgen05_fence_after();
                    const int m0 = ((n*p.H + y0 + q)*p.segs + x0/C3_SEG)*C3_SEG;
                    if (warp == 2 && lane == 0 && job == 0) STV_TRACE(3);
                    epilogue_tile(tmem_base + (uint32_t)acc*acc_cols, warp & 3, lane, m0, my_nt*p.bn, p.bn, Mv, p.Cout, p.C, rm, p.e,
                                  staging + ew*EPI_WARP_FLOATS, ci*32, chunks*32);
                    if (warp == 2 && lane == 0 && job == 0) STV_TRACE(4);
                    tc::tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
                }
            }
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) STV_TRACE(5);
    if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Channels-last (N,H,W,C) fp32 tensor read in TILED mode: boxes of {32 channels, `pixels` consecutive pixels of one image row};
// pixels / rows / channels outside the tensor (negative coordinates included) read as zeros.
static int make_tmap_rowseg(CUtensorMap* tm, const float* base, int N, int H, int W, int C, int pixels) {
    static EncodeTiledFn4 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn4)ptr;
    });
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the CUDA driver"); return STV_E_CUDA; }
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C*4, (cuuint64_t)W*C*4, (cuuint64_t)H*W*C*4};
    const cuuint32_t box[4] = {32u, (cuuint32_t)pixels, 1u, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): tensor (%d,%d,%d,%d), box of %d pixels", (int)r, N, H, W, C, pixels);
        return STV_E_CUDA;
    }
    return STV_OK;
}

// 0 = never (default: measured slower, see the header), 1 = heuristic, 2 = whenever the geometry allows (tests; read per call so
// that a test can switch it).
static int conv3_mode() {
    const char* v = getenv("STV_CONV_ROWSEG");
    return v ? atoi(v) : 0;
}

// Column-tile width, output rows per tile and ring depth, or false when the resident filter block does not fit: the widest
// multiple of 32 (<= 128, fewest padded columns first) whose R*S*Cin/32 tap tiles leave room for a >= 4-slab input ring beside the
// epilogue staging; rows per tile = as many accumulators as TMEM holds twice over (double-buffered across tiles), at most 4.
static bool conv3_shape(int Cin, int Cout, int R, int S, int& bn, int& NR, int& stages) {
    const int staging = (GEMM_THREADS_WIDE/32 - 2)*EPI_WARP_FLOATS*4;
    const int budget = 226*1024 - staging - 2048;
    int order[4], cost[4];
    for (int i = 0; i < 4; ++i) { order[i] = 128 - 32*i; cost[i] = ((Cout + order[i] - 1)/order[i])*order[i]; }
    for (int i = 1; i < 4; ++i)   // stable insertion sort by padded width (ties keep the wider tile first)
        for (int k = i; k > 0 && cost[k] < cost[k - 1]; --k) { const int c = cost[k], o = order[k]; cost[k] = cost[k - 1]; order[k] = order[k - 1]; cost[k - 1] = c; order[k - 1] = o; }
    for (int i = 0; i < 4; ++i) {
        const int res = R*S*((Cin + GEMM_BK - 1)/GEMM_BK)*order[i]*GEMM_BK*4;
        if (res + 4*C3_A_BYTES > budget) continue;
        if ((Cout + order[i] - 1)/order[i] > 148) continue;
        bn = order[i];
        const int acc_cols = bn <= 32 ? 32 : bn <= 64 ? 64 : 128;
        NR = 512/(2*acc_cols) < 4 ? 512/(2*acc_cols) : 4;
        stages = (budget - res)/C3_A_BYTES;
        if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
        return true;
    }
    return false;
}

// Geometry: S <= 3 taps along x, stride 1, 32-channel blocks, filters small enough to stay resident; the output grid is
// (N, oH, oW). Heuristic (mode 1): many pixels per channel — the regime where nine im2col boxes per input element saturate L2 -> SM
// — and rows long enough that the partly empty last segment costs less than the smaller operand stream saves.
bool conv3_eligible(int N, int oH, int oW, int Cin, int Cout, int R, int S, int stride) {
    const int mode = conv3_mode();
    // (a partial last channel block is fine: the copy engine zero-fills the slab's missing channels, which nulls whatever the filter box holds there)
    if (mode == 0 || stride != 1 || S < 2 || S > C3_MAX_S || R < 1 || R > 7 || Cin % 4 != 0 || Cout % 4 != 0) return false;
    if ((long long)N*oH*((oW + C3_SEG - 1)/C3_SEG)*C3_SEG >= (1ll << 31)) return false;
    int bn, NR, stages;
    if (!conv3_shape(Cin, Cout, R, S, bn, NR, stages)) return false;
    if (mode == 2) return true;
    const int segs = (oW + C3_SEG - 1)/C3_SEG;
    const double fill = (double)oW/(segs*C3_SEG);
    return fill >= 0.6 && (long long)N*oH*oW >= 100000;
}

// y (N, oH, oW, Cout) = epilogue( sum_{r,s,c} x[n, y + r + lh, x + s + lw, c] * filter(r, s, c, k) ), zero outside x.
//   fprop:  filters K-major (Cout x R*S*Cin, k = (r*S + s)*Cin + c)
//   dgrad:  filters MN-major (rows = reduction channel, column = tap*b_tap_cols + output channel), flipped taps
int conv3_launch(const float* x, int N, int iH, int iW, int Cin, int oH, int oW, int lw, int lh, int R, int S, const float* w, int Cout,
                 int flip, int b_mn, int b_tap_cols, long long b_rows, long long b_cols, float* y, long long ldc, const stv_gemm_epi* epi,
                 cudaStream_t st, const char* what) {
    Conv3Params p = {};
    int bn, NR, stages;
    if (!conv3_shape(Cin, Cout, R, S, bn, NR, stages)) { set_error("%s: the filters do not fit the row-segment kernel", what); return STV_E_ARG; }
    p.N = N; p.H = oH; p.W = oW; p.segs = (oW + C3_SEG - 1)/C3_SEG;
    p.NR = NR; p.nrb = (oH + NR - 1)/NR;
    p.Cout = Cout; p.R = R; p.S = S; p.cblocks = (Cin + GEMM_BK - 1)/GEMM_BK; p.lw = lw; p.lh = lh;
    p.flip = flip; p.b_mn = b_mn; p.b_tap_cols = b_tap_cols; p.Cin_k = Cin;
    p.bn = bn; p.stages = stages;
    p.C = y; p.ldc = ldc;
    if (epi) p.e = *epi;
    const int tap_bytes = bn*GEMM_BK*4, res_bytes = R*S*p.cblocks*tap_bytes;
    const int staging = (GEMM_THREADS_WIDE/32 - 2)*EPI_WARP_FLOATS*4;
    CUtensorMap tmA, tmB;
    if (int rc = make_tmap_rowseg(&tmA, x, N, iH, iW, Cin, C3_SEG + S - 1)) return rc;
    if (int rc = b_mn ? make_tmap_2d(&tmB, w, b_rows, b_cols, b_cols, 32, 1) : make_tmap_2d(&tmB, w, b_rows, b_cols, b_cols, bn, 0)) return rc;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(conv3_rowseg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024); });
    if (attr_err != cudaSuccess) { set_error("%s: cudaFuncSetAttribute failed (%s)", what, cudaGetErrorString(attr_err)); return STV_E_CUDA; }
    const size_t smem = (size_t)res_bytes + (size_t)stages*C3_A_BYTES + staging + 1024 + (2*GEMM_MAX_STAGES + 2*C3_MAX_ACC + 1)*8 + 16;
    const int nt = (Cout + bn - 1)/bn;
    const long long mtiles = (long long)N*p.nrb*p.segs;
    int sms = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    // a CTA keeps one column tile: the grid is a whole number of column-tile groups
    const long long per = mtiles < sms/nt ? mtiles : sms/nt;
    const int grid = (int)per*nt;
    conv3_rowseg_kernel<<<grid, GEMM_THREADS_WIDE, smem, st>>>(tmA, tmB, p);
    count_launch();
    return check_launch(what);
}

}  // namespace stv

#ifdef STV_GEMM_TRACE
extern "C" int stv_debug_conv3_trace(unsigned long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, stv::g_gemm_trace, sizeof(unsigned long long)*(size_t)(n < 1040 ? n : 1040));
}
#endif
