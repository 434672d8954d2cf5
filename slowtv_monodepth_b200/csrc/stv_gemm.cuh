// Shared declarations of the tcgen05 GEMM kernel (stv_gemm.cu) for the convolution front-ends (stv_conv.cu).
#pragma once
#include "stv_common.cuh"
#include "stv_tc.cuh"

namespace stv {

constexpr int GEMM_BM = 128, GEMM_BK = 32, GEMM_THREADS = 320, GEMM_MAX_STAGES = 8;  // warps: TMA, MMA, 8 x epilogue
constexpr int GEMM_THREADS_WIDE = 576;            // one CTA per SM: TMA, MMA, 16 x epilogue (the persistent kernels take either)
constexpr int GEMM_A_BYTES = GEMM_BM*GEMM_BK*4;  // 16 KB per stage
constexpr int SLAB_MN_BYTES = 32*128;            // MN-major slab: 32 k-rows x 128 B

// im2col addressing of one operand through a TMA im2col tensor map over a channels-last (N,H,W,C) tensor.
//   mode 1  A = im2col rows (fprop / stride-1 dgrad): GEMM row m = grid pixel (n, py, px); k-block = (filter tap, 32 channels)
//   mode 2  B = im2col rows (wgrad): reduction index = grid pixel; every 32-column slab of the tile = (filter tap, 32 channels)
struct ConvOperand {
    int mode;
    int gridH, gridW;   // pixel grid enumerated by the GEMM rows (mode 1) / the reduction index (mode 2)
    int stride, lw, lh; // TMA base coordinate of grid pixel (py, px): (lh + py*stride, lw + px*stride)
    int R, S, C;        // filter taps; channels of the im2col tensor
    int cblocks;        // mode 1: k-blocks per tap = ceil(C/32)
    int flip;           // dgrad: tap (r, s) reads offsets (R-1-r, S-1-s)
    int b_tap_cols;     // dgrad: B (filters, MN-major boxes) column = full_tap*b_tap_cols + n0 + 32*slab, row = channel block, where
    int r0, s0, tstep, Sfull;  //   full_tap = (r0 + tstep*r)*Sfull + (s0 + tstep*s): the sub-filter of one output parity (strided dgrad)
};

struct GemmParams {
    int M, N, K;
    int bn, stages, a_mn, b_mn;
    int pf;                          // L2 prefetch distance in k-blocks (0 = off), set by launch_gemm
    int kb_total, kb_per_split, splits;
    float* C;
    long long ldc;
    stv_gemm_epi e;
    ConvOperand cv;
    int remap, oH, oW, ost, oa, ob;  // RowMap of the output (stv_epi.cuh); rows enumerate (n, cv.gridH, cv.gridW) when remap != 0
    const float* pair_B;             // host only: B of a plain GEMM (the CTA-pair kernel re-encodes it with half-tile boxes), or NULL
    long long pair_ldb;
};

int make_tmap_2d(CUtensorMap* tm, const float* base, long long rows, long long cols, long long ld, int box_rows, int mn_major);
// Channels-last (N,H,W,C) fp32 tensor read in im2col mode: boxes of `pixels` base pixels x 32 channels, base pixels restricted
// to [lw, W + uw) x [lh, H + uh) and stepped by `stride`; out-of-image pixels / channels read as zeros.
int make_tmap_im2col(CUtensorMap* tm, const float* base, int N, int H, int W, int C, int lw, int lh, int uw, int uh, int stride,
                     int pixels, int mn_major);
// Dense fp32 (planes, H, W) tensor read as un-swizzled boxes {bw, bh, bp} (W % 4 == 0, bw % 4 == 0); zero OOB fill.
int make_tmap_3d(CUtensorMap* tm, const float* base, long long W, long long H, long long planes, int bw, int bh, int bp);
int pick_bn(int N, long long row_tiles);
// Fills p.stages, launches grid (ceil(M/128), ceil(N/bn), splits). p.bn, p.kb_total, p.kb_per_split must be set.
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, int splits, cudaStream_t stream, const char* what);

// Developer instrumentation (-DSTV_GEMM_TRACE, tools/gemm_trace.py; never in the shipped build): clock64 stamps of CTA 0's roles.
//   [0] entry  [1] set-up done  [2] last commit  [3] epilogue sees the accumulator  [4] epilogue done  [5] exit
//   [16 + i] producer passes the EMPTY wait of its i-th k-block   [528 + i] MMA thread passes the FULL wait of its i-th k-block
#ifdef STV_GEMM_TRACE
static __device__ unsigned long long g_gemm_trace[1040];   // one buffer per translation unit (no relocatable device code)
#define STV_TRACE(slot) do { if (blockIdx.x == 0) g_gemm_trace[(slot)] = (unsigned long long)clock64(); } while (0)
#define STV_TRACE_KB(base, i) do { if (blockIdx.x == 0 && (i) < 512) g_gemm_trace[(base) + (i)] = (unsigned long long)clock64(); } while (0)
#else
#define STV_TRACE(slot) do {} while (0)
#define STV_TRACE_KB(base, i) do {} while (0)
#endif

// Row-segment convolution (stv_conv3.cu): the taps of a filter row share one staged input slab (row-shifted UMMA descriptors).
bool conv3_eligible(int N, int oH, int oW, int Cin, int Cout, int R, int S, int stride);
int conv3_launch(const float* x, int N, int iH, int iW, int Cin, int oH, int oW, int lw, int lh, int R, int S, const float* w, int Cout,
                 int flip, int b_mn, int b_tap_cols, long long b_rows, long long b_cols, float* y, long long ldc, const stv_gemm_epi* epi,
                 cudaStream_t st, const char* what);

}  // namespace stv
