// Memory-bound companions of the implicit-GEMM convolutions (channels-last fp32):
//   stv_grad_pull   routes the data gradient the convolution kernel produced for its *virtual* input
//                   cat(up2(src1), src2), possibly on the reflection-padded grid, back to one real source tensor:
//                   channel slice + reflection fold + 2x2 sum-pool (the adjoint of nearest-x2 upsampling) in one pass
//                   (adjoints of decoders/monodepth.py:76-79 `F.interpolate` / `torch.cat` and utils.py:44-46 `padding_mode='reflect'`);
//   stv_act_bwd     dZ = dA * act'(Y) fused with the bias gradient (column sums of dZ);
//   stv_colsum      column sums of a row-major matrix (bias gradients of the GEMM layers).
#include "stv_common.cuh"
#include "stv_epi.cuh"

namespace stv {

// Sum over the positions of the padded axis (length n + 2*pad) that reflection padding maps onto index i of the original axis.
// Returns the number of such positions (1..3 for pad < n) and writes them to q[].
__device__ __forceinline__ int fold_positions(int i, int n, int pad, int* q) {
    int cnt = 0;
    q[cnt++] = i + pad;
    if (pad > 0) {
        if (i >= 1 && i <= pad) q[cnt++] = pad - i;                      // left border: padded index pad - i reflects to i
        if (i <= n - 2 && i >= n - 1 - pad) q[cnt++] = 2*(n - 1) - i + pad;  // right border
    }
    return cnt;
}

__global__ void grad_pull_kernel(int N, int H, int W, int C, const float* __restrict__ src, int Cs, int c_off, int pad, int pool,
                                 float* __restrict__ dst, int accumulate) {
    const int c4 = C >> 2;
    const long long total = (long long)N*H*W*c4;
    const int Hs = H*pool, Ws = W*pool, Hp = Hs + 2*pad, Wp = Ws + 2*pad;
    for (unsigned idx = blockIdx.x*blockDim.x + threadIdx.x; idx < (unsigned)total; idx += gridDim.x*blockDim.x) {  // 32-bit index math (count < 2^31, checked on the host)
        const int c = (int)(idx % c4)*4;
        unsigned r = idx/c4;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H);
        const int n = (int)(r/H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int dy = 0; dy < pool; ++dy) {
            int qy[3];
            const int ny = fold_positions(y*pool + dy, Hs, pad, qy);
            for (int dx = 0; dx < pool; ++dx) {
                int qx[3];
                const int nx = fold_positions(x*pool + dx, Ws, pad, qx);
                for (int a = 0; a < ny; ++a)
                    for (int b = 0; b < nx; ++b) {
                        const float4 v = __ldg((const float4*)(src + ((size_t)(n*Hp + qy[a])*Wp + qx[b])*Cs + c_off + c));
                        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                    }
            }
        }
        float4* d = (float4*)(dst + (size_t)idx*4);
        if (accumulate) { const float4 o = *d; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
        *d = acc;
    }
}

// dZ[m, c] = dA[m, c] * act'(Y[m, c]); dbias[c] += sum_m dZ[m, c].  Block = 32 x 8 threads: 32*VEC columns x ROWS rows.
template <int ACT>
__global__ void __launch_bounds__(256) act_bwd_kernel(long long M, int C, const float* __restrict__ dA, const float* __restrict__ Y,
                                                      float* __restrict__ dZ, float* __restrict__ dbias, int rows_per_block) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x*32 + tx;
    const long long r0 = (long long)blockIdx.y*rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s = 0.f;
    if (c < C) {
        for (long long r = r0 + ty; r < r1; r += 8) {
            const size_t o = (size_t)r*C + c;
            const float g = __ldg(dA + o)*act_bwd(ACT, __ldg(Y + o));
            dZ[o] = g;
            s += g;
        }
    }
    if (dbias) {
        red[ty][tx] = s;
        __syncthreads();
        if (ty == 0 && c < C) {
#pragma unroll
            for (int k = 1; k < 8; ++k) s += red[k][tx];
            atomicAdd(dbias + c, s);
        }
    }
}

__global__ void __launch_bounds__(256) colsum_kernel(long long M, int C, long long ld, const float* __restrict__ X, float* __restrict__ out,
                                                     int rows_per_block) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x*32 + tx;
    const long long r0 = (long long)blockIdx.y*rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s = 0.f;
    if (c < C)
        for (long long r = r0 + ty; r < r1; r += 8) s += __ldg(X + (size_t)r*ld + c);
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
#pragma unroll
        for (int k = 1; k < 8; ++k) s += red[k][tx];
        atomicAdd(out + c, s);
    }
}

// ---- vectorised row loops with per-column reductions (float4 = 4 channels per thread) ---------------------------------------------
// A (M, C) channels-last matrix has G = C/4 column groups. G >= 32: a block owns 32 consecutive groups (lane = group) and its 8
// warps stride over the rows. G < 32 (a power of two): one warp row covers 32/G matrix rows x all G groups, so every lane of
// every load is useful even for 16-64 channel tensors (the decoder / ResNet layers where the scalar mapping idled most lanes).
// Op::apply(o4, c, acc) handles the float4 at element offset 4*o4 (columns c..c+3) and adds into NACC float4 accumulators;
// the accumulators are combined lane -> warp -> block in a fixed order and leave the block as one atomic per column and block.
template <class Op, bool DBL>
__global__ void __launch_bounds__(256) colred4_kernel(long long M, int C, int rows_per_block, Op op, void* __restrict__ out0,
                                                      void* __restrict__ out1) {
    __shared__ float red[Op::NACC][8][32][4 + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int G = C >> 2;
    const bool wide = G >= 32;
    const int g = wide ? blockIdx.x*32 + lane : lane & (G - 1);
    const int rsub = wide ? 0 : lane/G, rpw = wide ? 1 : 32/G;           // row within a warp pass, rows per warp pass
    const long long r0 = (long long)blockIdx.y*rows_per_block, r1 = min(r0 + rows_per_block, M);
    float4 acc[Op::NACC];
#pragma unroll
    for (int k = 0; k < Op::NACC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g < G) {
        // four rows per trip in straight-line code: their loads are independent (read-only path), so a thread keeps 4-12 requests in
        // flight instead of 1-3 — these reductions were latency-bound at ~2.5x their HBM time (BatchNorm statistics / backward
        // sums, r1f profile). A `#pragma unroll` alone keeps the bound check between the copies and hoists nothing.
        const long long step = 8*rpw;
        long long r = r0 + (long long)wid*rpw + rsub;
        for (; r + 3*step < r1; r += 4*step) {
            op.apply((size_t)r*G + g, g*4, acc);
            op.apply((size_t)(r + step)*G + g, g*4, acc);
            op.apply((size_t)(r + 2*step)*G + g, g*4, acc);
            op.apply((size_t)(r + 3*step)*G + g, g*4, acc);
        }
        for (; r < r1; r += step) op.apply((size_t)r*G + g, g*4, acc);
    }
    if (!wide)  // lanes with the same group (G apart) -> lane < G
        for (int o = G; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < Op::NACC; ++k) {
                acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, o); acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, o);
                acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, o); acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, o);
            }
        }
    if (out0 == nullptr) return;
#pragma unroll
    for (int k = 0; k < Op::NACC; ++k) { red[k][wid][lane][0] = acc[k].x; red[k][wid][lane][1] = acc[k].y; red[k][wid][lane][2] = acc[k].z; red[k][wid][lane][3] = acc[k].w; }
    __syncthreads();
    const int nl = wide ? 32 : G;  // lanes holding distinct groups
    for (int e = threadIdx.x; e < Op::NACC*nl*4; e += 256) {
        const int k = e/(nl*4), l = (e/4) % nl, j = e & 3;
        const int gg = wide ? blockIdx.x*32 + l : l;
        if (gg >= G) continue;
        void* out = k == 0 ? out0 : out1;
        if (DBL) {
            double v = 0.;
#pragma unroll
            for (int w2 = 0; w2 < 8; ++w2) v += (double)red[k][w2][l][j];
            atomicAdd((double*)out + gg*4 + j, v);
        } else {
            float v = 0.f;
#pragma unroll
            for (int w2 = 0; w2 < 8; ++w2) v += red[k][w2][l][j];
            atomicAdd((float*)out + gg*4 + j, v);
        }
    }
}

__host__ __device__ inline bool colred4_ok(int C) { const int G = C >> 2; return (C & 3) == 0 && G >= 1 && (G >= 32 || (G & (G - 1)) == 0); }

struct OpColSum {
    static constexpr int NACC = 1;
    const float* X;
    __device__ __forceinline__ void apply(size_t o4, int, float4* acc) const {
        const float4 v = __ldg((const float4*)X + o4);
        acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
    }
};
template <int ACT>
struct OpActBwd {
    static constexpr int NACC = 1;
    const float* dA; const float* Y; float* dZ;
    __device__ __forceinline__ void apply(size_t o4, int, float4* acc) const {
        const float4 a = __ldg((const float4*)dA + o4), y = __ldg((const float4*)Y + o4);
        const float4 g = make_float4(a.x*act_bwd(ACT, y.x), a.y*act_bwd(ACT, y.y), a.z*act_bwd(ACT, y.z), a.w*act_bwd(ACT, y.w));
        ((float4*)dZ)[o4] = g;
        acc[0].x += g.x; acc[0].y += g.y; acc[0].z += g.z; acc[0].w += g.w;
    }
};
struct OpBnStats {
    static constexpr int NACC = 2;
    const float* X;
    __device__ __forceinline__ void apply(size_t o4, int, float4* acc) const {
        const float4 v = __ldg((const float4*)X + o4);
        acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
        acc[1].x = fmaf(v.x, v.x, acc[1].x); acc[1].y = fmaf(v.y, v.y, acc[1].y); acc[1].z = fmaf(v.z, v.z, acc[1].z); acc[1].w = fmaf(v.w, v.w, acc[1].w);
    }
};
struct OpBnBwdReduce {
    static constexpr int NACC = 2;
    const float* dY; const float* Y; const float* X; const float* mean; const float* rstd; int relu;
    __device__ __forceinline__ void apply(size_t o4, int c, float4* acc) const {
        float4 g = __ldg((const float4*)dY + o4);
        if (relu) {
            const float4 y = __ldg((const float4*)Y + o4);
            g.x = y.x > 0.f ? g.x : 0.f; g.y = y.y > 0.f ? g.y : 0.f; g.z = y.z > 0.f ? g.z : 0.f; g.w = y.w > 0.f ? g.w : 0.f;
        }
        const float4 x = __ldg((const float4*)X + o4), m = __ldg((const float4*)(mean + c)), r = __ldg((const float4*)(rstd + c));
        acc[0].x += g.x; acc[0].y += g.y; acc[0].z += g.z; acc[0].w += g.w;
        acc[1].x = fmaf(g.x, (x.x - m.x)*r.x, acc[1].x); acc[1].y = fmaf(g.y, (x.y - m.y)*r.y, acc[1].y);
        acc[1].z = fmaf(g.z, (x.z - m.z)*r.z, acc[1].z); acc[1].w = fmaf(g.w, (x.w - m.w)*r.w, acc[1].w);
    }
};

// Launch geometry of colred4_kernel: grid.x = 32-group column blocks (1 when G < 32), grid.y = row blocks.
static void colred4_grid(long long M, int C, dim3& grid, int& rpb) {
    const int G = C >> 2, colb = G >= 32 ? (G + 31)/32 : 1;
    long long want = (4ll*148*8 + colb - 1)/colb;   // ~4 waves of 8 resident blocks per SM
    rpb = (int)((M + want - 1)/want);
    const int min_rows = G >= 32 ? 64 : 64*(32/G);
    if (rpb < min_rows) rpb = min_rows;
    if (stv_deterministic()) rpb = (int)(M < 0x7fffffffll ? M : 0x7fffffffll);   // one row block: one contributor per column
    grid = dim3(colb, (unsigned)((M + rpb - 1)/rpb));
}

// ---- train-mode BatchNorm over the rows of a channels-last (M, C) matrix (M = N*H*W) ------------------------------------------
// Reference: the nn.BatchNorm2d layers of the timm ResNet encoder (src/networks/pose.py:40, depth.py:97), batch statistics
// per GPU (no SyncBN, api/train/train.py:105-119). Statistics are accumulated in double (block partials -> atomicAdd(double)).
__global__ void __launch_bounds__(256) bn_stats_kernel(long long M, int C, const float* __restrict__ X, double* __restrict__ sums,
                                                       int rows_per_block) {
    __shared__ float red[2][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x*32 + tx;
    const long long r0 = (long long)blockIdx.y*rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s = 0.f, q = 0.f;
    if (c < C)
        for (long long r = r0 + ty; r < r1; r += 8) {
            const float v = __ldg(X + (size_t)r*C + c);
            s += v;
            q = fmaf(v, v, q);
        }
    red[0][ty][tx] = s;
    red[1][ty][tx] = q;
    __syncthreads();
    if (ty == 0 && c < C) {
        double ds = 0., dq = 0.;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ds += red[0][k][tx]; dq += red[1][k][tx]; }
        atomicAdd(sums + c, ds);
        atomicAdd(sums + C + c, dq);
    }
}

// mean / rstd from the sums; running statistics updated like nn.BatchNorm2d (momentum, unbiased variance).
__global__ void bn_finalize_kernel(long long M, int C, const double* __restrict__ sums, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ run_mean, float* __restrict__ run_var) {
    const int c = blockIdx.x*blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c]/(double)M;
    double var = sums[C + c]/(double)M - m*m;
    var = var > 0. ? var : 0.;
    mean[c] = (float)m;
    rstd[c] = (float)(1.0/sqrt(var + (double)eps));
    if (run_mean) {
        run_mean[c] = (1.f - momentum)*run_mean[c] + momentum*(float)m;
        run_var[c] = (1.f - momentum)*run_var[c] + momentum*(float)(var*(double)M/(double)(M > 1 ? M - 1 : 1));
    }
}

// y = [relu]( (x - mean)*rstd*gamma + beta [+ res] ), 4 channels per thread.
__global__ void __launch_bounds__(256) bn_apply_kernel(long long total4, int C, const float* __restrict__ X, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ res, int relu,
                                                       float* __restrict__ Y) {
    const int c4 = C >> 2;
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < (unsigned)total4; i += gridDim.x*blockDim.x) {  // 32-bit index math (count < 2^31, checked on the host)
        const int c = (int)(i % c4)*4;
        const float4 x = __ldg((const float4*)X + i), m = __ldg((const float4*)(mean + c)), r = __ldg((const float4*)(rstd + c));
        const float4 g = __ldg((const float4*)(gamma + c)), b = __ldg((const float4*)(beta + c));
        float4 y = make_float4(fmaf((x.x - m.x)*r.x, g.x, b.x), fmaf((x.y - m.y)*r.y, g.y, b.y), fmaf((x.z - m.z)*r.z, g.z, b.z),
                               fmaf((x.w - m.w)*r.w, g.w, b.w));
        if (res) { const float4 s = __ldg((const float4*)res + i); y.x += s.x; y.y += s.y; y.z += s.z; y.w += s.w; }
        if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
        ((float4*)Y)[i] = y;
    }
}

// sums[c] += sum_m dz, sums[C + c] += sum_m dz*xhat, with dz = dy*(y > 0) when relu.
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(long long M, int C, const float* __restrict__ dY, const float* __restrict__ Y,
                                                            const float* __restrict__ X, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, int relu, double* __restrict__ sums,
                                                            int rows_per_block) {
    __shared__ float red[2][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x*32 + tx;
    const long long r0 = (long long)blockIdx.y*rows_per_block, r1 = min(r0 + rows_per_block, M);
    float s = 0.f, q = 0.f;
    if (c < C) {
        const float m = __ldg(mean + c), rs = __ldg(rstd + c);
        for (long long r = r0 + ty; r < r1; r += 8) {
            const size_t o = (size_t)r*C + c;
            float g = __ldg(dY + o);
            if (relu && !(__ldg(Y + o) > 0.f)) g = 0.f;
            s += g;
            q = fmaf(g, (__ldg(X + o) - m)*rs, q);
        }
    }
    red[0][ty][tx] = s;
    red[1][ty][tx] = q;
    __syncthreads();
    if (ty == 0 && c < C) {
        double ds = 0., dq = 0.;
#pragma unroll
        for (int k = 0; k < 8; ++k) { ds += red[0][k][tx]; dq += red[1][k][tx]; }
        atomicAdd(sums + c, ds);
        atomicAdd(sums + C + c, dq);
    }
}

// dx = gamma*rstd*(dz - dbeta/M - xhat*dgamma/M); dres = dz (nullable); also emits dgamma / dbeta as floats (block 0).
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(long long total4, long long M, int C, const float* __restrict__ dY,
                                                           const float* __restrict__ Y, const float* __restrict__ X,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, const double* __restrict__ sums, int relu,
                                                           float* __restrict__ dX, float* __restrict__ dRes, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int accumulate) {
    const int c4 = C >> 2;
    const float invM = 1.f/(float)M;
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)sums[c];
            dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)sums[C + c];
        }
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < (unsigned)total4; i += gridDim.x*blockDim.x) {  // 32-bit index math (count < 2^31, checked on the host)
        const int c = (int)(i % c4)*4;
        float4 g = __ldg((const float4*)dY + i);
        if (relu) {
            const float4 y = __ldg((const float4*)Y + i);
            g.x = y.x > 0.f ? g.x : 0.f; g.y = y.y > 0.f ? g.y : 0.f; g.z = y.z > 0.f ? g.z : 0.f; g.w = y.w > 0.f ? g.w : 0.f;
        }
        if (dRes) ((float4*)dRes)[i] = g;
        const float4 x = __ldg((const float4*)X + i), m = __ldg((const float4*)(mean + c)), r = __ldg((const float4*)(rstd + c));
        const float4 ga = __ldg((const float4*)(gamma + c));
        const float gv[4] = {g.x, g.y, g.z, g.w}, xv[4] = {x.x, x.y, x.z, x.w}, mv[4] = {m.x, m.y, m.z, m.w}, rv[4] = {r.x, r.y, r.z, r.w},
                    gav[4] = {ga.x, ga.y, ga.z, ga.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float xh = (xv[k] - mv[k])*rv[k];
            o[k] = gav[k]*rv[k]*(gv[k] - (float)sums[c + k]*invM - xh*(float)sums[C + c + k]*invM);
        }
        ((float4*)dX)[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// ---- batched 4x4 inverse (intrinsics K -> K^-1, backward poses T -> T^-1) ------------------------------------------------------
// Reference: `K.inverse()` in ViewSynth.forward (src/tools/geometry.py:383) and `T.inverse()` (src/core/trainer.py:253), which
// ATen sends to a batched LU in cuSOLVER/MAGMA — host-synchronising and not CUDA-graph capturable. One thread per matrix:
// Gauss-Jordan with partial pivoting in double, rounded once to fp32.
__global__ void inv4x4_kernel(int n, const float* __restrict__ A, float* __restrict__ B) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) { a[r][c] = (double)A[i*16 + r*4 + c]; a[r][4 + c] = r == c ? 1. : 0.; }
#pragma unroll
    for (int col = 0; col < 4; ++col) {
        int piv = col;
#pragma unroll
        for (int r = col + 1; r < 4; ++r) if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r == piv && piv != col) {
#pragma unroll
                for (int c = 0; c < 8; ++c) { const double t = a[col][c]; a[col][c] = a[r][c]; a[r][c] = t; }
            }
        const double inv = 1.0/a[col][col];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[col][c] *= inv;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (r != col) {
                const double f = a[r][col];
#pragma unroll
                for (int c = 0; c < 8; ++c) a[r][c] -= f*a[col][c];
            }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) B[i*16 + r*4 + c] = (float)a[r][4 + c];
}

// gA = -B^T gB B^T  with B = A^-1.
__global__ void inv4x4_bwd_kernel(int n, const float* __restrict__ B, const float* __restrict__ gB, float* __restrict__ gA) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    float b[16], g[16], t[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { b[k] = B[i*16 + k]; g[k] = gB[i*16 + k]; }
#pragma unroll
    for (int r = 0; r < 4; ++r)      // t = B^T g
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) v = fmaf(b[k*4 + r], g[k*4 + c], v);
            t[r*4 + c] = v;
        }
#pragma unroll
    for (int r = 0; r < 4; ++r)      // gA = -t B^T
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) v = fmaf(t[r*4 + k], b[c*4 + k], v);
            gA[i*16 + r*4 + c] = -v;
        }
}

// ---- 3x3 stride-2 max-pool (padding 1) of the ResNet stem, channels-last; 4 channels per thread -------------------------------
// Reference: `maxpool` of the timm ResNet encoder (src/networks/pose.py:40, depth.py:97). The forward stores the winning tap
// (0..8, first maximum in row-major window order, as ATen) per output element; the backward is a gather over the <= 4 windows
// that contain an input pixel (deterministic, no atomics).
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(int N, int H, int W, int C, int P, int Q, const float* __restrict__ x,
                                                          float* __restrict__ y, uint8_t* __restrict__ idx) {
    const int c4 = C >> 2;
    const long long total = (long long)N*P*Q*c4;
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x*blockDim.x) {  // 32-bit index math (count < 2^31, checked on the host)
        const int c = (int)(i % c4)*4;
        unsigned r = i/c4;
        const int q = (int)(r % Q); r /= Q;
        const int p = (int)(r % P);
        const int n = (int)(r/P);
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        uchar4 bi = make_uchar4(0, 0, 0, 0);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int yy = 2*p - 1 + dy;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int xx = 2*q - 1 + dx;
                if (xx < 0 || xx >= W) continue;
                const float4 v = __ldg((const float4*)(x + ((size_t)(n*H + yy)*W + xx)*C + c));
                const unsigned char t = (unsigned char)(dy*3 + dx);
                if (v.x > best.x) { best.x = v.x; bi.x = t; }
                if (v.y > best.y) { best.y = v.y; bi.y = t; }
                if (v.z > best.z) { best.z = v.z; bi.z = t; }
                if (v.w > best.w) { best.w = v.w; bi.w = t; }
            }
        }
        ((float4*)y)[i] = best;
        ((uchar4*)idx)[i] = bi;
    }
}

__global__ void __launch_bounds__(256) maxpool_bwd_kernel(int N, int H, int W, int C, int P, int Q, const float* __restrict__ dy,
                                                          const uint8_t* __restrict__ idx, float* __restrict__ dx) {
    const int c4 = C >> 2;
    const long long total = (long long)N*H*W*c4;
    for (unsigned i = blockIdx.x*blockDim.x + threadIdx.x; i < (unsigned)total; i += gridDim.x*blockDim.x) {  // 32-bit index math (count < 2^31, checked on the host)
        const int c = (int)(i % c4)*4;
        unsigned r = i/c4;
        const int xx = (int)(r % W); r /= W;
        const int yy = (int)(r % H);
        const int n = (int)(r/H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        // windows p with 2p-1 <= yy <= 2p+1
        for (int p = (yy > 0 ? (yy - 1 + 1)/2 : 0); p <= (yy + 1)/2 && p < P; ++p) {
            const int ty = yy - (2*p - 1);
            for (int q = (xx > 0 ? (xx - 1 + 1)/2 : 0); q <= (xx + 1)/2 && q < Q; ++q) {
                const int t = ty*3 + (xx - (2*q - 1));
                const size_t o = ((size_t)(n*P + p)*Q + q)*C + c;
                const uchar4 k = *(const uchar4*)(idx + o);
                const float4 g = __ldg((const float4*)(dy + o));
                if (k.x == t) acc.x += g.x;
                if (k.y == t) acc.y += g.y;
                if (k.z == t) acc.z += g.z;
                if (k.w == t) acc.w += g.w;
            }
        }
        ((float4*)dx)[i] = acc;
    }
}

static int rows_per_block(long long M, int C) {
    // ~4 waves of 148 SMs x 8 resident blocks, at least 64 rows per block.
    const long long col_blocks = (C + 31)/32;
    long long want = (4ll*148*8 + col_blocks - 1)/col_blocks;
    long long rpb = (M + want - 1)/want;
    if (rpb < 64) rpb = 64;
    if (stv_deterministic()) rpb = M < 0x7fffffffll ? M : 0x7fffffffll;   // one row block: one contributor per column
    return (int)rpb;
}

// ---------------------------------------------------------------------------------------------------------------------
// ConvNeXt block, parameter gradients behind the layer-scale: with out = res + gamma * (h W2^T + b2), G = g^T h (C x Hd) and
// gs = column sums of g (C):  dW2 += gamma[:,None] * G;  db2 += gamma * gs;  dgamma += sum_j W2[c,j] G[c,j] + b2 * gs.
// One warp per output channel c (one launch instead of five ATen launches per block). w2g = gamma[:,None] * W2 (forward fold).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ls_tail_kernel(int C, int Hd, const float* __restrict__ G, const float* __restrict__ w2,
                                                      const float* __restrict__ b2, const float* __restrict__ gamma,
                                                      const float* __restrict__ gs, float* __restrict__ dw2, float* __restrict__ db2,
                                                      float* __restrict__ dgamma) {
    const int c = blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    const float ga = __ldg(gamma + c);
    const float* g = G + (size_t)c*Hd;
    const float* w = w2 + (size_t)c*Hd;
    float* d = dw2 + (size_t)c*Hd;
    float dot = 0.f;
    for (int j = lane; j < Hd; j += 32) {
        const float gv = g[j];
        dot = fmaf(__ldg(w + j), gv, dot);
        d[j] = fmaf(ga, gv, d[j]);
    }
    dot = warp_sum(dot);
    if (lane == 0) {
        const float s = __ldg(gs + c);
        db2[c] = fmaf(ga, s, db2[c]);
        dgamma[c] += fmaf(__ldg(b2 + c), s, dot);
    }
}

__global__ void __launch_bounds__(256) rowscale_kernel(long long n, int Hd, const float* __restrict__ w, const float* __restrict__ gamma,
                                                       float* __restrict__ out) {
    const long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) out[i] = w[i]*__ldg(gamma + i/Hd);
}

}  // namespace stv

using namespace stv;

extern "C" int stv_grad_pull(int N, int H, int W, int C, const float* src, int Cs, int c_off, int pad, int pool, float* dst, int accumulate,
                             void* stream) {
    STV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && src && dst, "stv_grad_pull: empty tensor / null pointer");
    STV_REQUIRE(C % 4 == 0 && Cs % 4 == 0 && c_off % 4 == 0 && c_off + C <= Cs, "stv_grad_pull: channel slice [%d, %d) of %d must be 4-aligned", c_off, c_off + C, Cs);
    STV_REQUIRE((pool == 1 || pool == 2) && pad >= 0 && pad < H*pool && pad < W*pool, "stv_grad_pull: bad pool / pad");
    const long long total = (long long)N*H*W*(C/4);
    STV_REQUIRE(total < (1ll << 31) && (long long)N*(H*pool + 2*pad)*(W*pool + 2*pad)*(Cs/4) < (1ll << 31), "stv_grad_pull: tensor too large");
    const int blocks = (int)((total + 255)/256 < 148ll*32 ? (total + 255)/256 : 148ll*32);
    grad_pull_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(N, H, W, C, src, Cs, c_off, pad, pool, dst, accumulate);
    count_launch();
    return check_launch("stv_grad_pull");
}

extern "C" int stv_act_bwd(long long M, int C, const float* dA, const float* Y, int act, float* dZ, float* dbias, void* stream) {
    STV_REQUIRE(M > 0 && C > 0 && dA && Y && dZ, "stv_act_bwd: empty tensor / null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (colred4_ok(C) && (((uintptr_t)dA | (uintptr_t)Y | (uintptr_t)dZ) & 15) == 0) {
        dim3 grid; int rpb;
        colred4_grid(M, C, grid, rpb);
        switch (act) {
            case STV_ACT_RELU: colred4_kernel<OpActBwd<STV_ACT_RELU>, false><<<grid, 256, 0, st>>>(M, C, rpb, {dA, Y, dZ}, dbias, nullptr); break;
            case STV_ACT_ELU: colred4_kernel<OpActBwd<STV_ACT_ELU>, false><<<grid, 256, 0, st>>>(M, C, rpb, {dA, Y, dZ}, dbias, nullptr); break;
            case STV_ACT_SIGMOID: colred4_kernel<OpActBwd<STV_ACT_SIGMOID>, false><<<grid, 256, 0, st>>>(M, C, rpb, {dA, Y, dZ}, dbias, nullptr); break;
            case STV_ACT_GELU: colred4_kernel<OpActBwd<STV_ACT_GELU>, false><<<grid, 256, 0, st>>>(M, C, rpb, {dA, Y, dZ}, dbias, nullptr); break;
            case STV_ACT_NONE: colred4_kernel<OpActBwd<STV_ACT_NONE>, false><<<grid, 256, 0, st>>>(M, C, rpb, {dA, Y, dZ}, dbias, nullptr); break;
            default: STV_REQUIRE(false, "stv_act_bwd: unknown activation %d", act);
        }
        count_launch();
        return check_launch("stv_act_bwd");
    }
    const int rpb = rows_per_block(M, C);
    const dim3 grid((C + 31)/32, (unsigned)((M + rpb - 1)/rpb));
    switch (act) {
        case STV_ACT_NONE: act_bwd_kernel<STV_ACT_NONE><<<grid, 256, 0, st>>>(M, C, dA, Y, dZ, dbias, rpb); break;
        case STV_ACT_RELU: act_bwd_kernel<STV_ACT_RELU><<<grid, 256, 0, st>>>(M, C, dA, Y, dZ, dbias, rpb); break;
        case STV_ACT_GELU: act_bwd_kernel<STV_ACT_GELU><<<grid, 256, 0, st>>>(M, C, dA, Y, dZ, dbias, rpb); break;
        case STV_ACT_ELU: act_bwd_kernel<STV_ACT_ELU><<<grid, 256, 0, st>>>(M, C, dA, Y, dZ, dbias, rpb); break;
        case STV_ACT_SIGMOID: act_bwd_kernel<STV_ACT_SIGMOID><<<grid, 256, 0, st>>>(M, C, dA, Y, dZ, dbias, rpb); break;
        default: STV_REQUIRE(false, "stv_act_bwd: unknown activation %d", act);
    }
    count_launch();
    return check_launch("stv_act_bwd");
}

extern "C" int stv_colsum(long long M, int C, long long ld, const float* X, float* out, void* stream) {
    STV_REQUIRE(M > 0 && C > 0 && ld >= C && X && out, "stv_colsum: empty tensor / null pointer");
    if (ld == C && colred4_ok(C) && ((uintptr_t)X & 15) == 0) {
        dim3 grid; int rpb4;
        colred4_grid(M, C, grid, rpb4);
        colred4_kernel<OpColSum, false><<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, rpb4, {X}, out, nullptr);
        count_launch();
        return check_launch("stv_colsum");
    }
    const int rpb = rows_per_block(M, C);
    const dim3 grid((C + 31)/32, (unsigned)((M + rpb - 1)/rpb));
    colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, ld, X, out, rpb);
    count_launch();
    return check_launch("stv_colsum");
}

/* ---- BatchNorm (train mode) ---- */
extern "C" size_t stv_bn_workspace_bytes(int C) { return (size_t)2*C*sizeof(double); }

extern "C" int stv_bn_fwd(long long M, int C, const float* x, const float* gamma, const float* beta, const float* res, int relu, float eps,
                          float momentum, float* y, float* mean, float* rstd, float* run_mean, float* run_var, void* ws, size_t ws_bytes,
                          void* stream) {
    STV_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && x && gamma && beta && y && mean && rstd, "stv_bn_fwd: bad arguments (C must be a multiple of 4)");
    STV_REQUIRE(ws && ws_bytes >= stv_bn_workspace_bytes(C), "stv_bn_fwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* sums = (double*)ws;
    if (cudaMemsetAsync(sums, 0, stv_bn_workspace_bytes(C), st) != cudaSuccess) return check_launch("stv_bn_fwd(memset)");
    if (colred4_ok(C)) {
        dim3 grid; int rpb;
        colred4_grid(M, C, grid, rpb);
        colred4_kernel<OpBnStats, true><<<grid, 256, 0, st>>>(M, C, rpb, {x}, sums, sums + C);
    } else {
        const int rpb = rows_per_block(M, C);
        const dim3 grid((C + 31)/32, (unsigned)((M + rpb - 1)/rpb));
        bn_stats_kernel<<<grid, 256, 0, st>>>(M, C, x, sums, rpb);
    }
    bn_finalize_kernel<<<(C + 127)/128, 128, 0, st>>>(M, C, sums, eps, momentum, mean, rstd, run_mean, run_var);
    const long long total4 = M*(C/4);
    STV_REQUIRE(total4 < (1ll << 31), "stv_bn_fwd: tensor too large");
    const int blocks = (int)((total4 + 255)/256 < 148ll*16 ? (total4 + 255)/256 : 148ll*16);
    bn_apply_kernel<<<blocks, 256, 0, st>>>(total4, C, x, mean, rstd, gamma, beta, res, relu, y);
    count_launch(3);
    return check_launch("stv_bn_fwd");
}

extern "C" int stv_bn_bwd(long long M, int C, const float* dy, const float* y, const float* x, const float* mean, const float* rstd,
                          const float* gamma, int relu, float* dx, float* dres, float* dgamma, float* dbeta, int accumulate, void* ws,
                          size_t ws_bytes, void* stream) {
    STV_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && dy && x && mean && rstd && gamma && dx && dgamma && dbeta && (!relu || y),
                "stv_bn_bwd: bad arguments (C must be a multiple of 4)");
    STV_REQUIRE(ws && ws_bytes >= stv_bn_workspace_bytes(C), "stv_bn_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* sums = (double*)ws;
    if (cudaMemsetAsync(sums, 0, stv_bn_workspace_bytes(C), st) != cudaSuccess) return check_launch("stv_bn_bwd(memset)");
    if (colred4_ok(C)) {
        dim3 grid; int rpb;
        colred4_grid(M, C, grid, rpb);
        colred4_kernel<OpBnBwdReduce, true><<<grid, 256, 0, st>>>(M, C, rpb, {dy, y, x, mean, rstd, relu}, sums, sums + C);
    } else {
        const int rpb = rows_per_block(M, C);
        const dim3 grid((C + 31)/32, (unsigned)((M + rpb - 1)/rpb));
        bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(M, C, dy, y, x, mean, rstd, relu, sums, rpb);
    }
    const long long total4 = M*(C/4);
    STV_REQUIRE(total4 < (1ll << 31), "stv_bn_bwd: tensor too large");
    const int blocks = (int)((total4 + 255)/256 < 148ll*16 ? (total4 + 255)/256 : 148ll*16);
    bn_bwd_apply_kernel<<<blocks, 256, 0, st>>>(total4, M, C, dy, y, x, mean, rstd, gamma, sums, relu, dx, dres, dgamma, dbeta, accumulate);
    count_launch(2);
    return check_launch("stv_bn_bwd");
}

extern "C" int stv_inv4x4(int n, const float* A, float* B, void* stream) {
    STV_REQUIRE(n > 0 && A && B, "stv_inv4x4: empty batch / null pointer");
    inv4x4_kernel<<<(n + 63)/64, 64, 0, (cudaStream_t)stream>>>(n, A, B);
    count_launch();
    return check_launch("stv_inv4x4");
}

extern "C" int stv_inv4x4_bwd(int n, const float* B, const float* gB, float* gA, void* stream) {
    STV_REQUIRE(n > 0 && B && gB && gA, "stv_inv4x4_bwd: empty batch / null pointer");
    inv4x4_bwd_kernel<<<(n + 63)/64, 64, 0, (cudaStream_t)stream>>>(n, B, gB, gA);
    count_launch();
    return check_launch("stv_inv4x4_bwd");
}

extern "C" int stv_maxpool3x3s2_fwd(int N, int H, int W, int C, const float* x, float* y, uint8_t* idx, void* stream) {
    STV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && x && y && idx, "stv_maxpool3x3s2_fwd: bad arguments (C must be a multiple of 4)");
    const int P = (H - 1)/2 + 1, Q = (W - 1)/2 + 1;
    const long long total = (long long)N*P*Q*(C/4);
    STV_REQUIRE((long long)N*H*W*(C/4) < (1ll << 31), "stv_maxpool3x3s2_fwd: tensor too large");
    const int blocks = (int)((total + 255)/256 < 148ll*16 ? (total + 255)/256 : 148ll*16);
    maxpool_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(N, H, W, C, P, Q, x, y, idx);
    count_launch();
    return check_launch("stv_maxpool3x3s2_fwd");
}

extern "C" int stv_maxpool3x3s2_bwd(int N, int H, int W, int C, const float* dy, const uint8_t* idx, float* dx, void* stream) {
    STV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0 && dy && dx && idx, "stv_maxpool3x3s2_bwd: bad arguments (C must be a multiple of 4)");
    const int P = (H - 1)/2 + 1, Q = (W - 1)/2 + 1;
    const long long total = (long long)N*H*W*(C/4);
    STV_REQUIRE(total < (1ll << 31), "stv_maxpool3x3s2_bwd: tensor too large");
    const int blocks = (int)((total + 255)/256 < 148ll*16 ? (total + 255)/256 : 148ll*16);
    maxpool_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(N, H, W, C, P, Q, dy, idx, dx);
    count_launch();
    return check_launch("stv_maxpool3x3s2_bwd");
}

extern "C" int stv_ls_tail(int C, int Hd, const float* G, const float* w2, const float* b2, const float* gamma, const float* gs,
                           float* dw2, float* db2, float* dgamma, void* stream) {
    STV_REQUIRE(C > 0 && Hd > 0, "stv_ls_tail: bad shape (C=%d Hd=%d)", C, Hd);
    STV_REQUIRE(G && w2 && b2 && gamma && gs && dw2 && db2 && dgamma, "stv_ls_tail: NULL pointer");
    ls_tail_kernel<<<(C + 7)/8, 256, 0, (cudaStream_t)stream>>>(C, Hd, G, w2, b2, gamma, gs, dw2, db2, dgamma);
    count_launch();
    return check_launch("ls_tail_kernel");
}

extern "C" int stv_rowscale(int C, int Hd, const float* w, const float* gamma, float* out, void* stream) {
    STV_REQUIRE(C > 0 && Hd > 0 && w && gamma && out, "stv_rowscale: bad arguments");
    const long long n = (long long)C*Hd;
    rowscale_kernel<<<(unsigned)((n + 255)/256), 256, 0, (cudaStream_t)stream>>>(n, Hd, w, gamma, out);
    count_launch();
    return check_launch("rowscale_kernel");
}
