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
    for (long long idx = blockIdx.x*(long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x*blockDim.x) {
        const int c = (int)(idx % c4)*4;
        long long r = idx/c4;
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

static int rows_per_block(long long M, int C) {
    // ~4 waves of 148 SMs x 8 resident blocks, at least 64 rows per block.
    const long long col_blocks = (C + 31)/32;
    long long want = (4ll*148*8 + col_blocks - 1)/col_blocks;
    long long rpb = (M + want - 1)/want;
    if (rpb < 64) rpb = 64;
    return (int)rpb;
}

}  // namespace stv

using namespace stv;

extern "C" int stv_grad_pull(int N, int H, int W, int C, const float* src, int Cs, int c_off, int pad, int pool, float* dst, int accumulate,
                             void* stream) {
    STV_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && src && dst, "stv_grad_pull: empty tensor / null pointer");
    STV_REQUIRE(C % 4 == 0 && Cs % 4 == 0 && c_off % 4 == 0 && c_off + C <= Cs, "stv_grad_pull: channel slice [%d, %d) of %d must be 4-aligned", c_off, c_off + C, Cs);
    STV_REQUIRE((pool == 1 || pool == 2) && pad >= 0 && pad < H*pool && pad < W*pool, "stv_grad_pull: bad pool / pad");
    const long long total = (long long)N*H*W*(C/4);
    const int blocks = (int)((total + 255)/256 < 148ll*32 ? (total + 255)/256 : 148ll*32);
    grad_pull_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(N, H, W, C, src, Cs, c_off, pad, pool, dst, accumulate);
    count_launch();
    return check_launch("stv_grad_pull");
}

extern "C" int stv_act_bwd(long long M, int C, const float* dA, const float* Y, int act, float* dZ, float* dbias, void* stream) {
    STV_REQUIRE(M > 0 && C > 0 && dA && Y && dZ, "stv_act_bwd: empty tensor / null pointer");
    const int rpb = rows_per_block(M, C);
    const dim3 grid((C + 31)/32, (unsigned)((M + rpb - 1)/rpb));
    cudaStream_t st = (cudaStream_t)stream;
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
    const int rpb = rows_per_block(M, C);
    const dim3 grid((C + 31)/32, (unsigned)((M + rpb - 1)/rpb));
    colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, C, ld, X, out, rpb);
    count_launch();
    return check_launch("stv_colsum");
}
