// Logging statistics of the training step in ONE launch pair and one device-to-host copy.
//
// The reference's `summarize_depth` (src/core/trainer.py:486-503) calls `.mean().item()` and `.std().item()` on every
// up-sampled disparity and depth map — 4 scales x 2 tensors x 2 statistics = 16 reduction launches pairs and 16 host
// synchronisations per logging step (`summarize_pose` / `summarize_K`, :505-529, add a dozen more on tiny tensors). Here the mean
// and the unbiased standard deviation (torch.std default) of up to STV_STATS_MAX tensors are produced by one multi-tensor
// kernel (per-block double partial sums, fixed order) and a finalize kernel; the host reads them back with a single copy when
// it actually logs. HBM-bound: every element is read once (float4), ~3 flops per element.
#include "stv_common.cuh"

namespace stv {

struct StatsParams {
    const float* ptr[STV_STATS_MAX];
    long long n[STV_STATS_MAX];
    int k;
};

constexpr int ST_THREADS = 256;

// grid = (blocks_per_tensor, k). partial[(t*gridDim.x + b)*2 + {0,1}] = sum, sum of squares about a per-tensor shift (the first
// element), which keeps the one-pass variance well conditioned for maps far from zero (depth in [0.1, 100]).
__global__ void __launch_bounds__(ST_THREADS) moments_kernel(StatsParams p, double* __restrict__ partial) {
    __shared__ double red[2][ST_THREADS/32];
    const int t = blockIdx.y;
    const float* __restrict__ x = p.ptr[t];
    const long long n = p.n[t];
    const float shift = __ldg(x);
    double s = 0.0, q = 0.0;
    const long long n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n/4 : 0;
    for (long long i = (long long)blockIdx.x*ST_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x*ST_THREADS) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float a = v.x - shift, b = v.y - shift, c = v.z - shift, d = v.w - shift;
        s += (double)((a + b) + (c + d));
        q += (double)(fmaf(a, a, b*b) + fmaf(c, c, d*d));
    }
    for (long long i = n4*4 + (long long)blockIdx.x*ST_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x*ST_THREADS) {
        const float a = __ldg(x + i) - shift;
        s += (double)a; q += (double)(a*a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = s; red[1][wid] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ss = 0.0, qq = 0.0;
#pragma unroll
        for (int w = 0; w < ST_THREADS/32; ++w) { ss += red[0][w]; qq += red[1][w]; }
        partial[((size_t)t*gridDim.x + blockIdx.x)*2 + 0] = ss;
        partial[((size_t)t*gridDim.x + blockIdx.x)*2 + 1] = qq;
    }
}

// out[t*2 + 0] = mean, out[t*2 + 1] = unbiased standard deviation (NaN for n == 1, as torch.std). One warp per tensor.
__global__ void moments_finalize_kernel(StatsParams p, int blocks, const double* __restrict__ partial, float* __restrict__ out) {
    const int t = blockIdx.x, lane = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int b = lane; b < blocks; b += 32) { s += partial[((size_t)t*blocks + b)*2]; q += partial[((size_t)t*blocks + b)*2 + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) {
        const double n = (double)p.n[t], m = s/n;
        const double var = (q - s*m)/(n - 1.0);
        out[t*2 + 0] = (float)(m + (double)__ldg(p.ptr[t]));
        out[t*2 + 1] = (float)sqrt(var > 0.0 ? var : (n > 1.0 ? 0.0 : NAN));
    }
}

constexpr int ST_BLOCKS = 148;   // per tensor: one block per SM

}  // namespace stv

using namespace stv;

extern "C" size_t stv_mean_std_workspace_bytes(int k) { return k > 0 ? (size_t)k*ST_BLOCKS*2*sizeof(double) : 0; }

extern "C" int stv_mean_std(int k, const float* const* tensors, const long long* counts, float* out, void* ws, size_t ws_bytes, void* stream) {
    STV_REQUIRE(k > 0 && k <= STV_STATS_MAX, "stv_mean_std: k must be in [1, %d] (got %d)", STV_STATS_MAX, k);
    STV_REQUIRE(tensors && counts && out, "stv_mean_std: NULL pointer");
    if (!ws || ws_bytes < stv_mean_std_workspace_bytes(k)) {
        set_error("stv_mean_std: workspace too small (%zu < %zu bytes)", ws_bytes, stv_mean_std_workspace_bytes(k));
        return STV_E_WORKSPACE;
    }
    StatsParams p{};
    p.k = k;
    for (int t = 0; t < k; ++t) {
        STV_REQUIRE(tensors[t] != nullptr && counts[t] > 0, "stv_mean_std: tensor %d is NULL or empty", t);
        p.ptr[t] = tensors[t]; p.n[t] = counts[t];
    }
    moments_kernel<<<dim3(ST_BLOCKS, k), ST_THREADS, 0, (cudaStream_t)stream>>>(p, (double*)ws);
    count_launch();
    if (int rc = check_launch("moments_kernel")) return rc;
    moments_finalize_kernel<<<k, 32, 0, (cudaStream_t)stream>>>(p, ST_BLOCKS, (const double*)ws, out);
    count_launch();
    return check_launch("moments_finalize_kernel");
}
