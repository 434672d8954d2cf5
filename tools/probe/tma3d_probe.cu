// Developer probe: which 3-D TMA box shapes / start coordinates are legal on sm_100a. usage: probe bw bh bp x y z
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, int x, int y, int z, uint32_t bytes, float* out, int n) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* buf = (float*)raw;
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(buf)), "l"(&tm), "r"(smem_u32(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    __syncthreads();
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
    int bw = atoi(argv[1]), bh = atoi(argv[2]), bp = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]), z = atoi(argv[6]);
    const int W = 192, H = 128, P = 18;
    float* d; cudaMalloc(&d, (size_t)W*H*P*4);
    float* h = (float*)malloc((size_t)W*H*P*4);
    for (int i = 0; i < W*H*P; ++i) h[i] = (float)i;
    cudaMemcpy(d, h, (size_t)W*H*P*4, cudaMemcpyHostToDevice);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* ptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t dims[3] = {W, H, P}, strides[2] = {W*4, (cuuint64_t)W*H*4};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bp}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)ptr)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %d %d %d at %d %d %d: encode=%d ", bw, bh, bp, x, y, z, (int)r);
    if (r) { printf("\n"); return 0; }
    int n = bw*bh*bp; float* out; cudaMalloc(&out, n*4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, n*4);
    k<<<1, 128, n*4>>>(tm, x, y, z, n*4, out, n);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        float* o = (float*)malloc(n*4); cudaMemcpy(o, out, n*4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < bp; ++c) for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
            int X = x + i, Y = y + j, Z = z + c;
            float want = (X < 0 || X >= W || Y < 0 || Y >= H || Z < 0 || Z >= P) ? 0.f : (float)((Z*H + Y)*W + X);
            if (o[(c*bh + j)*bw + i] != want) ++bad;
        }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
