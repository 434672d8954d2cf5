// Developer probe for round 2 (NOT part of libstv): can a K-major, 128-byte-swizzled shared-memory slab written by ONE TMA box be
// consumed by tcgen05.mma through descriptors whose start address is shifted by s rows (s * 128 B)? If yes, a 3x3 convolution can
// stage a (rows + halo) input tile once and read its nine taps as row-shifted views instead of nine im2col TMA boxes (DESIGN.md,
// round-2 plan item 2). For s = 0..8 the kernel computes  D_s[128, 32] = A[s : s+128, 0:32] * B[0:32, 0:32]^T  from a 136-row slab
// and the host compares with the CPU product, for both settings of the descriptor's base-offset field ([49,52)):
//   mode 0: base offset 0;   mode 1: base offset = (start address >> 7) & 7  (PTX: "matrix base offset" for unaligned starts).
// STATUS: compiles for sm_100a; NOT YET RUN. usage: timeout 20 ./umma_rowshift_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr, int mode) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((16u >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    if (mode == 1) d |= (uint64_t)((addr >> 7) & 7u) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr int ROWS = 136, NS = 9;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ D,
                                            int mode) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t* sa = smem;                       // 136 rows x 128 B
    uint8_t* sb = smem + 18*1024;             // 32 rows x 128 B, 1024-aligned
    __shared__ uint64_t full, done;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&full, 1); mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full)), "r"((uint32_t)((ROWS + 32)*128)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sa)), "l"(&tmA), "r"(smem_u32(&full)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sb)), "l"(&tmB), "r"(smem_u32(&full)), "r"(0), "r"(0) : "memory");
        mbar_wait(&full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int s = 0; s < NS; ++s)
            for (int k8 = 0; k8 < 4; ++k8) {
                const uint64_t da = desc_kmajor(smem_u32(sa) + s*128 + k8*32, mode), db = desc_kmajor(smem_u32(sb) + k8*32, mode);
                const uint32_t accum = k8 > 0 ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem + (uint32_t)(s*32)), "l"(da), "l"(db), "r"(IDESC), "r"(accum) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    }
    __syncwarp();
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int s = 0; s < NS; ++s) {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                       "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(tmem + ((uint32_t)(warp*32) << 16) + (uint32_t)(s*32)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[((size_t)s*128 + warp*32 + lane)*32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static int make_map(CUtensorMap* tm, const float* base, int rows, int box_rows) {
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return 1;
    const cuuint64_t dims[2] = {32, (cuuint64_t)rows}, strides[1] = {32*4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows}, es[2] = {1u, 1u};
    return (int)((Fn)ptr)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
    std::vector<float> A((size_t)ROWS*32), B(32*32), D((size_t)NS*128*32);
    srand(2);
    for (auto& v : A) v = (float)((rand() % 17) - 8)/8.f;
    for (auto& v : B) v = (float)((rand() % 17) - 8)/8.f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size()*4); cudaMalloc(&dB, B.size()*4); cudaMalloc(&dD, D.size()*4);
    cudaMemcpy(dA, A.data(), A.size()*4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size()*4, cudaMemcpyHostToDevice);
    CUtensorMap tmA, tmB;
    if (make_map(&tmA, dA, ROWS, ROWS) || make_map(&tmB, dB, 32, 32)) { printf("tensor map encode failed\n"); return 1; }
    const size_t smem = 18*1024 + 4*1024 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dD, 0xFF, D.size()*4);
        probe<<<1, 128, smem>>>(tmA, tmB, dD, mode);
        const cudaError_t e = cudaDeviceSynchronize();
        printf("base-offset mode %d: run=%s |", mode, cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("\n"); return 0; }
        cudaMemcpy(D.data(), dD, D.size()*4, cudaMemcpyDeviceToHost);
        for (int s = 0; s < NS; ++s) {
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 32; ++n) {
                    double r = 0;
                    for (int k = 0; k < 32; ++k) r += (double)A[(size_t)(s + m)*32 + k]*B[(size_t)n*32 + k];
                    if (!(fabs(r - D[((size_t)s*128 + m)*32 + n]) <= 1e-4)) ++bad;
                }
            printf(" shift %d: %d bad", s, bad);
        }
        printf("\n");
    }
    return 0;
}
