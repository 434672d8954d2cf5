// Developer probe for round 2 (NOT part of libstv): the smallest cta_group::2 TF32 GEMM — one CTA pair computes
//   C[256, N] = A[256, K] * B[N, K]^T      (K-major operands, 128-byte swizzle, fp32 accumulate in TMEM)
// with each CTA TMA-loading ITS 128 rows of A and ITS N/2 rows of B, the leader issuing tcgen05.mma.cta_group::2 (M = 256) and
// committing with .multicast::cluster to both CTAs. It answers, before the product kernel is touched:
//   (1) do our UMMA descriptors / instruction descriptor carry over to M = 256, (2) which CTA holds which half of B,
//   (3) does the leader-barrier transaction accounting (peer CTA's TMA -> leader's mbarrier) behave as CUTLASS documents.
//   (4) with `stages` < K/32: does a smem ring shared by the pair work when the leader's multicast commit releases stage s in BOTH
//       CTAs (each producer waits on its own empty[s]; the leader alone posts expect_tx for the bytes of both CTAs).
// STATUS: compiles for sm_100a; NOT YET RUN (written after the round-1 GPU budget was spent). Expected output: "max |err|" ~1e-3
// (TF32-exact inputs: expect 0) and "mismatches=0". usage: timeout 20 ./cta2_gemm_probe [N=128] [K=64] [stages=K/32]   (ALWAYS under a timeout: a
// wrong barrier protocol hangs the pair)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
// 2-SM TMA load: the transaction bytes land on the LEADER CTA's mbarrier (peer bit of the shared::cluster address cleared),
// cute::SM100_TMA_2SM_LOAD_2D.
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t slab, int k8) {   // as stv_tc.cuh
    uint64_t d = 0;
    d |= (uint64_t)(((slab + k8*32) >> 4) & 0x3FFFu);
    d |= (uint64_t)((16u >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int BK = 32, THREADS = 128, MAXST = 8;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS)
cta2_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C, int N, int K, int stages) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int kb_total = K/BK, halfN = N/2;
    const uint32_t a_bytes = 128*BK*4, b_bytes = (uint32_t)halfN*BK*4, stage = a_bytes + b_bytes;
    __shared__ uint64_t full[MAXST], empty[MAXST], done;
    __shared__ uint32_t tmem_slot;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // same warp id in both CTAs, same smem slot (cute::TMEM::Allocator2Sm preconditions)
        const uint32_t cols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : 256u;
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();   // barriers of the leader are initialised before the peer's TMA signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    // Producer (thread 0 of BOTH CTAs): this CTA's 128 rows of A and N/2 rows of B per k-block into ring stage kb % stages.
    // The stage is free again once the leader's multicast commit has arrived on THIS CTA's empty[s].
    if (warp == 0 && lane == 0) {
        for (int kb = 0; kb < kb_total; ++kb) {
            const int s = kb % stages;
            mbar_wait(&empty[s], (uint32_t)((kb/stages) & 1) ^ 1u);   // first pass: passes immediately (fresh barrier, parity 1)
            if (rank == 0) mbar_expect_tx(&full[s], 2*stage);          // bytes of BOTH CTAs arrive on the leader's barrier
            uint8_t* a = smem + (size_t)s*stage;
            tma_load_2d_2sm(a, &tmA, &full[s], kb*BK, (int)rank*128);
            tma_load_2d_2sm(a + a_bytes, &tmB, &full[s], kb*BK, (int)rank*halfN);
        }
    }
    // MMA issuer (thread 32 of the LEADER): separate warp, so that the producer above can run ahead through the ring
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = idesc_tf32(256, N);
        for (int kb = 0; kb < kb_total; ++kb) {
            const int s = kb % stages;
            mbar_wait(&full[s], (uint32_t)(kb/stages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a = smem_u32(smem + (size_t)s*stage), b = a + a_bytes;
            for (int k8 = 0; k8 < BK/8; ++k8) {
                const uint64_t da = umma_desc_kmajor(a, k8), db = umma_desc_kmajor(b, k8);
                const uint32_t accum = (kb > 0 || k8 > 0) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
            }
            // release stage s in BOTH CTAs once these MMAs have read it
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&empty[s])), "h"((uint16_t)0b11) : "memory");
        }
        // arrive on `done` in BOTH CTAs once every MMA above has completed
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&done)), "h"((uint16_t)0b11) : "memory");
    }
    __syncwarp();
    // epilogue (both CTAs): warp w drains TMEM lanes 32w..32w+31 = rows rank*128 + 32w + lane
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = (int)rank*128 + warp*32 + lane;
    for (int c = 0; c < N; c += 32) {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                       "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(tmem + ((uint32_t)(warp*32) << 16) + (uint32_t)c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) C[(size_t)row*N + c + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();   // both CTAs are done with TMEM (and the peer's smem) before it is released
    if (warp == 0) {
        const uint32_t cols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : 256u;
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols) : "memory");
    }
}

static int make_map(CUtensorMap* tm, const float* base, int rows, int cols, int box_rows) {
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) return 1;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols*4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows}, es[2] = {1u, 1u};
    return (int)((Fn)ptr)(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 128, K = argc > 2 ? atoi(argv[2]) : 64, M = 256;
    int stages = argc > 3 ? atoi(argv[3]) : K/BK;
    if (stages > MAXST) stages = MAXST;
    if (N % 32 || N > 256 || K % BK || stages < 1) { printf("need N %% 32 == 0, N <= 256, K %% 32 == 0, stages >= 1\n"); return 1; }
    std::vector<float> A((size_t)M*K), B((size_t)N*K), Cc((size_t)M*N);
    srand(1);
    auto tf32ish = [] { return (float)((rand() % 17) - 8)/8.f; };   // exactly representable in TF32: the product must be exact
    for (auto& v : A) v = tf32ish();
    for (auto& v : B) v = tf32ish();
    float *dA, *dB, *dC;
    cudaMalloc(&dA, A.size()*4); cudaMalloc(&dB, B.size()*4); cudaMalloc(&dC, Cc.size()*4);
    cudaMemcpy(dA, A.data(), A.size()*4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size()*4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xFF, Cc.size()*4);
    CUtensorMap tmA, tmB;
    if (make_map(&tmA, dA, M, K, 128) || make_map(&tmB, dB, N, K, N/2)) { printf("tensor map encode failed\n"); return 1; }
    const size_t smem = (size_t)stages*(128*BK*4 + (N/2)*BK*4) + 1024;
    cudaFuncSetAttribute(cta2_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cta2_gemm<<<2, THREADS, smem>>>(tmA, tmB, dC, N, K, stages);
    const cudaError_t e = cudaDeviceSynchronize();
    printf("M=256 N=%d K=%d stages=%d: run=%s ", N, K, stages, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        cudaMemcpy(Cc.data(), dC, Cc.size()*4, cudaMemcpyDeviceToHost);
        int bad = 0; double worst = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0;
                for (int k = 0; k < K; ++k) r += (double)A[(size_t)m*K + k]*B[(size_t)n*K + k];
                const double d = fabs(r - Cc[(size_t)m*N + n]);
                if (!(d <= 1e-4)) ++bad;
                if (d > worst || d != d) worst = d;
            }
        printf("max |err| = %.3g mismatches=%d", worst, bad);
    }
    printf("\n");
    return 0;
}
