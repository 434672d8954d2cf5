// Blackwell (sm_100a) tensor-core plumbing for libstv: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (TMEM allocation,
// UMMA shared-memory / instruction descriptors, tcgen05.mma kind::tf32, tcgen05.ld) as thin inline-PTX wrappers.
//
// Operand staging convention used by every GEMM-shaped kernel in this library ("slabs"):
//   a slab is R rows x 128 bytes (32 fp32) of shared memory, 1024-byte aligned, swizzled as a TMA box {32 floats, R rows} is.
//   * K-major operand  (reduction index contiguous in memory): rows = M/N index, the 128 bytes = 32 consecutive k.
//       128-byte swizzle (16-byte chunk index XOR (row & 7); CU_TENSOR_MAP_SWIZZLE_128B; UMMA layout type SWIZZLE_128B).
//       descriptor: SBO = 1024 B (next 8-row group), one MMA (K = 8 tf32 = 32 B) advances the start address by 32 B.
//   * MN-major operand (M/N index contiguous in memory): rows = k index, the 128 bytes = 32 consecutive m (or n).
//       For 32-bit (tf32) data the tensor core accepts ONE MN-major layout: "128B swizzle with a 32-byte atom"
//       (32-byte chunk index XOR (row & 3); TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B; UMMA layout type SWIZZLE_128B_BASE32B),
//       whose swizzle atom is 4 k-rows deep. descriptor: LBO = slab size (next 32 m), SBO = 512 B (next 4 k);
//       one MMA (K = 8) consumes 8 rows = advances the start address by 1024 B.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace stv { namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t globaltimer() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- mbarrier -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocking wait with a watchdog: a pipeline bug must surface as a trapped kernel (CUDA error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(32);  // back off: polling warps must not take issue slots from the warps they are waiting for
        if ((++spins & 0xFFFu) == 0) {
            const uint64_t now = globaltimer();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}

// Pure spin (mbarrier.try_wait already suspends the thread in hardware for a bounded time): for the single-thread producer and
// MMA-issuer roles, whose wake-up latency is on the critical path of the pipeline and who do not crowd anybody's issue slots.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0xFFFFu) == 0) {
            const uint64_t now = globaltimer();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}

// ---- proxies / fences -----------------------------------------------------------------------------------------------
// Makes generic-proxy shared-memory writes (st.shared / cp.async) visible to the async proxy (TMA, tcgen05.mma).
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
// 2-D tiled load: box origin (x = inner/contiguous coordinate, y = row) -> smem, completes `bar` by byte count.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}

// 3-D tiled load: box origin (x, y, z), x the contiguous coordinate; out-of-bounds elements (negative coordinates included) are
// zero-filled.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// im2col-mode load from a channels-last (N,H,W,C) tensor map: `pixelsPerColumn` base pixels starting at (n, h, w) — walking W,
// then H, then N inside the map's bounding box with its traversal strides — each read at the filter-tap offset (off_h, off_w),
// channels c .. c+31; pixels / channels outside the tensor are zero-filled. Rows land as a TMA box does (128 B, swizzled).
__device__ __forceinline__ void tma_load_im2col_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

// ---- cp.async (generic-proxy gather path of the implicit-GEMM loaders) ----------------------------------------------
// 16-byte copy; src_bytes = 0 writes zeros (padding) without touching `src`.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// Waits until at most `n` (0..7, warp-uniform) of this thread's most recent cp.async groups are still pending.
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
    switch (n) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        default: cp_async_wait<7>(); break;
    }
}

// ---- TMEM -----------------------------------------------------------------------------------------------------------
// Whole-warp calls. ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp receives lane (addr.lane + t), columns addr.col .. +31.
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors -----------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 | [46,48) version = 1
//   [49,52) base offset = 0 (slabs are 1024 B aligned) | [61,64) layout type: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
// K-major slab (rows = m/n, 128-byte swizzle): the MMA at k-offset `k8` (units of 8 tf32 = 32 B) within the 32-wide k-block.
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t slab, int k8) { return umma_desc(slab + k8*32, 16, 1024, 2); }
// MN-major slabs (rows = k, 32-byte-atom swizzle), `slab_bytes` apart for consecutive groups of 32 m/n.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t slab, int k8, uint32_t slab_bytes) {
    return umma_desc(slab + k8*1024, slab_bytes, 512, 1);
}
// Instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor bit layout):
//   [4,6) D format = 1 (F32) | [7,10) A format = 2 (TF32) | [10,13) B format = 2 | [15] A MN-major | [16] B MN-major
//   [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on `bar` once every previously issued tcgen05.mma of this thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One lane of a converged warp (PTX elect.sync): the compiler knows exactly one thread is active under this predicate, so values
// it feeds to uniform-register operands (UTMALDG / UTCHMMA descriptors) need no per-value election loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- lean single-issuer loops ----------------------------------------------------------------------------------------
// The producer and MMA-issuer loops run on ONE thread each and sit on the critical path of every k-block (4 MMAs = ~280 tensor
// cycles at M = N = 128): ~150 scalar instructions per k-block (runtime `%`, descriptor assembly, per-operand election loops that
// move divergent-code values into uniform registers) made the issuing thread, not the tensor pipe or L2, the bottleneck
// (profiles/r2_gemm_issue_bound.txt). These variants take 32-bit shared-memory addresses, so that the whole warp can run the loop
// control with warp-uniform values and one lane issues.
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin_s(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_s(bar, parity)) return;
    uint64_t t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait_s(bar, parity)) {
        if ((++spins & 0xFFFFu) == 0) {   // watchdog: a pipeline bug must trap, never hang the GPU
            const uint64_t now = globaltimer();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_s(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_s(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n,
                                                     uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
// L2 prefetch of a box a few k-blocks ahead of its load: the load then finds its lines in L2 (the 2-CTA/SM kernel has room for a
// 2-3 stage ring only, so its k-block period is the load latency over the ring depth — profiles/r2_gemm_timeline.txt).
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int x, int y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(m), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_prefetch_im2col_4d(const CUtensorMap* m, int c, int w, int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.im2col [%0, {%1, %2, %3, %4}], {%5, %6};"
                 ::"l"(m), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void umma_commit_s(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Descriptor of a slab at shared-memory address 0 (the caller adds address >> 4 into the low 14 bits: the ring lies below 256 KB)
// and the increment, in descriptor units, from one K = 8 MMA to the next.
__device__ __forceinline__ uint64_t umma_desc_template(bool mn_major, uint32_t slab_bytes) {
    return mn_major ? umma_desc(0, slab_bytes, 512, 1) : umma_desc(0, 16, 1024, 2);
}
__device__ __forceinline__ uint32_t umma_desc_kstep(bool mn_major) { return mn_major ? (1024u >> 4) : (32u >> 4); }

// ---- cta_group::2 (CTA pair = one 256-row MMA; tools/probe/cta2_gemm_probe.cu established the protocol on a B200) --------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA load: data lands in THIS CTA's shared memory, the transaction bytes on the LEADER CTA's mbarrier at the same offset
// (peer bit of the shared::cluster address cleared; cute::SM100_TMA_2SM_LOAD_2D).
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA's smem] * B[N columns: N/2 from each CTA's smem]^T, issued by ONE thread
// of the leader CTA.
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on `bar` (same offset) in BOTH CTAs of the pair once every previously issued MMA has completed.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)0b11) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm_s(uint32_t dst, const CUtensorMap* m, uint32_t bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(m), "r"(bar & 0xFEFFFFFFu), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_2sm_s(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h, int n,
                                                         uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(m), "r"(bar & 0xFEFFFFFFu), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_s(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)0b11) : "memory");
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// ---- epilogue stores ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}}  // namespace stv::tc
