// TF32 tensor-core GEMM for the network layers (tcgen05.mma kind::tf32, fp32 operands straight from HBM via TMA, fp32
// accumulators in TMEM) with fused epilogues. One kernel covers the forward, data-gradient and weight-gradient products of
// every Linear / 1x1-convolution layer by choosing, per operand, whether the reduction index is the contiguous one in memory
// ("K-major") or the strided one ("MN-major"):
//
//   C[M,N] (+)= epilogue( sum_k A[m,k] * B[n,k] )
//     forward   Y  = X  W^T      A = X  (M x in, K-major)      B = W (out x in, K-major)
//     dgrad     dX = dY W        A = dY (M x out, K-major)     B = W (out x in) read as MN-major (reduction over `out`)
//     wgrad     dW = dY^T X      A = dY read MN-major          B = X read MN-major (reduction over the M pixels), split-K
//
// Reference numerics being matched: fp32 storage with TF32 matmuls (`torch.set_float32_matmul_precision('high')`,
// src/core/trainer.py:30; cfg/default.yaml:171). Call sites replaced: the `nn.Linear` / 1x1 `nn.Conv2d` layers of the timm
// encoders built at src/networks/depth.py:97 / pose.py:40 and of the pose heads (src/networks/pose.py:46,75-106).
//
// Structure (one 128 x BN output tile per CTA, 320 threads):
//   warp 0    TMA producer: per 32-wide k-block, box loads of A and B slabs into a ring of smem stages (mbarrier expect_tx)
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases stages / signals the epilogue
//   warps 2-9 epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> shared-memory transpose -> bias / activation /
//             layer-scale / residual / activation-backward / column sums -> coalesced 128-bit global stores, or
//             red.global.add.v4 when accumulating (split-K, grad buffers)
#include <cstdlib>
#include <mutex>

#include "stv_common.cuh"
#include "stv_tc.cuh"
#include "stv_epi.cuh"
#include "stv_gemm.cuh"

namespace stv {

__global__ void __launch_bounds__(GEMM_THREADS)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);  // slabs need 1024-byte alignment (128 B swizzle atom)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_bytes = p.bn*GEMM_BK*4;
    const int stage_bytes = GEMM_A_BYTES + b_bytes;
    uint64_t* full = (uint64_t*)(smem + (size_t)p.stages*stage_bytes);
    uint64_t* empty = full + GEMM_MAX_STAGES;
    uint64_t* tmem_full = empty + GEMM_MAX_STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

    // Column tiles are the fastest grid index: the CTAs that share one 128-row A tile are co-scheduled, so A is fetched from HBM
    // once and re-read from L2; B (the filters) is small and always L2-resident.
    const int m0 = blockIdx.y*GEMM_BM, n0 = blockIdx.x*p.bn;
    const int kb0 = blockIdx.z*p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
    const uint32_t tmem_cols = p.bn <= 32 ? 32u : p.bn <= 64 ? 64u : p.bn <= 128 ? 128u : 256u;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(tmem_full, 1);
        tc::fence_barrier_init();
    } else if (warp == 1) {
        tc::tmem_alloc(tmem_slot, tmem_cols);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const ConvOperand& cv = p.cv;
            // im2col state of the single producer thread (all integer divisions happen here, once per CTA / k-block)
            int bn_ = 0, bw = 0, bh = 0;              // mode 1: image index and TMA base coordinate of the tile's first row
            int tap = 0, cb = 0;                      // mode 1: current filter tap and channel block
            int slab_c[8], slab_rs[8];                // mode 2: channel origin and (r << 8 | s) of every 32-column slab
            if (cv.mode == 1) {
                const int hw = cv.gridH*cv.gridW;
                bn_ = m0/hw;
                const int rem = m0 - bn_*hw, py = rem/cv.gridW, px = rem - py*cv.gridW;
                bw = cv.lw + px*cv.stride; bh = cv.lh + py*cv.stride;
                tap = kb0/cv.cblocks; cb = kb0 - tap*cv.cblocks;
            } else if (cv.mode == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = n0 + 32*j, t = col/cv.C;
                    const bool ok = j < p.bn/32 && t < cv.R*cv.S;
                    slab_c[j] = ok ? col - t*cv.C : cv.C;  // channel coordinate C = fully out of bounds = zeros
                    const int r = ok ? t/cv.S : 0, sx = ok ? t - r*cv.S : 0;
                    slab_rs[j] = (r << 8) | sx;
                }
            }
            int it = 0;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it/p.stages) & 1u;
                tc::mbar_wait_spin(&empty[s], ph ^ 1u);
                tc::mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
                uint8_t* a = smem + (size_t)s*stage_bytes;
                uint8_t* b = a + GEMM_A_BYTES;
                const int k = kb*GEMM_BK;
                if (cv.mode == 1) {
                    const int r = tap/cv.S, sx = tap - r*cv.S;
                    tc::tma_load_im2col_4d(a, &tmA, &full[s], cb*GEMM_BK, bw, bh, bn_, (uint16_t)(cv.flip ? cv.S - 1 - sx : sx),
                                           (uint16_t)(cv.flip ? cv.R - 1 - r : r));
                    if (!p.b_mn) tc::tma_load_2d(b, &tmB, &full[s], k, n0);
                    else
                        for (int j = 0; j < p.bn/32; ++j)
                            tc::tma_load_2d(b + j*SLAB_MN_BYTES, &tmB, &full[s], ((cv.r0 + cv.tstep*r)*cv.Sfull + cv.s0 + cv.tstep*sx)*cv.b_tap_cols + n0 + 32*j,
                                            cb*GEMM_BK);
                    if (++cb == cv.cblocks) { cb = 0; ++tap; }
                    continue;
                }
                if (!p.a_mn) tc::tma_load_2d(a, &tmA, &full[s], k, m0);
                else
                    for (int j = 0; j < GEMM_BM/32; ++j) tc::tma_load_2d(a + j*SLAB_MN_BYTES, &tmA, &full[s], m0 + 32*j, k);
                if (cv.mode == 2) {
                    const int hw = cv.gridH*cv.gridW, n = k/hw, rem = k - n*hw, py = rem/cv.gridW, px = rem - py*cv.gridW;
                    const int w = cv.lw + px*cv.stride, h = cv.lh + py*cv.stride;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < p.bn/32)
                            tc::tma_load_im2col_4d(b + j*SLAB_MN_BYTES, &tmB, &full[s], slab_c[j], w, h, n, (uint16_t)(slab_rs[j] & 255),
                                                   (uint16_t)(slab_rs[j] >> 8));
                } else if (!p.b_mn) tc::tma_load_2d(b, &tmB, &full[s], k, n0);
                else
                    for (int j = 0; j < p.bn/32; ++j) tc::tma_load_2d(b + j*SLAB_MN_BYTES, &tmB, &full[s], n0 + 32*j, k);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_tf32(GEMM_BM, p.bn, p.a_mn != 0, p.b_mn != 0);
            int it = 0;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % p.stages;
                const uint32_t ph = (uint32_t)(it/p.stages) & 1u;
                tc::mbar_wait_spin(&full[s], ph);
                tc::tcgen05_fence_after();
                const uint32_t a = tc::smem_u32(smem + (size_t)s*stage_bytes), b = a + GEMM_A_BYTES;
#pragma unroll
                for (int k8 = 0; k8 < GEMM_BK/8; ++k8) {
                    const uint64_t da = p.a_mn ? tc::umma_desc_mnmajor(a, k8, SLAB_MN_BYTES) : tc::umma_desc_kmajor(a, k8);
                    const uint64_t db = p.b_mn ? tc::umma_desc_mnmajor(b, k8, SLAB_MN_BYTES) : tc::umma_desc_kmajor(b, k8);
                    tc::umma_tf32(tmem_base, da, db, idesc, (it > 0 || k8 > 0) ? 1u : 0u);
                }
                tc::umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
            }
            tc::umma_commit(tmem_full);      // accumulator complete
        }
        __syncwarp();
    } else {
        tc::mbar_wait(tmem_full, 0);
        tc::tcgen05_fence_after();
        // Every MMA has retired (tmem_full), so the operand ring is free: its first bytes stage the transposed output chunks.
        const RowMap rm = {p.ldc, p.remap, p.cv.gridH, p.cv.gridW, p.oH, p.oW, p.ost, p.oa, p.ob};
        // Warps 2..9: lane quarter = warp & 3 (the TMEM lanes a warp may read), column chunks interleaved between the two warps
        // of a quarter — the epilogue math (GELU, GELU') is what bounds the narrow-K layers, so it gets 8 of the 10 warps.
        const int ew = warp - 2;
        epilogue_tile(tmem_base, warp & 3, lane, m0, n0, p.bn, p.M, p.N, p.C, rm, p.e, (float*)smem + ew*EPI_WARP_FLOATS, (ew >> 2)*32, 64);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// ---- persistent variant ------------------------------------------------------------------------------------------------------
// Same roles, but a CTA (two resident per SM) walks a strided list of output tiles: the barriers, the TMEM allocation and the
// tensor-map prefetch are paid once; the producer keeps streaming k-blocks of the NEXT tile into the ring while the epilogue
// warps drain the current one; and the accumulator is double-buffered in TMEM (2 x cols columns), so the MMAs of tile j+1
// overlap the epilogue of tile j. The epilogue stages its transposed chunks in a dedicated region (the ring stays busy).
//   tmem_full[a]  : MMA warp -> epilogue ("accumulator a holds a finished tile"), one tcgen05.commit per tile
//   tmem_empty[a] : epilogue -> MMA warp ("accumulator a has been read out"), one arrival per epilogue warp
__global__ void __launch_bounds__(GEMM_THREADS_WIDE, 1)
gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_bytes = p.bn*GEMM_BK*4;
    const int stage_bytes = GEMM_A_BYTES + b_bytes;
    const int nepi = (int)(blockDim.x >> 5) - 2;     // epilogue warps: 8 (two CTAs per SM) or 16 (one)
    float* staging = (float*)(smem + (size_t)p.stages*stage_bytes);
    uint64_t* full = (uint64_t*)(staging + nepi*EPI_WARP_FLOATS);
    uint64_t* empty = full + GEMM_MAX_STAGES;
    uint64_t* tmem_full = empty + GEMM_MAX_STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

    const int nt = (p.N + p.bn - 1)/p.bn, mt = (p.M + GEMM_BM - 1)/GEMM_BM;
    const int total = nt*mt*p.splits;
    const uint32_t acc_cols = p.bn <= 32 ? 32u : p.bn <= 64 ? 64u : 128u;
    if (threadIdx.x == 0) STV_TRACE(0);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
        for (int s = 0; s < p.stages; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&tmem_full[a], 1);
            tc::mbar_init(&tmem_empty[a], (uint32_t)nepi);
        }
        tc::fence_barrier_init();
    } else if (warp == 1) {
        tc::tmem_alloc(tmem_slot, 2*acc_cols);
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) STV_TRACE(1);
    [[maybe_unused]] int trace_i = 0;

    // Producer and MMA issuer: the WHOLE warp runs the loop control (warp-uniform values: ring slot / phase counters instead of
    // `it % stages`, incremental tap / pixel counters instead of per-k-block divisions, descriptors = template + address) and
    // lane 0 issues — see the note in stv_tc.cuh: with everything inside `if (lane == 0)` the issuing thread needed ~150
    // instructions per k-block and was the bottleneck of the kernel.
    const uint32_t ring = tc::smem_u32(smem), full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
    const int stages = p.stages;
    if (warp == 0) {
        const ConvOperand& cv = p.cv;
        const int mode = cv.mode, nslab = p.bn >> 5, pf = p.pf;
        int s = 0;
        uint32_t ph = 0;   // ring slot and its phase carry over from tile to tile
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            const int n0 = (t % nt)*p.bn, m0 = ((t/nt) % mt)*GEMM_BM, z = t/(nt*mt);
            const int kb0 = z*p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            int bn_ = 0, bw = 0, bh = 0, r = 0, sx = 0, cb = 0;   // mode 1: tile origin, current tap and channel block
            int pr = 0, psx = 0, pcb = 0;                         // mode 1: the same counters `pf` k-blocks ahead (L2 prefetch)
            int pn = 0, py = 0, px = 0;                           // mode 2: grid pixel of the k-block's first row
            int slab_c[8], slab_rs[8];
            if (mode == 1) {
                const int hw = cv.gridH*cv.gridW;
                bn_ = m0/hw;
                const int rem = m0 - bn_*hw, qy = rem/cv.gridW, qx = rem - qy*cv.gridW;
                bw = cv.lw + qx*cv.stride; bh = cv.lh + qy*cv.stride;
                const int tap = kb0/cv.cblocks;
                cb = kb0 - tap*cv.cblocks; r = tap/cv.S; sx = tap - r*cv.S;
                const int ptap = (kb0 + pf)/cv.cblocks;
                pcb = kb0 + pf - ptap*cv.cblocks; pr = ptap/cv.S; psx = ptap - pr*cv.S;
            } else if (mode == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int col = n0 + 32*j, tp = col/cv.C;
                    const bool ok = j < nslab && tp < cv.R*cv.S;
                    slab_c[j] = ok ? col - tp*cv.C : cv.C;
                    const int rr = ok ? tp/cv.S : 0, ss = ok ? tp - rr*cv.S : 0;
                    slab_rs[j] = (rr << 8) | ss;
                }
                const int hw = cv.gridH*cv.gridW, k = kb0*GEMM_BK;
                pn = k/hw;
                const int rem = k - pn*hw;
                py = rem/cv.gridW; px = rem - py*cv.gridW;
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                tc::mbar_wait_spin_s(empty0 + 8*s, ph ^ 1u);
                if (lane == 0) { STV_TRACE_KB(16, trace_i); ++trace_i; }
                if (tc::elect_one()) {
                    const uint32_t fb = full0 + 8*s, a = ring + (uint32_t)(s*stage_bytes), b = a + GEMM_A_BYTES;
                    tc::mbar_arrive_expect_tx_s(fb, (uint32_t)stage_bytes);
                    const int k = kb*GEMM_BK;
                    if (mode == 1) {
                        tc::tma_load_im2col_4d_s(a, &tmA, fb, cb*GEMM_BK, bw, bh, bn_, (uint16_t)(cv.flip ? cv.S - 1 - sx : sx),
                                                 (uint16_t)(cv.flip ? cv.R - 1 - r : r));
                        if (!p.b_mn) tc::tma_load_2d_s(b, &tmB, fb, k, n0);
                        else {
                            const int col = ((cv.r0 + cv.tstep*r)*cv.Sfull + cv.s0 + cv.tstep*sx)*cv.b_tap_cols + n0;
                            for (int j = 0; j < nslab; ++j) tc::tma_load_2d_s(b + j*SLAB_MN_BYTES, &tmB, fb, col + 32*j, cb*GEMM_BK);
                        }
                    } else {
                        if (!p.a_mn) tc::tma_load_2d_s(a, &tmA, fb, k, m0);
                        else {
#pragma unroll
                            for (int j = 0; j < GEMM_BM/32; ++j) tc::tma_load_2d_s(a + j*SLAB_MN_BYTES, &tmA, fb, m0 + 32*j, k);
                        }
                        if (mode == 2) {
                            const int w = cv.lw + px*cv.stride, h = cv.lh + py*cv.stride;
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (j < nslab)
                                    tc::tma_load_im2col_4d_s(b + j*SLAB_MN_BYTES, &tmB, fb, slab_c[j], w, h, pn, (uint16_t)(slab_rs[j] & 255),
                                                             (uint16_t)(slab_rs[j] >> 8));
                        } else if (!p.b_mn) tc::tma_load_2d_s(b, &tmB, fb, k, n0);
                        else
                            for (int j = 0; j < nslab; ++j) tc::tma_load_2d_s(b + j*SLAB_MN_BYTES, &tmB, fb, n0 + 32*j, k);
                    }
                    // L2 prefetch of the STREAMED operand(s) `pf` k-blocks ahead (activations / gradients come from HBM; filters and
                    // weights are L2-resident and are not prefetched)
                    if (pf > 0 && kb + pf < kb1) {
                        const int kp = (kb + pf)*GEMM_BK;
                        if (mode == 1) tc::tma_prefetch_im2col_4d(&tmA, pcb*GEMM_BK, bw, bh, bn_, (uint16_t)(cv.flip ? cv.S - 1 - psx : psx),
                                                                  (uint16_t)(cv.flip ? cv.R - 1 - pr : pr));
                        else if (!p.a_mn) tc::tma_prefetch_2d(&tmA, kp, m0);
                        else {
#pragma unroll
                            for (int j = 0; j < GEMM_BM/32; ++j) tc::tma_prefetch_2d(&tmA, m0 + 32*j, kp);
                            if (mode == 0 && p.b_mn)
                                for (int j = 0; j < nslab; ++j) tc::tma_prefetch_2d(&tmB, n0 + 32*j, kp);
                        }
                    }
                }
                if (mode == 1) {
                    if (++cb == cv.cblocks) { cb = 0; if (++sx == cv.S) { sx = 0; ++r; } }
                    if (++pcb == cv.cblocks) { pcb = 0; if (++psx == cv.S) { psx = 0; ++pr; } }
                }
                else if (mode == 2) {
                    px += GEMM_BK;
                    while (px >= cv.gridW) { px -= cv.gridW; if (++py == cv.gridH) { py = 0; ++pn; } }
                }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = tc::umma_idesc_tf32(GEMM_BM, p.bn, p.a_mn != 0, p.b_mn != 0);
        const uint64_t da0 = tc::umma_desc_template(p.a_mn != 0, SLAB_MN_BYTES), db0 = tc::umma_desc_template(p.b_mn != 0, SLAB_MN_BYTES);
        const uint32_t ka = tc::umma_desc_kstep(p.a_mn != 0), kbs = tc::umma_desc_kstep(p.b_mn != 0);
        const uint32_t tmem_full0 = tc::smem_u32(tmem_full), tmem_empty0 = tc::smem_u32(tmem_empty);
        int s = 0, j = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
            const int z = t/(nt*mt);
            const int kb0 = z*p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            const int acc = j & 1;
            tc::mbar_wait_spin_s(tmem_empty0 + 8*acc, ((uint32_t)(j >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
            tc::tcgen05_fence_after();
            const uint32_t d = tmem_base + (uint32_t)acc*acc_cols;
            for (int kb = kb0; kb < kb1; ++kb) {
                tc::mbar_wait_spin_s(full0 + 8*s, ph);
                if (lane == 0) { STV_TRACE_KB(528, trace_i); ++trace_i; }
                tc::tcgen05_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a = (ring + (uint32_t)(s*stage_bytes)) >> 4;
                    uint64_t da = da0 + a, db = db0 + (a + (GEMM_A_BYTES >> 4));
                    tc::umma_tf32(d, da, db, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
                    for (int k8 = 1; k8 < GEMM_BK/8; ++k8) {
                        da += ka; db += kbs;
                        tc::umma_tf32(d, da, db, idesc, 1u);
                    }
                    tc::umma_commit_s(empty0 + 8*s);   // frees the stage once these MMAs have read it
                }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            if (tc::elect_one()) tc::umma_commit_s(tmem_full0 + 8*acc);
            if (lane == 0 && j == 0) STV_TRACE(2);
        }
        __syncwarp();
    } else {
        const RowMap rm = {p.ldc, p.remap, p.cv.gridH, p.cv.gridW, p.oH, p.oW, p.ost, p.oa, p.ob};
        const int ew = warp - 2;
        int j = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
            const int n0 = (t % nt)*p.bn, m0 = ((t/nt) % mt)*GEMM_BM;
            const int acc = j & 1;
            tc::mbar_wait(&tmem_full[acc], (uint32_t)(j >> 1) & 1u);
            tc::tcgen05_fence_after();
            if (warp == 2 && lane == 0 && j == 0) STV_TRACE(3);
            epilogue_tile(tmem_base + (uint32_t)acc*acc_cols, warp & 3, lane, m0, n0, p.bn, p.M, p.N, p.C, rm, p.e,
                          staging + ew*EPI_WARP_FLOATS, (ew >> 2)*32, 8*nepi);
            if (warp == 2 && lane == 0 && j == 0) STV_TRACE(4);
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
        }
    }
    tc::tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) STV_TRACE(5);
    if (warp == 1) tc::tmem_dealloc(tmem_base, 2*acc_cols);
}

// ---- CTA-pair variant (cta_group::2) ------------------------------------------------------------------------------------------
// Plain matrices only (no im2col operand). A cluster of two CTAs on neighbouring SMs owns a 256 x BN output tile (BN <= 256): each
// CTA loads ITS 128 rows of A and ITS half of the B tile (BN/2 rows of a K-major B, or half of the 32-column slabs of an MN-major
// B), the leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256) against both CTAs' shared memory, and each CTA's
// epilogue warps drain their own 128 TMEM lanes. Per flop a 256 x 256 pair tile moves HALF the operand bytes of 128 x 128 tiles
// through L2 -> shared memory and reads half as much shared memory per MMA — TF32 operands are 4 bytes, so the 128 x 128 kernel
// needs 118 B/clk of operand reads per SM at full tensor rate. One CTA per SM (the accumulators are double-buffered
// over all 512 TMEM columns at BN = 256), 5-8 stage ring.
//   full[s]        leader only: bytes of BOTH CTAs' TMA loads (2-SM loads signal the leader's barrier)
//   empty[s]       both CTAs: multicast commit of the leader once the MMAs have read stage s
//   tmem_full[a]   both CTAs: multicast commit once accumulator a holds a finished tile
//   tmem_empty[a]  leader only: 16 arrivals = the 8 epilogue warps of each CTA (remote arrive from the peer)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS_WIDE, 1)
gemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = tc::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int half_bn = p.bn/2;
    const int b_bytes = half_bn*GEMM_BK*4;
    const int stage_bytes = GEMM_A_BYTES + b_bytes;
    const int stages = p.stages;
    const int nepi = (int)(blockDim.x >> 5) - 2;
    float* staging = (float*)(smem + (size_t)stages*stage_bytes);
    uint64_t* full = (uint64_t*)(staging + nepi*EPI_WARP_FLOATS);
    uint64_t* empty = full + GEMM_MAX_STAGES;
    uint64_t* tmem_full = empty + GEMM_MAX_STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

    const int nt = (p.N + p.bn - 1)/p.bn, mt = (p.M + 2*GEMM_BM - 1)/(2*GEMM_BM);
    const int total = nt*mt*p.splits;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const uint32_t acc_cols = p.bn <= 64 ? 64u : p.bn <= 128 ? 128u : 256u;
    if (threadIdx.x == 0) STV_TRACE(0);
    [[maybe_unused]] int trace_i = 0;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
        for (int s = 0; s < stages; ++s) {
            tc::mbar_init(&full[s], 1);
            tc::mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(&tmem_full[a], 1);
            tc::mbar_init(&tmem_empty[a], (uint32_t)(2*nepi));
        }
        tc::fence_barrier_init();
    } else if (warp == 1) {
        tc::tmem_alloc_2sm(tmem_slot, 2*acc_cols);
    }
    tc::tcgen05_fence_before();
    tc::cluster_sync();   // the leader's barriers exist before the peer's TMA signals them
    tc::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t ring = tc::smem_u32(smem), full0 = tc::smem_u32(full), empty0 = tc::smem_u32(empty);
    if (threadIdx.x == 0) STV_TRACE(1);

    if (warp == 0) {
        const ConvOperand& cv = p.cv;
        const int mode = cv.mode, nslab = half_bn >> 5;
        int s = 0;
        uint32_t ph = 0;
        for (int t = pair; t < total; t += npairs) {
            const int n0 = (t % nt)*p.bn + (int)rank*half_bn, m0 = ((t/nt) % mt)*2*GEMM_BM + (int)rank*GEMM_BM, z = t/(nt*mt);
            const int kb0 = z*p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            int bn_ = 0, bw = 0, bh = 0, r = 0, sx = 0, cb = 0;   // mode 1 (A = im2col rows): this CTA's tile origin, current tap and channel block
            if (mode == 1) {
                const int hw = cv.gridH*cv.gridW;
                bn_ = m0/hw;
                const int rem = m0 - bn_*hw, qy = rem/cv.gridW, qx = rem - qy*cv.gridW;
                bw = cv.lw + qx*cv.stride; bh = cv.lh + qy*cv.stride;
                const int tap = kb0/cv.cblocks;
                cb = kb0 - tap*cv.cblocks; r = tap/cv.S; sx = tap - r*cv.S;
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                tc::mbar_wait_spin_s(empty0 + 8*s, ph ^ 1u);
                if (lane == 0) { STV_TRACE_KB(16, trace_i); ++trace_i; }
                if (tc::elect_one()) {
                    const uint32_t fb = full0 + 8*s, a = ring + (uint32_t)(s*stage_bytes), b = a + GEMM_A_BYTES;
                    if (rank == 0) tc::mbar_arrive_expect_tx_s(fb, 2u*(uint32_t)stage_bytes);
                    const int k = kb*GEMM_BK;
                    if (mode == 1) {
                        tc::tma_load_im2col_4d_2sm_s(a, &tmA, fb, cb*GEMM_BK, bw, bh, bn_, (uint16_t)(cv.flip ? cv.S - 1 - sx : sx),
                                                     (uint16_t)(cv.flip ? cv.R - 1 - r : r));
                        if (!p.b_mn) tc::tma_load_2d_2sm_s(b, &tmB, fb, k, n0);
                        else {
                            const int col = ((cv.r0 + cv.tstep*r)*cv.Sfull + cv.s0 + cv.tstep*sx)*cv.b_tap_cols + n0;
                            for (int j = 0; j < nslab; ++j) tc::tma_load_2d_2sm_s(b + j*SLAB_MN_BYTES, &tmB, fb, col + 32*j, cb*GEMM_BK);
                        }
                    } else {
                        if (!p.a_mn) tc::tma_load_2d_2sm_s(a, &tmA, fb, k, m0);
                        else {
#pragma unroll
                            for (int j = 0; j < GEMM_BM/32; ++j) tc::tma_load_2d_2sm_s(a + j*SLAB_MN_BYTES, &tmA, fb, m0 + 32*j, k);
                        }
                        if (!p.b_mn) tc::tma_load_2d_2sm_s(b, &tmB, fb, k, n0);
                        else
                            for (int j = 0; j < nslab; ++j) tc::tma_load_2d_2sm_s(b + j*SLAB_MN_BYTES, &tmB, fb, n0 + 32*j, k);
                    }
                }
                if (mode == 1) { if (++cb == cv.cblocks) { cb = 0; if (++sx == cv.S) { sx = 0; ++r; } } }
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const uint32_t idesc = tc::umma_idesc_tf32(2*GEMM_BM, p.bn, p.a_mn != 0, p.b_mn != 0);
            const uint64_t da0 = tc::umma_desc_template(p.a_mn != 0, SLAB_MN_BYTES), db0 = tc::umma_desc_template(p.b_mn != 0, SLAB_MN_BYTES);
            const uint32_t ka = tc::umma_desc_kstep(p.a_mn != 0), kbs = tc::umma_desc_kstep(p.b_mn != 0);
            const uint32_t tmem_full0 = tc::smem_u32(tmem_full), tmem_empty0 = tc::smem_u32(tmem_empty);
            int s = 0, j = 0;
            uint32_t ph = 0;
            for (int t = pair; t < total; t += npairs, ++j) {
                const int z = t/(nt*mt);
                const int kb0 = z*p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
                const int acc = j & 1;
                tc::mbar_wait_spin_s(tmem_empty0 + 8*acc, ((uint32_t)(j >> 1) & 1u) ^ 1u);  // both CTAs have drained this accumulator
                tc::tcgen05_fence_after();
                const uint32_t d = tmem_base + (uint32_t)acc*acc_cols;
                for (int kb = kb0; kb < kb1; ++kb) {
                    tc::mbar_wait_spin_s(full0 + 8*s, ph);
                    if (lane == 0) { STV_TRACE_KB(528, trace_i); ++trace_i; }
                    tc::tcgen05_fence_after();
                    if (tc::elect_one()) {
                        const uint32_t a = (ring + (uint32_t)(s*stage_bytes)) >> 4;
                        uint64_t da = da0 + a, db = db0 + (a + (GEMM_A_BYTES >> 4));
                        tc::umma_tf32_2sm(d, da, db, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
                        for (int k8 = 1; k8 < GEMM_BK/8; ++k8) {
                            da += ka; db += kbs;
                            tc::umma_tf32_2sm(d, da, db, idesc, 1u);
                        }
                        tc::umma_commit_2sm_s(empty0 + 8*s);
                    }
                    if (++s == stages) { s = 0; ph ^= 1u; }
                }
                if (tc::elect_one()) tc::umma_commit_2sm_s(tmem_full0 + 8*acc);
                if (lane == 0 && j == 0) STV_TRACE(2);
            }
        }
        __syncwarp();
    } else {
        const RowMap rm = {p.ldc, p.remap, p.cv.gridH, p.cv.gridW, p.oH, p.oW, p.ost, p.oa, p.ob};
        const int ew = warp - 2;
        int j = 0;
        for (int t = pair; t < total; t += npairs, ++j) {
            const int n0 = (t % nt)*p.bn, m0 = ((t/nt) % mt)*2*GEMM_BM + (int)rank*GEMM_BM;
            const int acc = j & 1;
            tc::mbar_wait(&tmem_full[acc], (uint32_t)(j >> 1) & 1u);
            tc::tcgen05_fence_after();
            if (warp == 2 && lane == 0 && j == 0) STV_TRACE(3);
            epilogue_tile(tmem_base + (uint32_t)acc*acc_cols, warp & 3, lane, m0, n0, p.bn, p.M, p.N, p.C, rm, p.e,
                          staging + ew*EPI_WARP_FLOATS, (ew >> 2)*32, 8*nepi);
            if (warp == 2 && lane == 0 && j == 0) STV_TRACE(4);
            tc::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_remote(&tmem_empty[acc], 0);   // the leader's barrier (also from the leader itself)
        }
    }
    tc::tcgen05_fence_before();
    tc::cluster_sync();   // both CTAs are done with TMEM and with each other's shared memory
    if (threadIdx.x == 0) STV_TRACE(5);
    if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, 2*acc_cols);
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    });
    return fn;
}

// fp32 row-major matrix [rows][ld] with `cols` valid columns; box = {32 floats, box_rows}, zero OOB fill.
// mn_major = 0: 128-byte swizzle (K-major slabs); 1: 128-byte swizzle with 32-byte atoms (MN-major tf32 slabs).
int make_tmap_2d(CUtensorMap* tm, const float* base, long long rows, long long cols, long long ld, int box_rows, int mn_major) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the CUDA driver"); return STV_E_CUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld*4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a %lld x %lld matrix, ld %lld, box rows %d, base %p", (int)r, rows, cols, ld,
                  box_rows, (const void*)base);
        return STV_E_CUDA;
    }
    return STV_OK;
}

int make_tmap_3d(CUtensorMap* tm, const float* base, long long W, long long H, long long planes, int bw, int bh, int bp) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the CUDA driver"); return STV_E_CUDA; }
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W*4, (cuuint64_t)W*H*4};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bp};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for a (%lld,%lld,%lld) tensor, box (%d,%d,%d), base %p", (int)r, planes, H, W, bp, bh, bw,
                  (const void*)base);
        return STV_E_CUDA;
    }
    return STV_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*,
                                   const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_im2col(CUtensorMap* tm, const float* base, int N, int H, int W, int C, int lw, int lh, int uw, int uh, int stride,
                     int pixels, int mn_major) {
    static EncodeIm2colFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeIm2colFn)ptr;
    });
    if (!fn) { set_error("cuTensorMapEncodeIm2col is not available from the CUDA driver"); return STV_E_CUDA; }
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C*4, (cuuint64_t)W*C*4, (cuuint64_t)H*W*C*4};
    const int lower[2] = {lw, lh}, upper[2] = {uw, uh};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, lower, upper, 32u, (cuuint32_t)pixels, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeIm2col failed (%d): tensor (%d,%d,%d,%d), corners (%d,%d)/(%d,%d), stride %d, %d pixels", (int)r, N, H, W, C,
                  lw, lh, uw, uh, stride, pixels);
        return STV_E_CUDA;
    }
    return STV_OK;
}

// Tile width: the multiple of 32 (<= 128) that wastes the fewest columns (ties -> wider); when the resulting grid would leave
// SMs idle (few row tiles: the deep, low-resolution layers) it is narrowed, down to 64, until the grid covers the 148 SMs.
int pick_bn(int N, long long row_tiles) {
    // <= 128 columns: a 128 x 128 tile keeps the operand ring at 32 KB per stage, so two CTAs (20 warps) stay resident per SM and
    // one CTA's epilogue overlaps the other's main loop; wider tiles halve the residency and left the epilogue-heavy layers
    // (GELU / GELU' over 4C columns with only 3-12 k-blocks) latency-bound.
    // Fewest column tiles first, then the narrowest width that still gives that count: a kind::tf32 MMA costs the same ~80 clocks
    // for every N <= 128 (profiles/r2_conv3_timeline.txt), so the MMA count — not the padded columns — is what a tile width buys.
    // (N = 160: two 96-wide tiles instead of five 32-wide ones; N = 576: five 128-wide instead of six 96-wide.)
    // STV_GEMM_BN_RULE=0 restores the round-1 rule (fewest padded columns) for A/B runs.
    static const int rule = getenv("STV_GEMM_BN_RULE") ? atoi(getenv("STV_GEMM_BN_RULE")) : 1;
    int best = 32, best_cost = 1 << 30;
    if (rule == 0) {
        for (int bn = 128; bn >= 32; bn -= 32) {
            const int tiles = (N + bn - 1)/bn, cost = tiles*bn;
            if (cost < best_cost) { best = bn; best_cost = cost; }
        }
    } else {
        const int tiles = (N + 127)/128;
        best = ((N + tiles - 1)/tiles + 31)/32*32;
    }
    // Measured and rejected (profiles/r2_rejected_variants.txt): narrowing only while every tile still gets an SM of its own
    // (STV_GEMM_NARROW=0). The deep, narrow layers this changes (stage-3 ConvNeXt products: 90 full-width tiles instead of 180
    // half-width ones) are bound by the L2 -> SM operand stream (276 MB in 50 us = 5.5 TB/s), not by the MMA rate: no gain.
    static const int narrow_always = getenv("STV_GEMM_NARROW") ? atoi(getenv("STV_GEMM_NARROW")) : 1;
    while (best > 64 && best % 64 == 0 && row_tiles*((N + best - 1)/best) < 148 &&
           (narrow_always || row_tiles*((N + best/2 - 1)/(best/2)) <= 148)) best /= 2;
    return best;
}

static bool persistent_enabled() {
    static int on = -1;
    if (on < 0) { const char* v = getenv("STV_GEMM_PERSISTENT"); on = (v && v[0] == '0') ? 0 : 1; }  // developer switch for A/B runs
    return on != 0;
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmParams& p, int splits, cudaStream_t stream, const char* what) {
    const int stage_bytes = GEMM_A_BYTES + p.bn*GEMM_BK*4;
    const int nt = (p.N + p.bn - 1)/p.bn, mt = (p.M + GEMM_BM - 1)/GEMM_BM;
    p.splits = splits;
    if (mt > 65535) { set_error("%s: too many row tiles (%d)", what, mt); return STV_E_ARG; }
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024);
        if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(gemm_tf32_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024);
    });
    if (attr_err != cudaSuccess) { set_error("%s: cudaFuncSetAttribute failed (%s)", what, cudaGetErrorString(attr_err)); return STV_E_CUDA; }
    const long long total = (long long)nt*mt*splits;
    // CTA pairs (cta_group::2): plain matrices and im2col-A convolutions (forward / data gradient) whose 256 x bn2 pair tiles keep most of the 74 SM pairs busy. bn2 = the multiple of 64
    // (<= 256) that wastes the fewest columns (ties -> wider): each CTA's half of the B tile is whole 32-column slabs.
    static const int pair_mode = getenv("STV_GEMM_PAIR") ? atoi(getenv("STV_GEMM_PAIR")) : 1;   // developer switch: 0 off, 1 heuristic, 2 whenever legal
    static const int pair_conv = getenv("STV_GEMM_PAIR_CONV") ? atoi(getenv("STV_GEMM_PAIR_CONV")) : 0;   // im2col-A pairs: measured slower on the step
    if (pair_mode && (p.cv.mode == 0 || (p.cv.mode == 1 && (pair_conv || pair_mode == 2))) && p.pair_B != nullptr && p.M > GEMM_BM && p.N >= 64) {
        int bn2 = 128, best_cost = 1 << 30;
        for (int bn = 256; bn >= 64; bn -= 64) {
            const int cost = ((p.N + bn - 1)/bn)*bn;
            if (cost < best_cost) { bn2 = bn; best_cost = cost; }
        }
        const int mt2 = (p.M + 2*GEMM_BM - 1)/(2*GEMM_BM);
        const int npairs_max = sm_count()/2;
        const int nt2 = (p.N + bn2 - 1)/bn2;
        const long long tiles2 = (long long)nt2*mt2*splits;
        const bool waste_ok = (long long)nt2*bn2*4 <= (long long)p.N*5;   // <= 25 % padded columns
        // Measured (profiles/r2_gemm_pair_vs_single.txt): the pair kernel wins 20-30 % on the plain / bias / residual / split-K
        // products once ~80 % of the 74 SM pairs have a tile; it loses on the GELU / GELU' epilogues, which are issue-bound and
        // want the 16 epilogue warps per SM of the two-CTA kernel.
        const bool heavy_epi = p.e.act == STV_ACT_GELU || (p.e.dact_src != nullptr && p.e.dact == STV_ACT_GELU);
        if (pair_mode == 2 || (tiles2 >= (npairs_max*4)/5 && waste_ok && !heavy_epi)) {
            static std::once_flag once2;
            static cudaError_t err2 = cudaSuccess;
            std::call_once(once2, [] { err2 = cudaFuncSetAttribute(gemm_tf32_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024); });
            if (err2 != cudaSuccess) { set_error("%s: cudaFuncSetAttribute failed (%s)", what, cudaGetErrorString(err2)); return STV_E_CUDA; }
            CUtensorMap tmB2 = tmB;
            if (!p.b_mn) { if (int rc = make_tmap_2d(&tmB2, p.pair_B, p.N, p.K, p.pair_ldb, bn2/2, 0)) return rc; }
            GemmParams q = p;
            q.bn = bn2;
            const int sb = GEMM_A_BYTES + (bn2/2)*GEMM_BK*4;
            static const int pair_threads = (getenv("STV_GEMM_PAIR_EPI") && atoi(getenv("STV_GEMM_PAIR_EPI")) == 8) ? GEMM_THREADS : GEMM_THREADS_WIDE;
            const int staging = (pair_threads/32 - 2)*EPI_WARP_FLOATS*4;
            int stages = (226*1024 - staging - 2048)/sb;
            stages = stages > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : stages;
            q.stages = stages;
            const size_t smem = (size_t)stages*sb + staging + 1024 + (2*GEMM_MAX_STAGES + 4)*8 + 16;
            const int pairs = (int)(tiles2 < npairs_max ? tiles2 : npairs_max);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2*pairs); cfg.blockDim = dim3(pair_threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tf32_pair_kernel, tmA, tmB2, q);
            count_launch();
            if (le != cudaSuccess) { set_error("%s: cluster launch failed (%s)", what, cudaGetErrorString(le)); return STV_E_CUDA; }
            return check_launch(what);
        }
    }
    if (persistent_enabled() && total < (1ll << 30)) {
        // Two resident CTAs per SM (ring + epilogue staging + barriers within ~112 KB each: one CTA's epilogue overlaps the other's
        // main loop), or ONE with the whole shared memory as a deeper ring (developer switch STV_GEMM_RESIDENT=1).
        // STV_GEMM_RESIDENT: 1 = always one CTA per SM (deep ring), 2 = one CTA per SM for launches of at most one tile per SM,
        // 3 = always two (default). Measured (profiles/r2_rejected_variants.txt): mode 2 changes nothing on the single-wave stage-3
        // products (0.050 vs 0.047 ms for dx = dz.W1) — they are L2-stream-bound, the deeper ring has nothing to hide.
        static const int res_cfg = getenv("STV_GEMM_RESIDENT") ? atoi(getenv("STV_GEMM_RESIDENT")) : 3;
        const int per_sm = (res_cfg == 1 || (res_cfg == 2 && total <= sm_count() && p.kb_per_split >= 16)) ? 1 : 2;
        const int threads = per_sm == 1 ? GEMM_THREADS_WIDE : GEMM_THREADS;
        const int staging = (threads/32 - 2)*EPI_WARP_FLOATS*4;
        int stages = ((per_sm == 1 ? 226 : 112)*1024 - staging - 2048)/stage_bytes;
        stages = stages > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : stages;
        stages = stages < 2 ? 2 : stages;
        p.stages = stages;
        // L2 prefetch distance of the streamed operand in k-blocks. Measured (profiles/r2_rejected_variants.txt): 2 / 4 / 8 are all SLOWER than none on the
        // step (431 -> 415 images/s) — the load latency of the 2-stage ring is not DRAM-miss time — so it stays a developer switch.
        static const int pf_cfg = getenv("STV_GEMM_PF") ? atoi(getenv("STV_GEMM_PF")) : 0;
        p.pf = pf_cfg;
        const size_t smem = (size_t)stages*stage_bytes + staging + 1024 /*alignment slack*/ + (2*GEMM_MAX_STAGES + 4)*8 + 16;
        const int resident = per_sm*sm_count();
        const int grid = (int)(total < resident ? total : resident);
        gemm_tf32_persistent_kernel<<<grid, threads, smem, stream>>>(tmA, tmB, p);
        count_launch();
        return check_launch(what);
    }
    int stages = (110*1024)/stage_bytes;                         // two resident CTAs per SM (bn <= 128 -> at least 3 stages)
    stages = stages > GEMM_MAX_STAGES ? GEMM_MAX_STAGES : stages;
    stages = stages > p.kb_per_split ? (p.kb_per_split < 2 ? 2 : p.kb_per_split) : stages;
    p.stages = stages;
    const size_t smem = (size_t)stages*stage_bytes + 1024 /*alignment slack*/ + (2*GEMM_MAX_STAGES + 1)*8 + 16;
    const dim3 grid(nt, mt, splits);
    gemm_tf32_kernel<<<grid, GEMM_THREADS, smem, stream>>>(tmA, tmB, p);
    count_launch();
    return check_launch(what);
}

}  // namespace stv

using namespace stv;

#ifdef STV_GEMM_TRACE
extern "C" int stv_debug_gemm_trace(unsigned long long* out, int n) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, g_gemm_trace, sizeof(unsigned long long)*(size_t)(n < 1040 ? n : 1040));
}
#endif

extern "C" int stv_gemm_tf32(int M, int N, int K, const float* A, long long lda, int a_mn, const float* B, long long ldb, int b_mn,
                             float* C, long long ldc, const stv_gemm_epi* epi, int split_k, void* stream) {
    STV_REQUIRE(M > 0 && N > 0 && K > 0, "stv_gemm_tf32: empty problem (M=%d N=%d K=%d)", M, N, K);
    STV_REQUIRE(A && B && C, "stv_gemm_tf32: null operand");
    STV_REQUIRE(N % 4 == 0 && ldc % 4 == 0 && ((uintptr_t)C & 15) == 0, "stv_gemm_tf32: N and ldc must be multiples of 4, C 16-byte aligned");
    STV_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0,
                "stv_gemm_tf32: lda/ldb must be multiples of 4 and A/B 16-byte aligned (TMA)");
    STV_REQUIRE(lda >= (a_mn ? M : K) && ldb >= (b_mn ? N : K), "stv_gemm_tf32: leading dimension smaller than the row length");
    stv_gemm_epi e = {};
    if (epi) e = *epi;
    STV_REQUIRE(split_k >= 1, "stv_gemm_tf32: split_k must be >= 1");
    if (stv_deterministic()) {   // one contributor per output element; the fused column sums have one per row tile, so they are refused
        split_k = 1;
        STV_REQUIRE(e.colsum == nullptr, "stv_gemm_tf32: the fused column sums are not reproducible (STV_DETERMINISTIC=1): use stv_colsum");
    }
    STV_REQUIRE(split_k == 1 || e.accumulate, "stv_gemm_tf32: split_k > 1 needs an accumulating epilogue");
    STV_REQUIRE(!(e.accumulate && (e.aux || e.act || e.res)), "stv_gemm_tf32: accumulate cannot be combined with aux/act/res");
    if (e.bias) STV_REQUIRE(((uintptr_t)e.bias & 15) == 0, "stv_gemm_tf32: bias must be 16-byte aligned");
    if (e.gamma) STV_REQUIRE(((uintptr_t)e.gamma & 15) == 0, "stv_gemm_tf32: gamma must be 16-byte aligned");

    GemmParams p = {};
    p.M = M; p.N = N; p.K = K;
    p.bn = pick_bn(N, (M + GEMM_BM - 1)/GEMM_BM);
    p.a_mn = a_mn != 0; p.b_mn = b_mn != 0;
    p.kb_total = (K + GEMM_BK - 1)/GEMM_BK;
    split_k = split_k < p.kb_total ? split_k : p.kb_total;
    p.kb_per_split = (p.kb_total + split_k - 1)/split_k;
    split_k = (p.kb_total + p.kb_per_split - 1)/p.kb_per_split;  // every split non-empty
    p.C = C; p.ldc = ldc; p.e = e;
    CUtensorMap tmA, tmB;
    int rc = a_mn ? make_tmap_2d(&tmA, A, K, M, lda, 32, 1) : make_tmap_2d(&tmA, A, M, K, lda, GEMM_BM, 0);
    if (rc) return rc;
    rc = b_mn ? make_tmap_2d(&tmB, B, K, N, ldb, 32, 1) : make_tmap_2d(&tmB, B, N, K, ldb, p.bn, 0);
    if (rc) return rc;
    p.pair_B = B; p.pair_ldb = ldb;   // the CTA-pair kernel re-encodes B with half-tile boxes
    return launch_gemm(tmA, tmB, p, split_k, (cudaStream_t)stream, "stv_gemm_tf32");
}
