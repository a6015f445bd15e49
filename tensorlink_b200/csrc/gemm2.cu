// 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a thread-block cluster of two CTAs on one TPC computes a 256x256
// output tile with ONE tcgen05.mma M=256 instruction stream issued by the leader CTA.
//
//   CTA r of the pair owns accumulator rows [m0 + 128 r, +128) in ITS tensor memory and loads, per k-block,
//   its 128 rows of A (16 KB) and HALF of the B tile (rows n0 + 128 r .. +127, 16 KB): the B operand is fetched
//   once per pair instead of once per CTA, so shared-memory fill traffic per flop halves and 6 stages fit.
//   TMA completions of both CTAs are counted on the leader's full barrier; tcgen05.commit multicasts the
//   "slot free" / "accumulator ready" arrivals to both CTAs; the peer's epilogue warps arrive remotely on the
//   leader's accumulator-empty barrier.
// Warp roles per CTA as in gemm.cu (warp 0 TMA producer, warp 1 TMEM owner [+ MMA issuer in the leader], warps 2..5
// epilogue).  Same operand-major options and the same epilogue (gemm_common.cuh).
#include "gemm_common.cuh"

namespace tl {

constexpr int G2_THREADS = 192;
constexpr int G2_BN = 256;                       // tile N (both halves)
constexpr int G2_A_BYTES = BM * BK * 2;          // 16 KB: this CTA's 128 rows of A
constexpr int G2_B_BYTES = (G2_BN / 2) * BK * 2; // 16 KB: this CTA's half of B
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_STAGES = 6;
constexpr int G2_TMEM_COLS = 2 * G2_BN;
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + 1024 + 256 + EPI_SMEM_BYTES;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D tiled TMA load into LOCAL shared memory whose completion is counted on a (possibly remote) mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint32_t mbar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(mbar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair once all prior MMAs have retired
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, void* __restrict__ Cv, int M,
                  int N, int K, int ldc, const bf16* __restrict__ bias, const bf16* __restrict__ residual, int ldr, int flags) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G2_STAGES * G2_STAGE_BYTES);
    uint64_t* full_bar = bars;                          // [STAGES]  used in the leader only
    uint64_t* empty_bar = bars + G2_STAGES;             // [STAGES]  both CTAs (multicast commit)
    uint64_t* tmem_full = bars + 2 * G2_STAGES;         // [2]       both CTAs (multicast commit)
    uint64_t* tmem_empty = bars + 2 * G2_STAGES + 2;    // [2]       leader only, 8 arrivals (4 epilogue warps x 2 CTAs)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * G2_STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = cluster_ctarank();
    const bool leader = cta_rank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + G2_BN - 1) / G2_BN;
    const int total_tiles = tiles_m * tiles_n;
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < G2_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 8);
        }
        fence_barrier_init();
    }
    cluster_sync_all();                                   // barriers of both CTAs initialised before any remote use
    if (warp == 1) tmem_alloc_2sm<G2_TMEM_COLS>(tmem_base_slot);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        // ===================================================================== TMA producer (each CTA loads its share)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < total_tiles; t += n_clusters) {
                const int m0 = (t % tiles_m) * (2 * BM) + (int)cta_rank * BM;
                const int n0 = (t / tiles_m) * G2_BN + (int)cta_rank * (G2_BN / 2);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * G2_STAGE_BYTES;
                    unsigned char* sb = sa + G2_A_BYTES;
                    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);   // bytes of BOTH CTAs
                    if (!A_MN) {
                        tma_load_2d_2sm(sa, &tmA, full_leader, kb * BK, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(sa + j * 8192, &tmA, full_leader, m0 + 64 * j, kb * BK);
                    }
                    if (!B_MN) {
                        tma_load_2d_2sm(sb, &tmB, full_leader, kb * BK, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < (G2_BN / 2) / 64; ++j) tma_load_2d_2sm(sb + j * 8192, &tmB, full_leader, n0 + 64 * j, kb * BK);
                    }
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer: leader CTA only
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * BM, G2_BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = cluster_id; t < total_tiles; t += n_clusters) {
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * G2_BN);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
                    const uint32_t sb = sa + G2_A_BYTES;
                    const uint64_t da = A_MN ? make_smem_desc_sw128(sa, 8192, 1024) : make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t db = B_MN ? make_smem_desc_sw128(sb, 8192, 1024) : make_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t ak = da + (uint64_t)((A_MN ? 2048 : 32) * k >> 4);
                        const uint64_t bk = db + (uint64_t)((B_MN ? 2048 : 32) * k >> 4);
                        umma_bf16_2sm(d_tmem, ak, bk, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_2sm(&empty_bar[stage]);
                    if (kb == num_k - 1) umma_commit_2sm(&tmem_full[acc]);
                    if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================================================================== epilogue warps: this CTA's 128 rows
        const int quarter = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = cluster_id; t < total_tiles; t += n_clusters) {
            const int m0 = (t % tiles_m) * (2 * BM) + (int)cta_rank * BM, n0 = (t / tiles_m) * G2_BN;
            mbar_wait(&tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            const int row0 = m0 + quarter * 32;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * G2_BN);
            unsigned char* stg = smem + G2_STAGES * G2_STAGE_BYTES + 256 + (warp - 2) * EPI_STAGE_BYTES;
#pragma unroll 1
            for (int c = 0; c < G2_BN / 64; ++c) {
                uint32_t r0[32], r1[32];
                tmem_ld32(taddr + (uint32_t)(c * 64), r0);
                tmem_ld32(taddr + (uint32_t)(c * 64 + 32), r1);
                tmem_ld_wait();
                gemm_epilogue_chunk64(r0, r1, stg, Cv, row0, lane, n0 + c * 64, M, N, ldc, bias, residual, ldr, flags);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[acc]), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();                                   // no CTA frees tensor memory the pair's MMAs may still target
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc_2sm<G2_TMEM_COLS>(tmem_base);
    }
}

template <bool A_MN, bool B_MN>
static int launch_gemm2(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc, const void* bias,
                        const void* residual, int flags, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    int rc;
    rc = A_MN ? make_tensor_map(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK)
              : make_tensor_map(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
    if (rc != TL_OK) return rc;
    rc = B_MN ? make_tensor_map(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK)
              : make_tensor_map(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, G2_BN / 2);
    if (rc != TL_OK) return rc;
    auto kern = gemm2_bf16_kernel<A_MN, B_MN>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES) != cudaSuccess)
            return check_launch("tl_gemm_bf16/2cta (smem attr)");
        attr_done = true;
    }
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + G2_BN - 1) / G2_BN);
    int clusters = sm_count() / 2;
    if (clusters > tiles) clusters = tiles;
    kern<<<2 * clusters, G2_THREADS, G2_SMEM_BYTES, st>>>(tmA, tmB, C, M, N, K, ldc, (const bf16*)bias, (const bf16*)residual, ldc,
                                                          flags);
    return check_launch("tl_gemm_bf16/2cta");
}

// returns TL_OK / error, or 1 when the 1-CTA kernel should be used instead
int gemm2_dispatch(bool a_mn, bool b_mn, const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                   const void* bias, const void* residual, int flags, cudaStream_t st) {
    if (!a_mn && !b_mn) return launch_gemm2<false, false>(A, B, C, M, N, K, lda, ldb, ldc, bias, residual, flags, st);
    if (!a_mn && b_mn) return launch_gemm2<false, true>(A, B, C, M, N, K, lda, ldb, ldc, bias, residual, flags, st);
    if (a_mn && !b_mn) return launch_gemm2<true, false>(A, B, C, M, N, K, lda, ldb, ldc, bias, residual, flags, st);
    return launch_gemm2<true, true>(A, B, C, M, N, K, lda, ldb, ldc, bias, residual, flags, st);
}

}  // namespace tl
