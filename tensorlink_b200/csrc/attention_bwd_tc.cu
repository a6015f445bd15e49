// Causal GQA attention backward on tcgen05 / TMEM (training sequences of at least one 128-row tile; attention_bwd.cu keeps
// the mma.sync kernels for shorter runs and computes D = rowsum(dO * O) for both).
//
// Same two passes without atomics as attention_bwd.cu, with every product on the 5th-generation tensor cores:
//   dQ  kernel  CTA = 128 query rows of one head; loops over 64-key tiles up to the diagonal:
//                 S = Q K_j^T, dP = dO V_j^T            (tcgen05.mma, both accumulators in TMEM, double-buffered)
//                 dS = P * (dP - D) * scale -> bf16 -> shared memory (thread = query row, tcgen05.ld)
//                 dQ += dS K_j                          (K_j tile re-used as an MN-major B operand, accumulates in TMEM)
//   dKV kernel  CTA = 128 keys of one QUERY head (one partial per query head, summed over the GQA group by
//               rope_kv_bwd); loops over 64-query tiles from the diagonal down:
//                 S^T = K Q_i^T, dP^T = V dO_i^T        (thread = key row)
//                 P^T, dS^T -> bf16 -> shared memory
//                 dV += P^T dO_i, dK += dS^T Q_i        (Q_i / dO_i tiles re-used as MN-major B operands)
// Operand tiles are TMA-loaded, 128-byte swizzled, K-major in shared memory ([rows][64 columns] blocks); the same tile
// serves as a K-major operand of one product and as the MN-major operand of another through a second descriptor, like V
// in the forward kernel (attention_tc2.cu).  TMEM: 2 x 64 (S) + 2 x 64 (dP) + D (dQ)  /  + 2 D (dV, dK) columns of 512.
// Masked positions (key after query, rows beyond S: the TMA boxes run into the neighbouring sequence) get P = dS = 0.
#include <cuda.h>

#include "gemm_common.cuh"

namespace tl {

constexpr int BT_THREADS = 160;                 // 4 row warps + 1 control warp
constexpr float BT_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float bt_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void bt_bar_sync_rows() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// 64 bf16 of row `row` of a [128 rows][128 B] swizzled tile, from 64 floats
__device__ __forceinline__ void bt_store_row64(unsigned char* tile, int row, const float* v) {
    unsigned char* prow = tile + (row >> 3) * 1024 + (row & 7) * 128;
    const int sw = row & 7;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(prow + ((q ^ sw) << 4)) =
            make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                       pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
}

template <int D>
struct BtCfg {
    static constexpr int DB = D / 64;
    static constexpr int T128 = 128 * D * 2;        // a 128-row operand tile
    static constexpr int T64 = 64 * D * 2;          // a 64-row operand tile
    static constexpr int P_BYTES = 128 * 64 * 2;    // [128 rows][64 columns] bf16
    static constexpr int DQ_SMEM = 2 * T128 + 4 * T64 + P_BYTES + 256;                 // 144.25 KB at D = 128
    static constexpr int DKV_SMEM = 2 * T128 + 4 * T64 + 2 * P_BYTES + 1024 + 256;     // 161.25 KB at D = 128
    static constexpr uint32_t COL_S = 0, COL_DP = 128, COL_ACC0 = 256, COL_ACC1 = 256 + D;
};

// ------------------------------------------------------------------------------------------------ dQ
template <int D>
__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const float* __restrict__ lse, const float* __restrict__ Dv, bf16* __restrict__ dq, int S, int n_h, int n_kv,
                      int T_max, float scale) {
    using Cfg = BtCfg<D>;
    constexpr int DB = Cfg::DB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();
    unsigned char* sQ = smem_raw;                                  // [DB][128][128 B]
    unsigned char* sdO = sQ + Cfg::T128;
    unsigned char* sK = sdO + Cfg::T128;                           // [2][DB][64][128 B]
    unsigned char* sV = sK + 2 * Cfg::T64;
    unsigned char* sdS = sV + 2 * Cfg::T64;                        // [128][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sdS + Cfg::P_BYTES);
    uint64_t* bar_q = bars;            // 1
    uint64_t* bar_kv = bars + 1;       // 2
    uint64_t* bar_s = bars + 3;        // 2   S_j, dP_j in TMEM
    uint64_t* bar_sfree = bars + 5;    // 2   rows done reading S / dP buffer u          (128 arrivals)
    uint64_t* bar_p = bars + 7;        // 1   dS_j in shared memory                      (128 arrivals)
    uint64_t* bar_acc = bars + 8;      // 1   dQ += dS_j K_j done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;           // heaviest query tiles first
    const int h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (n_h / n_kv);
    const int q0 = qt * 128;
    const int n_tiles = (min(S, q0 + 128) + 63) / 64;
    const int kv_row0 = (b * n_kv + kvh) * T_max;

    if (tid == 0) {
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_kv[i], 1);
            mbar_init(&bar_s[i], 1);
            mbar_init(&bar_sfree[i], 128);
        }
        mbar_init(bar_p, 128);
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 4) {
        tmem_alloc<512>(tmem_slot);
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmdO);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            auto load_kv = [&](int j) {
                const int u = j & 1;
                mbar_expect_tx(&bar_kv[u], 2 * Cfg::T64);
#pragma unroll
                for (int db = 0; db < DB; ++db) {
                    tma_load_2d(sK + u * Cfg::T64 + db * 8192, &tmK, &bar_kv[u], 64 * db, kv_row0 + j * 64);
                    tma_load_2d(sV + u * Cfg::T64 + db * 8192, &tmV, &bar_kv[u], 64 * db, kv_row0 + j * 64);
                }
            };
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0u, 0u);
            constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, 0u, 1u);
            auto issue_s = [&](int j) {                                // S_j = Q K_j^T and dP_j = dO V_j^T
                const int u = j & 1;
                mbar_wait(&bar_kv[u], (uint32_t)(j >> 1) & 1u);
                if (j >= 2) mbar_wait(&bar_sfree[u], (uint32_t)((j >> 1) - 1) & 1u);
                tcgen05_fence_after();
                const uint64_t dq_ = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
                const uint64_t ddo = make_smem_desc_sw128(smem_u32(sdO), 16, 1024);
                const uint64_t dk_ = make_smem_desc_sw128(smem_u32(sK + u * Cfg::T64), 16, 1024);
                const uint64_t dv_ = make_smem_desc_sw128(smem_u32(sV + u * Cfg::T64), 16, 1024);
#pragma unroll
                for (int db = 0; db < DB; ++db)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offa = (uint64_t)((db * 16384 + 32 * k) >> 4);
                        const uint64_t offb = (uint64_t)((db * 8192 + 32 * k) >> 4);
                        umma_bf16(tmem_base + Cfg::COL_S + u * 64, dq_ + offa, dk_ + offb, idesc_s, (db | k) ? 1u : 0u);
                    }
#pragma unroll
                for (int db = 0; db < DB; ++db)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offa = (uint64_t)((db * 16384 + 32 * k) >> 4);
                        const uint64_t offb = (uint64_t)((db * 8192 + 32 * k) >> 4);
                        umma_bf16(tmem_base + Cfg::COL_DP + u * 64, ddo + offa, dv_ + offb, idesc_s, (db | k) ? 1u : 0u);
                    }
                umma_commit(&bar_s[u]);
            };
            mbar_expect_tx(bar_q, 2 * Cfg::T128);
#pragma unroll
            for (int db = 0; db < DB; ++db) {
                tma_load_2d(sQ + db * 16384, &tmQ, bar_q, h * D + 64 * db, b * S + q0);
                tma_load_2d(sdO + db * 16384, &tmdO, bar_q, h * D + 64 * db, b * S + q0);
            }
            load_kv(0);
            if (n_tiles > 1) load_kv(1);
            mbar_wait(bar_q, 0);
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int u = j & 1;
                if (j + 1 < n_tiles) issue_s(j + 1);
                mbar_wait(bar_p, (uint32_t)j & 1u);
                tcgen05_fence_after();
                const uint64_t dds = make_smem_desc_sw128(smem_u32(sdS), 16, 1024);
                const uint64_t dkm = make_smem_desc_sw128(smem_u32(sK + u * Cfg::T64), 8192, 1024);     // K_j as [keys][d], MN-major
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + Cfg::COL_ACC0, dds + (uint64_t)((32 * k) >> 4), dkm + (uint64_t)((2048 * k) >> 4), idesc_acc,
                              (j | k) ? 1u : 0u);
                umma_commit(bar_acc);
                if (j + 2 < n_tiles) {
                    mbar_wait(bar_acc, (uint32_t)j & 1u);              // everything issued so far is done: slot u is free
                    load_kv(j + 2);
                }
            }
        }
    } else {
        // ===================================================================== thread = query row
        const int row = tid;
        const int qpos = q0 + row;
        const bool live = qpos < S;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const float lse2 = live ? lse[((size_t)b * n_h + h) * S + qpos] * BT_LOG2E : 0.f;
        const float dlt = live ? Dv[((size_t)b * n_h + h) * S + qpos] : 0.f;
        const float sl2 = scale * BT_LOG2E;
        for (int j = 0; j < n_tiles; ++j) {
            const int u = j & 1;
            const int k0 = j * 64;
            mbar_wait(&bar_s[u], (uint32_t)(j >> 1) & 1u);
            tcgen05_fence_after();
            uint32_t s[64], dp[64];
            tmem_ld32(tmem_base + lane_base + Cfg::COL_S + u * 64, s);
            tmem_ld32(tmem_base + lane_base + Cfg::COL_S + u * 64 + 32, s + 32);
            tmem_ld32(tmem_base + lane_base + Cfg::COL_DP + u * 64, dp);
            tmem_ld32(tmem_base + lane_base + Cfg::COL_DP + u * 64 + 32, dp + 32);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&bar_sfree[u]);
            float ds[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const int key = k0 + i;
                const bool ok = live && key <= qpos && key < S;
                const float p = ok ? bt_exp2(__uint_as_float(s[i]) * sl2 - lse2) : 0.f;
                ds[i] = ok ? p * (__uint_as_float(dp[i]) - dlt) * scale : 0.f;
            }
            if (j > 0) mbar_wait(bar_acc, (uint32_t)(j - 1) & 1u);    // dS_{j-1} has been consumed
            bt_store_row64(sdS, row, ds);
            fence_proxy_async();
            mbar_arrive(bar_p);
        }
        mbar_wait(bar_acc, (uint32_t)(n_tiles - 1) & 1u);
        tcgen05_fence_after();
        bf16* dst = dq + ((size_t)b * S + qpos) * n_h * D + (size_t)h * D;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_base + lane_base + Cfg::COL_ACC0 + c * 32, o);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) =
                        make_uint4(pack_bf16(__uint_as_float(o[8 * q]), __uint_as_float(o[8 * q + 1])),
                                   pack_bf16(__uint_as_float(o[8 * q + 2]), __uint_as_float(o[8 * q + 3])),
                                   pack_bf16(__uint_as_float(o[8 * q + 4]), __uint_as_float(o[8 * q + 5])),
                                   pack_bf16(__uint_as_float(o[8 * q + 6]), __uint_as_float(o[8 * q + 7])));
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------ dK, dV
template <int D>
__global__ void __launch_bounds__(BT_THREADS, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const float* __restrict__ lse, const float* __restrict__ Dv, bf16* __restrict__ dk, bf16* __restrict__ dv, int S,
                       int n_h, int n_kv, int T_max, float scale) {
    using Cfg = BtCfg<D>;
    constexpr int DB = Cfg::DB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();
    unsigned char* sK = smem_raw;                                  // [DB][128][128 B]
    unsigned char* sV = sK + Cfg::T128;
    unsigned char* sQ = sV + Cfg::T128;                            // [2][DB][64][128 B]
    unsigned char* sdO = sQ + 2 * Cfg::T64;
    unsigned char* sP = sdO + 2 * Cfg::T64;                        // [128 keys][64 queries]
    unsigned char* sdS = sP + Cfg::P_BYTES;
    float* s_lse = reinterpret_cast<float*>(sdS + Cfg::P_BYTES);   // [2][64] lse * log2(e)
    float* s_dlt = s_lse + 128;                                    // [2][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_lse + 256);
    uint64_t* bar_kv = bars;           // 1
    uint64_t* bar_q = bars + 1;        // 2
    uint64_t* bar_s = bars + 3;        // 2
    uint64_t* bar_sfree = bars + 5;    // 2   (128 arrivals)
    uint64_t* bar_p = bars + 7;        // 1   (128 arrivals)
    uint64_t* bar_acc = bars + 8;      // 1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;     // (the first key tiles see the most query tiles: heaviest first)
    const int kvh = h / (n_h / n_kv);
    const int kv0 = kt * 128;
    const int i0 = kv0 / 64;                                       // first 64-query tile that can see a key of this tile
    const int n_it = (S + 63) / 64 - i0;
    const int kv_row0 = (b * n_kv + kvh) * T_max;

    if (tid == 0) {
        mbar_init(bar_kv, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_q[i], 1);
            mbar_init(&bar_s[i], 1);
            mbar_init(&bar_sfree[i], 128);
        }
        mbar_init(bar_p, 128);
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 4) {
        tmem_alloc<512>(tmem_slot);
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmdO);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            auto load_q = [&](int it) {
                const int u = it & 1;
                mbar_expect_tx(&bar_q[u], 2 * Cfg::T64);
#pragma unroll
                for (int db = 0; db < DB; ++db) {
                    tma_load_2d(sQ + u * Cfg::T64 + db * 8192, &tmQ, &bar_q[u], h * D + 64 * db, b * S + (i0 + it) * 64);
                    tma_load_2d(sdO + u * Cfg::T64 + db * 8192, &tmdO, &bar_q[u], h * D + 64 * db, b * S + (i0 + it) * 64);
                }
            };
            constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0u, 0u);
            constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, 0u, 1u);
            auto issue_s = [&](int it) {                               // S^T = K Q_i^T and dP^T = V dO_i^T
                const int u = it & 1;
                mbar_wait(&bar_q[u], (uint32_t)(it >> 1) & 1u);
                if (it >= 2) mbar_wait(&bar_sfree[u], (uint32_t)((it >> 1) - 1) & 1u);
                tcgen05_fence_after();
                const uint64_t dk_ = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
                const uint64_t dv_ = make_smem_desc_sw128(smem_u32(sV), 16, 1024);
                const uint64_t dq_ = make_smem_desc_sw128(smem_u32(sQ + u * Cfg::T64), 16, 1024);
                const uint64_t ddo = make_smem_desc_sw128(smem_u32(sdO + u * Cfg::T64), 16, 1024);
#pragma unroll
                for (int db = 0; db < DB; ++db)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offa = (uint64_t)((db * 16384 + 32 * k) >> 4);
                        const uint64_t offb = (uint64_t)((db * 8192 + 32 * k) >> 4);
                        umma_bf16(tmem_base + Cfg::COL_S + u * 64, dk_ + offa, dq_ + offb, idesc_s, (db | k) ? 1u : 0u);
                    }
#pragma unroll
                for (int db = 0; db < DB; ++db)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offa = (uint64_t)((db * 16384 + 32 * k) >> 4);
                        const uint64_t offb = (uint64_t)((db * 8192 + 32 * k) >> 4);
                        umma_bf16(tmem_base + Cfg::COL_DP + u * 64, dv_ + offa, ddo + offb, idesc_s, (db | k) ? 1u : 0u);
                    }
                umma_commit(&bar_s[u]);
            };
            mbar_expect_tx(bar_kv, 2 * Cfg::T128);
#pragma unroll
            for (int db = 0; db < DB; ++db) {
                tma_load_2d(sK + db * 16384, &tmK, bar_kv, 64 * db, kv_row0 + kv0);
                tma_load_2d(sV + db * 16384, &tmV, bar_kv, 64 * db, kv_row0 + kv0);
            }
            load_q(0);
            if (n_it > 1) load_q(1);
            mbar_wait(bar_kv, 0);
            issue_s(0);
            for (int it = 0; it < n_it; ++it) {
                const int u = it & 1;
                if (it + 1 < n_it) issue_s(it + 1);
                mbar_wait(bar_p, (uint32_t)it & 1u);
                tcgen05_fence_after();
                const uint64_t dp_ = make_smem_desc_sw128(smem_u32(sP), 16, 1024);
                const uint64_t dds = make_smem_desc_sw128(smem_u32(sdS), 16, 1024);
                const uint64_t dom = make_smem_desc_sw128(smem_u32(sdO + u * Cfg::T64), 8192, 1024);   // dO_i as [queries][d], MN-major
                const uint64_t dqm = make_smem_desc_sw128(smem_u32(sQ + u * Cfg::T64), 8192, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + Cfg::COL_ACC0, dp_ + (uint64_t)((32 * k) >> 4), dom + (uint64_t)((2048 * k) >> 4), idesc_acc,
                              (it | k) ? 1u : 0u);                     // dV += P^T dO_i
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + Cfg::COL_ACC1, dds + (uint64_t)((32 * k) >> 4), dqm + (uint64_t)((2048 * k) >> 4), idesc_acc,
                              (it | k) ? 1u : 0u);                     // dK += dS^T Q_i
                umma_commit(bar_acc);
                if (it + 2 < n_it) {
                    mbar_wait(bar_acc, (uint32_t)it & 1u);
                    load_q(it + 2);
                }
            }
        }
    } else {
        // ===================================================================== thread = key row
        const int row = tid;
        const int kpos = kv0 + row;
        const bool live = kpos < S;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const float sl2 = scale * BT_LOG2E;
        const float* lse_h = lse + ((size_t)b * n_h + h) * S;
        const float* dlt_h = Dv + ((size_t)b * n_h + h) * S;
        for (int it = 0; it < n_it; ++it) {
            const int u = it & 1;
            const int q0 = (i0 + it) * 64;
            {   // per-query lse / D of this tile (buffer u was last read two tiles ago, before the previous tile's bar.sync)
                const int c = tid & 63, qp = q0 + c;
                if (tid < 64) s_lse[u * 64 + c] = qp < S ? lse_h[qp] * BT_LOG2E : 0.f;
                else s_dlt[u * 64 + c] = qp < S ? dlt_h[qp] : 0.f;
            }
            bt_bar_sync_rows();
            mbar_wait(&bar_s[u], (uint32_t)(it >> 1) & 1u);
            tcgen05_fence_after();
            uint32_t s[64], dp[64];
            tmem_ld32(tmem_base + lane_base + Cfg::COL_S + u * 64, s);
            tmem_ld32(tmem_base + lane_base + Cfg::COL_S + u * 64 + 32, s + 32);
            tmem_ld32(tmem_base + lane_base + Cfg::COL_DP + u * 64, dp);
            tmem_ld32(tmem_base + lane_base + Cfg::COL_DP + u * 64 + 32, dp + 32);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&bar_sfree[u]);
            float p[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const int qp = q0 + i;
                const bool ok = live && qp >= kpos && qp < S;
                p[i] = ok ? bt_exp2(__uint_as_float(s[i]) * sl2 - s_lse[u * 64 + i]) : 0.f;
            }
            if (it > 0) mbar_wait(bar_acc, (uint32_t)(it - 1) & 1u);  // P^T / dS^T of the previous tile have been consumed
            bt_store_row64(sP, row, p);
#pragma unroll
            for (int i = 0; i < 64; ++i) p[i] = p[i] * (__uint_as_float(dp[i]) - s_dlt[u * 64 + i]) * scale;   // (P = 0 where masked)
            bt_store_row64(sdS, row, p);
            fence_proxy_async();
            mbar_arrive(bar_p);
        }
        mbar_wait(bar_acc, (uint32_t)(n_it - 1) & 1u);
        tcgen05_fence_after();
        bf16* dv_dst = dv + (((size_t)b * n_h + h) * T_max + kpos) * D;
        bf16* dk_dst = dk + (((size_t)b * n_h + h) * T_max + kpos) * D;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            bf16* dst = half ? dk_dst : dv_dst;
            const uint32_t col = half ? Cfg::COL_ACC1 : Cfg::COL_ACC0;
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                uint32_t o[32];
                tmem_ld32(tmem_base + lane_base + col + c * 32, o);
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) =
                            make_uint4(pack_bf16(__uint_as_float(o[8 * q]), __uint_as_float(o[8 * q + 1])),
                                       pack_bf16(__uint_as_float(o[8 * q + 2]), __uint_as_float(o[8 * q + 3])),
                                       pack_bf16(__uint_as_float(o[8 * q + 4]), __uint_as_float(o[8 * q + 5])),
                                       pack_bf16(__uint_as_float(o[8 * q + 6]), __uint_as_float(o[8 * q + 7])));
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int D>
static int launch_bwd_tc(const void* q, const void* k_cache, const void* v_cache, const void* dout, const float* lse, const float* Dv,
                         void* dq, void* dk, void* dv, int B, int S, int n_h, int n_kv, int T_max, float scale, cudaStream_t st) {
    using Cfg = BtCfg<D>;
    CUtensorMap tmQ128, tmdO128, tmK64, tmV64, tmQ64, tmdO64, tmK128, tmV128;
    int rc;
    if ((rc = make_tensor_map(&tmQ128, q, (uint64_t)n_h * D, (uint64_t)B * S, (uint64_t)n_h * D, 64, 128)) != TL_OK) return rc;
    if ((rc = make_tensor_map(&tmdO128, dout, (uint64_t)n_h * D, (uint64_t)B * S, (uint64_t)n_h * D, 64, 128)) != TL_OK) return rc;
    if ((rc = make_tensor_map(&tmK64, k_cache, (uint64_t)D, (uint64_t)B * n_kv * T_max, (uint64_t)D, 64, 64)) != TL_OK) return rc;
    if ((rc = make_tensor_map(&tmV64, v_cache, (uint64_t)D, (uint64_t)B * n_kv * T_max, (uint64_t)D, 64, 64)) != TL_OK) return rc;
    if ((rc = make_tensor_map(&tmQ64, q, (uint64_t)n_h * D, (uint64_t)B * S, (uint64_t)n_h * D, 64, 64)) != TL_OK) return rc;
    if ((rc = make_tensor_map(&tmdO64, dout, (uint64_t)n_h * D, (uint64_t)B * S, (uint64_t)n_h * D, 64, 64)) != TL_OK) return rc;
    if ((rc = make_tensor_map(&tmK128, k_cache, (uint64_t)D, (uint64_t)B * n_kv * T_max, (uint64_t)D, 64, 128)) != TL_OK) return rc;
    if ((rc = make_tensor_map(&tmV128, v_cache, (uint64_t)D, (uint64_t)B * n_kv * T_max, (uint64_t)D, 64, 128)) != TL_OK) return rc;
    auto k1 = attn_bwd_dq_tc_kernel<D>;
    auto k2 = attn_bwd_dkv_tc_kernel<D>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::DQ_SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::DKV_SMEM) != cudaSuccess)
            return check_launch("tl_attn_bwd/tc (smem attr)");
        attr_done = true;
    }
    const dim3 grid((S + 127) / 128, n_h, B);
    k1<<<grid, BT_THREADS, Cfg::DQ_SMEM, st>>>(tmQ128, tmdO128, tmK64, tmV64, lse, Dv, (bf16*)dq, S, n_h, n_kv, T_max, scale);
    k2<<<grid, BT_THREADS, Cfg::DKV_SMEM, st>>>(tmQ64, tmdO64, tmK128, tmV128, lse, Dv, (bf16*)dk, (bf16*)dv, S, n_h, n_kv, T_max, scale);
    return check_launch("tl_attn_bwd/tc");
}

// returns TL_OK / error, or 1 = not applicable (the caller runs the mma.sync kernels)
int attn_bwd_tc_dispatch(const void* q, const void* k_cache, const void* v_cache, const void* dout, const float* lse, const float* Dv,
                         void* dq, void* dk, void* dv, int B, int S, int n_h, int n_kv, int d, int T_max, float scale,
                         cudaStream_t st) {
    if ((((uintptr_t)q | (uintptr_t)k_cache | (uintptr_t)v_cache | (uintptr_t)dout | (uintptr_t)dq | (uintptr_t)dk | (uintptr_t)dv) & 15) != 0)
        return 1;
    if (d == 128) return launch_bwd_tc<128>(q, k_cache, v_cache, dout, lse, Dv, dq, dk, dv, B, S, n_h, n_kv, T_max, scale, st);
    if (d == 64) return launch_bwd_tc<64>(q, k_cache, v_cache, dout, lse, Dv, dq, dk, dv, B, S, n_h, n_kv, T_max, scale, st);
    return 1;
}

}  // namespace tl
