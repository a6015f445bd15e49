// Causal GQA attention forward (prefill / training) on the 5th-generation tensor cores.
//
// Replaces the `mma.sync` tiles of attention.cu for the SDPA call of Qwen2Attention.forward
// (site-packages/transformers/models/qwen2/modeling_qwen2.py:161-184, 232-243).  One CTA = 128 query rows of one head;
// it walks the causal range of the KV cache in tiles of 128 keys:
//
//   control warp (one elected lane):  TMA loads Q once and K_j / V_j into a 2-deep ring (128B-swizzled boxes),
//                                     S_j = Q K_j^T      tcgen05.mma 128x128xD   -> TMEM S[j & 1]
//                                     O_j = P_j V_j      tcgen05.mma 128xDx128   -> TMEM O   (P from shared memory, V MN-major)
//   4 softmax warps (thread = row):   tcgen05.ld S_j, scale, causal mask, running max / sum (fp32, exp2), P_j -> bf16 ->
//                                     shared memory in the swizzled K-major operand layout, then tcgen05.ld O_j and
//                                     O = O * alpha + O_j in registers.
//
// S_{j+1} is issued before the softmax of tile j starts, so the tensor pipe computes the next scores while the SFU works.
// Numerics follow attention.cu: P is rounded to bf16 before P V (SDPA contract), the row sum uses the unrounded fp32 p,
// lse = m ln2 + log(l).
#include <cuda.h>

#include "gemm_common.cuh"

namespace tl {

constexpr int TC_BQ = 128;
constexpr int TC_THREADS = 160;                 // 4 softmax warps + 1 control warp

// BKV = 128: one CTA per SM (192 KB of shared memory at D = 128, 512 TMEM columns).
// BKV = 64:  112 KB and 256 TMEM columns, so TWO CTAs share an SM: while one CTA's softmax warps work (one warp per
//            scheduler cannot hide its own TMEM / SFU latencies) the other CTA's MMAs and softmax fill the gaps.
template <int D, int BKV>
struct TcCfg {
    static constexpr int Q_BYTES = 128 * D * 2;
    static constexpr int KV_BYTES = BKV * D * 2;                    // one K or V tile
    static constexpr int P_BYTES = 128 * BKV * 2;
    // no alignment slack: two CTAs of 112 KB + barriers + the 1 KB the system reserves per CTA must fit in 228 KB; the
    // dynamic segment is declared 1024-byte aligned (the kernel has no static shared memory) and the kernel traps if not
    static constexpr int SMEM_BYTES = Q_BYTES + 4 * KV_BYTES + P_BYTES + 256 /*barriers*/;
    static constexpr uint32_t TMEM_COLS = (BKV == 128) ? 512 : 256;
    static constexpr uint32_t COL_S0 = 0, COL_S1 = BKV, COL_O = 2 * BKV;
    static constexpr int CTAS_PER_SM = (BKV == 128) ? 1 : 2;
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D, int BKV>
__global__ void __launch_bounds__(TC_THREADS, TcCfg<D, BKV>::CTAS_PER_SM)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out, float* __restrict__ lse, int S, int past_len,
                       int n_h, int n_kv, int T_max, float scale_log2) {
    using Cfg = TcCfg<D, BKV>;
    constexpr int DB = D / 64;                                       // 64-element blocks along the head dimension
    constexpr int KB = BKV / 64;                                     // 64-key blocks per KV tile
    constexpr int TC_BKV = BKV;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();                        // swizzled tiles need a 1024-byte aligned base
    unsigned char* smem = smem_raw;
    unsigned char* sQ = smem;
    unsigned char* sK = sQ + Cfg::Q_BYTES;                           // [2][KV tile]: [D/64 blocks][BKV keys][128 B]
    unsigned char* sV = sK + 2 * Cfg::KV_BYTES;                      // [2][KV tile]: [key block][D/64 blocks][64 keys][128 B]
    unsigned char* sP = sV + 2 * Cfg::KV_BYTES;                      // [key block][128 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
    uint64_t* bar_q = bars;            // 1
    uint64_t* bar_k = bars + 1;        // 2
    uint64_t* bar_v = bars + 3;        // 2
    uint64_t* bar_s = bars + 5;        // 2   S_j in TMEM
    uint64_t* bar_sfree = bars + 7;    // 2   softmax done reading S[u]            (128 arrivals)
    uint64_t* bar_p = bars + 9;        // 1   P_j in shared memory                 (128 arrivals)
    uint64_t* bar_o = bars + 10;       // 1   O_j in TMEM
    uint64_t* bar_ofree = bars + 11;   // 1   softmax done reading O              (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;             // heaviest (last) query tiles first
    const int h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (n_h / n_kv);
    const int q0 = qt * TC_BQ;
    const int T = past_len + S;
    const int kv_end = min(T, past_len + q0 + TC_BQ);
    const int n_tiles = (kv_end + TC_BKV - 1) / TC_BKV;
    const int kv_row0 = (b * n_kv + kvh) * T_max;                    // row of key 0 in the [B*n_kv*T_max, D] view

    if (tid == 0) {
        mbar_init(bar_q, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_k[i], 1);
            mbar_init(&bar_v[i], 1);
            mbar_init(&bar_s[i], 1);
            mbar_init(&bar_sfree[i], 128);
        }
        mbar_init(bar_p, 128);
        mbar_init(bar_o, 1);
        mbar_init(bar_ofree, 128);
        fence_barrier_init();
    }
    if (warp == 4) {
        tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmV);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===================================================================== control: TMA + MMA issue (one lane)
        if (lane == 0) {
            auto load_kv = [&](int j) {
                const int u = j & 1;
                mbar_expect_tx(&bar_k[u], Cfg::KV_BYTES);
#pragma unroll
                for (int db = 0; db < DB; ++db)
                    tma_load_2d(sK + u * Cfg::KV_BYTES + db * (BKV * 128), &tmK, &bar_k[u], 64 * db, kv_row0 + j * TC_BKV);
                mbar_expect_tx(&bar_v[u], Cfg::KV_BYTES);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                    for (int nb = 0; nb < DB; ++nb)
                        tma_load_2d(sV + u * Cfg::KV_BYTES + (kb * DB + nb) * 8192, &tmV, &bar_v[u], 64 * nb,
                                    kv_row0 + j * TC_BKV + 64 * kb);
            };
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BKV, 0u, 0u);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, D, 0u, 1u);
            auto issue_s = [&](int j) {
                const int u = j & 1;
                mbar_wait(&bar_k[u], (uint32_t)(j >> 1) & 1u);
                if (j >= 2) mbar_wait(&bar_sfree[u], (uint32_t)((j >> 1) - 1) & 1u);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (u ? Cfg::COL_S1 : Cfg::COL_S0);
                const uint64_t da = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
                const uint64_t dbk = make_smem_desc_sw128(smem_u32(sK + u * Cfg::KV_BYTES), 16, 1024);
#pragma unroll
                for (int db = 0; db < DB; ++db)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t offq = (uint64_t)((db * 16384 + 32 * k) >> 4);
                        const uint64_t offk = (uint64_t)((db * (BKV * 128) + 32 * k) >> 4);
                        umma_bf16(d_tmem, da + offq, dbk + offk, idesc_s, (db | k) ? 1u : 0u);
                    }
                umma_commit(&bar_s[u]);
            };
            mbar_expect_tx(bar_q, Cfg::Q_BYTES);
#pragma unroll
            for (int db = 0; db < DB; ++db) tma_load_2d(sQ + db * 16384, &tmQ, bar_q, h * D + 64 * db, b * S + q0);
            load_kv(0);
            if (n_tiles > 1) load_kv(1);
            mbar_wait(bar_q, 0);
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int u = j & 1;
                if (j + 1 < n_tiles) issue_s(j + 1);
                mbar_wait(bar_p, (uint32_t)j & 1u);
                mbar_wait(&bar_v[u], (uint32_t)(j >> 1) & 1u);
                if (j >= 1) mbar_wait(bar_ofree, (uint32_t)(j - 1) & 1u);
                tcgen05_fence_after();
                const uint64_t dp = make_smem_desc_sw128(smem_u32(sP), 16, 1024);
                const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + u * Cfg::KV_BYTES), 8192, 1024);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t pa = dp + (uint64_t)((kb * 16384 + 32 * k) >> 4);
                        const uint64_t vb = dv + (uint64_t)((kb * DB * 8192 + 2048 * k) >> 4);
                        umma_bf16(tmem_base + Cfg::COL_O, pa, vb, idesc_o, (kb | k) ? 1u : 0u);
                    }
                umma_commit(bar_o);
                if (j + 2 < n_tiles) {             // ring slot u is free once O_j (and S_j long before it) has been computed
                    mbar_wait(bar_o, (uint32_t)j & 1u);
                    load_kv(j + 2);
                }
            }
        }
    } else {
        // ===================================================================== softmax warps: thread = query row
        const int row = tid;                                          // == TMEM lane
        const int qpos = past_len + q0 + row;                         // absolute position of this query
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        float o[D];
#pragma unroll
        for (int i = 0; i < D; ++i) o[i] = 0.f;
        unsigned char* prow = sP + (row >> 3) * 1024 + (row & 7) * 128;
        const int sw = row & 7;

        for (int j = 0; j < n_tiles; ++j) {
            const int u = j & 1;
            const uint32_t ts = tmem_base + lane_base + (u ? Cfg::COL_S1 : Cfg::COL_S0);
            const int k0 = j * TC_BKV;
            const bool need_mask = (k0 + TC_BKV - 1 > past_len + q0) || (k0 + TC_BKV > T);   // warp-uniform
            mbar_wait(&bar_s[u], (uint32_t)(j >> 1) & 1u);
            tcgen05_fence_after();
            float mx = -INFINITY, rs = 0.f;
            float m_new, m_use, alpha;
            if constexpr (BKV == 64 && D == 64) {
                // ---- one pass: the whole score row (64 fp32) stays in registers beside the 64 output accumulators; both
                // TMEM loads are issued before the wait.  (At D = 128 the extra 64 registers spill: 292 vs 319 TFLOP/s at
                // S = 4096, so that case keeps the two-pass form below.)
                uint32_t r[64];
                tmem_ld32(ts, r);
                tmem_ld32(ts + 32, r + 32);
                tmem_ld_wait();
                tcgen05_fence_before();
                mbar_arrive(&bar_sfree[u]);                           // S[u] may be overwritten by S_{j+2}
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const int kp = k0 + i;
                    if (need_mask && !(kp <= qpos && kp < T)) r[i] = __float_as_uint(-INFINITY);
                    mx = fmaxf(mx, __uint_as_float(r[i]));
                }
                m_new = fmaxf(m_run, mx * scale_log2);
                m_use = (m_new == -INFINITY) ? 0.f : m_new;           // fully masked so far (rows beyond S): keep p = 0
                alpha = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run - m_use);
#pragma unroll
                for (int q = 0; q < 8; ++q) {                         // eight 16-byte chunks (8 keys each)
                    float p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        p[i] = fast_exp2(__uint_as_float(r[8 * q + i]) * scale_log2 - m_use);   // exp2(-inf) = 0 for masked keys
                        rs += p[i];
                    }
                    *reinterpret_cast<uint4*>(prow + ((q ^ sw) << 4)) =
                        make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
                }
            } else {
                // ---- pass A: row maximum
#pragma unroll
                for (int c = 0; c < BKV / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(ts + c * 32, r);
                    tmem_ld_wait();
                    if (need_mask) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kp = k0 + c * 32 + i;
                            if (kp <= qpos && kp < T) mx = fmaxf(mx, __uint_as_float(r[i]));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
                    }
                }
                m_new = fmaxf(m_run, mx * scale_log2);
                m_use = (m_new == -INFINITY) ? 0.f : m_new;
                alpha = (m_run == -INFINITY) ? 0.f : fast_exp2(m_run - m_use);
                // ---- pass B: p = exp2(s * scale_log2 - m), P tile (bf16) to shared memory, row sum
#pragma unroll
                for (int c = 0; c < BKV / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(ts + c * 32, r);
                    tmem_ld_wait();
                    float p[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int kp = k0 + c * 32 + i;
                        const bool ok = !need_mask || (kp <= qpos && kp < T);
                        p[i] = ok ? fast_exp2(__uint_as_float(r[i]) * scale_log2 - m_use) : 0.f;
                        rs += p[i];
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {                     // four 16-byte chunks (8 keys each)
                        const int cc = c * 4 + q;                     // chunk index 0..15 within the row
                        const int kb = cc >> 3, within = cc & 7;
                        *reinterpret_cast<uint4*>(prow + kb * 16384 + ((within ^ sw) << 4)) =
                            make_uint4(pack_bf16(p[8 * q], p[8 * q + 1]), pack_bf16(p[8 * q + 2], p[8 * q + 3]),
                                       pack_bf16(p[8 * q + 4], p[8 * q + 5]), pack_bf16(p[8 * q + 6], p[8 * q + 7]));
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(&bar_sfree[u]);                           // S[u] may be overwritten by S_{j+2}
            }
            fence_proxy_async();                                      // P stores -> visible to the tensor core's reads
            mbar_arrive(bar_p);
            l_run = l_run * alpha + rs;
            m_run = m_new;
            // ---- O = O * alpha + P_j V_j
            mbar_wait(bar_o, (uint32_t)j & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < D / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + Cfg::COL_O + c * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] = o[c * 32 + i] * alpha + __uint_as_float(r[i]);
            }
            tcgen05_fence_before();
            mbar_arrive(bar_ofree);
        }
        // ---- normalise and store this row
        if (q0 + row < S) {
            const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
            bf16* dst = out + ((size_t)b * S + q0 + row) * n_h * D + (size_t)h * D;
#pragma unroll
            for (int c = 0; c < D / 8; ++c)
                *reinterpret_cast<uint4*>(dst + c * 8) =
                    make_uint4(pack_bf16(o[8 * c] * inv, o[8 * c + 1] * inv), pack_bf16(o[8 * c + 2] * inv, o[8 * c + 3] * inv),
                               pack_bf16(o[8 * c + 4] * inv, o[8 * c + 5] * inv), pack_bf16(o[8 * c + 6] * inv, o[8 * c + 7] * inv));
            if (lse) lse[((size_t)b * n_h + h) * S + q0 + row] = m_run * 0.6931471805599453f + logf(l_run);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) {
        tcgen05_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int D, int BKV>
static int launch_tc(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, int B, int S, int past_len,
                     int n_h, int n_kv, int T_max, float scale, cudaStream_t st) {
    using Cfg = TcCfg<D, BKV>;
    CUtensorMap tmQ, tmK, tmV;
    int rc = make_tensor_map(&tmQ, q, (uint64_t)n_h * D, (uint64_t)B * S, (uint64_t)n_h * D, 64, 128);
    if (rc != TL_OK) return rc;
    rc = make_tensor_map(&tmK, k_cache, (uint64_t)D, (uint64_t)B * n_kv * T_max, (uint64_t)D, 64, BKV);
    if (rc != TL_OK) return rc;
    rc = make_tensor_map(&tmV, v_cache, (uint64_t)D, (uint64_t)B * n_kv * T_max, (uint64_t)D, 64, 64);
    if (rc != TL_OK) return rc;
    auto kern = attn_prefill_tc_kernel<D, BKV>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
            return check_launch("tl_attn_prefill_fwd/tc (smem attr)");
        attr_done = true;
    }
    const dim3 grid((S + TC_BQ - 1) / TC_BQ, n_h, B);
    kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(tmQ, tmK, tmV, (bf16*)out, lse, S, past_len, n_h, n_kv, T_max,
                                                   scale * 1.4426950408889634f);
    return check_launch("tl_attn_prefill_fwd/tc");
}

// returns TL_OK / error, or 1 = not applicable (caller uses the mma.sync kernel)
int attn_prefill_tc_dispatch(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, int B, int S,
                             int past_len, int n_h, int n_kv, int d, int T_max, float scale, cudaStream_t st) {
    if ((((uintptr_t)q | (uintptr_t)k_cache | (uintptr_t)v_cache | (uintptr_t)out) & 15) != 0) return 1;
    const char* e = getenv("TL_ATTN_BKV");                  // 64 (default: two CTAs per SM) or 128
    const bool wide = e && e[0] == '1';
    if (d == 128) return wide ? launch_tc<128, 128>(q, k_cache, v_cache, out, lse, B, S, past_len, n_h, n_kv, T_max, scale, st)
                              : launch_tc<128, 64>(q, k_cache, v_cache, out, lse, B, S, past_len, n_h, n_kv, T_max, scale, st);
    if (d == 64) return wide ? launch_tc<64, 128>(q, k_cache, v_cache, out, lse, B, S, past_len, n_h, n_kv, T_max, scale, st)
                             : launch_tc<64, 64>(q, k_cache, v_cache, out, lse, B, S, past_len, n_h, n_kv, T_max, scale, st);
    return 1;
}

}  // namespace tl
