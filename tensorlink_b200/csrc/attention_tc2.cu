// Causal GQA attention forward on tcgen05 / TMEM, second form: the output accumulator never leaves tensor memory.
//
// Same tiling, operand layouts and barriers as attention_tc.cu (128 query rows x 64-key tiles, two CTAs per SM), but
//   * P_j V_j is issued with accumulate = 1 into ONE TMEM accumulator O, so the tensor core sums the tiles itself and
//     the softmax warps neither read O every tile nor hold 64-128 fp32 of it in registers;
//   * the exponent reference m_ref follows the running row maximum lazily: p = exp2(s*c - m_ref), and only when a row's
//     maximum has grown by more than 8 (log2 units; p <= 256 until then) is that row of O rescaled in place
//     (tcgen05.ld -> * 2^(m_ref - m_new) -> tcgen05.st, decided per warp) - typically in the first one or two tiles of a
//     row and never again.  P is still rounded to bf16 before P V and the row sum still uses the unrounded fp32 p; only
//     the reference of the exponent differs, so results agree with attention.cu to the usual bf16-P noise (1e-3 rel.);
//   * the whole 64-value score row is read in one pass (both tcgen05.ld issued before the wait).
// Per tile the dependent chain is therefore softmax only: P_j V_j runs under the softmax of tile j+1.
#include <cuda.h>

#include "gemm_common.cuh"

namespace tl {

constexpr int T2_BQ = 128, T2_BKV = 64;
constexpr int T2_THREADS = 160;                 // 4 softmax warps + 1 control warp
constexpr float T2_RESCALE = 8.0f;              // log2 units

template <int D>
struct T2Cfg {
    static constexpr int Q_BYTES = 128 * D * 2;
    static constexpr int KV_BYTES = T2_BKV * D * 2;
    static constexpr int P_BYTES = 128 * T2_BKV * 2;
    static constexpr int SMEM_BYTES = Q_BYTES + 4 * KV_BYTES + P_BYTES + 256 /*barriers*/;   // 112.25 KB at D = 128: 2 CTAs / SM
    static constexpr uint32_t TMEM_COLS = 256;
    static constexpr uint32_t COL_S0 = 0, COL_S1 = 64, COL_O = 128;
};

__device__ __forceinline__ float t2_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 32 lanes x 32 columns of fp32 back into tensor memory (thread i writes row lane_base + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int D>
__global__ void __launch_bounds__(T2_THREADS, 2)
attn_prefill_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ out, float* __restrict__ lse, int S, int past_len,
                        int n_h, int n_kv, int T_max, float scale_log2) {
    using Cfg = T2Cfg<D>;
    constexpr int DB = D / 64;
    constexpr int BKV = T2_BKV;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (smem_u32(smem_raw) & 1023u) __trap();                        // swizzled tiles need a 1024-byte aligned base
    unsigned char* sQ = smem_raw;
    unsigned char* sK = sQ + Cfg::Q_BYTES;                           // [2][D/64 blocks][64 keys][128 B]
    unsigned char* sV = sK + 2 * Cfg::KV_BYTES;                      // [2][D/64 blocks][64 keys][128 B]  (MN-major operand)
    unsigned char* sP = sV + 2 * Cfg::KV_BYTES;                      // [128 rows][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
    uint64_t* bar_q = bars;            // 1
    uint64_t* bar_k = bars + 1;        // 2
    uint64_t* bar_v = bars + 3;        // 2
    uint64_t* bar_s = bars + 5;        // 2   S_j in TMEM
    uint64_t* bar_sfree = bars + 7;    // 2   softmax done reading S[u]            (128 arrivals)
    uint64_t* bar_p = bars + 9;        // 1   P_j in shared memory, O rescaled     (128 arrivals)
    uint64_t* bar_o = bars + 10;       // 1   P_j V_j accumulated into O
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qt = (int)gridDim.x - 1 - (int)blockIdx.x;             // heaviest (last) query tiles first
    const int h = blockIdx.y, b = blockIdx.z;
    const int kvh = h / (n_h / n_kv);
    const int q0 = qt * T2_BQ;
    const int T = past_len + S;
    const int kv_end = min(T, past_len + q0 + T2_BQ);
    const int n_tiles = (kv_end + BKV - 1) / BKV;
    const int kv_row0 = (b * n_kv + kvh) * T_max;

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
                    tma_load_2d(sK + u * Cfg::KV_BYTES + db * 8192, &tmK, &bar_k[u], 64 * db, kv_row0 + j * BKV);
                mbar_expect_tx(&bar_v[u], Cfg::KV_BYTES);
#pragma unroll
                for (int nb = 0; nb < DB; ++nb)
                    tma_load_2d(sV + u * Cfg::KV_BYTES + nb * 8192, &tmV, &bar_v[u], 64 * nb, kv_row0 + j * BKV);
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
                        const uint64_t offk = (uint64_t)((db * 8192 + 32 * k) >> 4);
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
                mbar_wait(bar_p, (uint32_t)j & 1u);                   // P_j stored (and O rescaled where needed)
                mbar_wait(&bar_v[u], (uint32_t)(j >> 1) & 1u);
                tcgen05_fence_after();
                const uint64_t dp = make_smem_desc_sw128(smem_u32(sP), 16, 1024);
                const uint64_t dv = make_smem_desc_sw128(smem_u32(sV + u * Cfg::KV_BYTES), 8192, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + Cfg::COL_O, dp + (uint64_t)((32 * k) >> 4), dv + (uint64_t)((2048 * k) >> 4), idesc_o,
                              (j | k) ? 1u : 0u);                     // the tensor core accumulates the tiles
                umma_commit(bar_o);
                if (j + 2 < n_tiles) {             // ring slot u is free once P_j V_j has been computed
                    mbar_wait(bar_o, (uint32_t)j & 1u);
                    load_kv(j + 2);
                }
            }
        }
    } else {
        // ===================================================================== softmax warps: thread = query row
        const int row = tid;                                          // == TMEM lane
        const int qpos = past_len + q0 + row;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float m_ref = -INFINITY, l_run = 0.f;                         // exponent reference (lags the true maximum), row sum
        unsigned char* prow = sP + (row >> 3) * 1024 + (row & 7) * 128;
        const int sw = row & 7;

        for (int j = 0; j < n_tiles; ++j) {
            const int u = j & 1;
            const uint32_t ts = tmem_base + lane_base + (u ? Cfg::COL_S1 : Cfg::COL_S0);
            const int k0 = j * BKV;
            const bool need_mask = (k0 + BKV - 1 > past_len + q0) || (k0 + BKV > T);   // warp-uniform
            mbar_wait(&bar_s[u], (uint32_t)(j >> 1) & 1u);
            tcgen05_fence_after();
            uint32_t r[64];
            tmem_ld32(ts, r);
            tmem_ld32(ts + 32, r + 32);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(&bar_sfree[u]);                               // S[u] may be overwritten by S_{j+2}
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                const int kp = k0 + i;
                if (need_mask && !(kp <= qpos && kp < T)) r[i] = __float_as_uint(-INFINITY);
                mx = fmaxf(mx, __uint_as_float(r[i]));
            }
            const float m_new = fmaxf(m_ref, mx * scale_log2);
            // lazy reference update: first finite maximum, or growth beyond the threshold
            const bool grow = m_new > m_ref + T2_RESCALE || (m_ref == -INFINITY && m_new != -INFINITY);
            const float factor = (grow && m_ref != -INFINITY) ? t2_exp2(m_ref - m_new) : 1.0f;
            if (grow) {
                m_ref = m_new;
                l_run *= factor;
            }
            // everything issued so far has accumulated into O once P_{j-1} V_{j-1} is done (it ran under this softmax)
            if (j > 0) {
                mbar_wait(bar_o, (uint32_t)(j - 1) & 1u);
                if (__any_sync(0xffffffffu, factor != 1.0f)) {        // rare: rescale these 32 rows of O in place
                    tcgen05_fence_after();
#pragma unroll
                    for (int c = 0; c < D / 32; ++c) {
                        uint32_t o[32];
                        tmem_ld32(tmem_base + lane_base + Cfg::COL_O + c * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
                        tmem_st32(tmem_base + lane_base + Cfg::COL_O + c * 32, o);
                    }
                    tmem_st_wait();
                    tcgen05_fence_before();
                }
            }
            const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;  // fully masked so far (rows beyond S): p = 0
            float rs = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {                             // eight 16-byte chunks (8 keys each)
                float p[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    p[i] = t2_exp2(__uint_as_float(r[8 * q + i]) * scale_log2 - m_use);   // exp2(-inf) = 0 for masked keys
                    rs += p[i];
                }
                *reinterpret_cast<uint4*>(prow + ((q ^ sw) << 4)) =
                    make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
            }
            l_run += rs;
            fence_proxy_async();                                      // P stores -> visible to the tensor core's reads
            mbar_arrive(bar_p);
        }
        // ---- O / l -> global
        mbar_wait(bar_o, (uint32_t)(n_tiles - 1) & 1u);
        tcgen05_fence_after();
        const bool live = q0 + row < S;
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
        bf16* dst = out + ((size_t)b * S + q0 + row) * n_h * D + (size_t)h * D;
#pragma unroll
        for (int c = 0; c < D / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_base + lane_base + Cfg::COL_O + c * 32, o);
            tmem_ld_wait();
            if (live) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) =
                        make_uint4(pack_bf16(__uint_as_float(o[8 * q]) * inv, __uint_as_float(o[8 * q + 1]) * inv),
                                   pack_bf16(__uint_as_float(o[8 * q + 2]) * inv, __uint_as_float(o[8 * q + 3]) * inv),
                                   pack_bf16(__uint_as_float(o[8 * q + 4]) * inv, __uint_as_float(o[8 * q + 5]) * inv),
                                   pack_bf16(__uint_as_float(o[8 * q + 6]) * inv, __uint_as_float(o[8 * q + 7]) * inv));
            }
        }
        if (live && lse) lse[((size_t)b * n_h + h) * S + q0 + row] = m_ref * 0.6931471805599453f + logf(l_run);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) {
        tcgen05_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int D>
static int launch_tc2(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, int B, int S, int past_len,
                      int n_h, int n_kv, int T_max, float scale, cudaStream_t st) {
    using Cfg = T2Cfg<D>;
    CUtensorMap tmQ, tmK, tmV;
    int rc = make_tensor_map(&tmQ, q, (uint64_t)n_h * D, (uint64_t)B * S, (uint64_t)n_h * D, 64, 128);
    if (rc != TL_OK) return rc;
    rc = make_tensor_map(&tmK, k_cache, (uint64_t)D, (uint64_t)B * n_kv * T_max, (uint64_t)D, 64, 64);
    if (rc != TL_OK) return rc;
    rc = make_tensor_map(&tmV, v_cache, (uint64_t)D, (uint64_t)B * n_kv * T_max, (uint64_t)D, 64, 64);
    if (rc != TL_OK) return rc;
    auto kern = attn_prefill_tc2_kernel<D>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess)
            return check_launch("tl_attn_prefill_fwd/tc2 (smem attr)");
        attr_done = true;
    }
    const dim3 grid((S + T2_BQ - 1) / T2_BQ, n_h, B);
    kern<<<grid, T2_THREADS, Cfg::SMEM_BYTES, st>>>(tmQ, tmK, tmV, (bf16*)out, lse, S, past_len, n_h, n_kv, T_max,
                                                   scale * 1.4426950408889634f);
    return check_launch("tl_attn_prefill_fwd/tc2");
}

// returns TL_OK / error, or 1 = not applicable
int attn_prefill_tc2_dispatch(const void* q, const void* k_cache, const void* v_cache, void* out, float* lse, int B, int S,
                              int past_len, int n_h, int n_kv, int d, int T_max, float scale, cudaStream_t st) {
    if ((((uintptr_t)q | (uintptr_t)k_cache | (uintptr_t)v_cache | (uintptr_t)out) & 15) != 0) return 1;
    if (d == 128) return launch_tc2<128>(q, k_cache, v_cache, out, lse, B, S, past_len, n_h, n_kv, T_max, scale, st);
    if (d == 64) return launch_tc2<64>(q, k_cache, v_cache, out, lse, B, S, past_len, n_h, n_kv, T_max, scale, st);
    return 1;
}

}  // namespace tl
