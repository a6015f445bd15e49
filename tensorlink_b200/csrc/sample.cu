// Token sampling on the device: temperature -> top-k -> top-p -> multinomial draw, one CTA per row.
//
// The reference hands generation to HF ``generate`` (tensorlink/ml/module.py:763-769, ml/worker.py:403-404), whose
// sampling path is ``TemperatureLogitsWarper`` -> ``TopKLogitsWarper`` -> ``TopPLogitsWarper`` -> ``torch.multinomial``
// on the host's copy of the logits.  Here the logits never leave the last stage: this kernel reads the bf16 row the
// lm_head just wrote and stores one int64 id (into the first stage's mailbox on a multi-stage job).
//
// Everything is decided on a histogram over the 65536 possible bf16 bit patterns (integer atomics: deterministic):
//   * max                = highest occupied bin
//   * top-k threshold    = value of the k-th largest logit; every logit >= it is kept (HF: `scores < topk[-1]` removed)
//   * top-p threshold    = walking the bins downwards, a bin is kept while the probability mass ABOVE it is < top_p
//                          (HF removes ascending-cumulative-mass <= 1 - top_p and always keeps the top token); bf16
//                          logits tie often, and ties at the threshold are all kept
//   * draw               = Philox4x32-10(seed; row, counter) -> u in [0,1); inverse CDF over the kept tokens in index
//                          order (chunked prefix sums in a fixed order: a seed reproduces its tokens)
// The per-row counter lives in device memory and is advanced by the kernel, so a captured CUDA graph draws a fresh
// number on every replay.
#include "common.cuh"

namespace tl {

constexpr int SM_THREADS = 1024;
constexpr int SM_BINS = 65536;
constexpr int SM_PER = SM_BINS / SM_THREADS;      // bins per thread in the scans

__device__ __forceinline__ uint32_t bf16_key(uint16_t bits) {       // monotone: larger value -> larger key
    return (bits & 0x8000u) ? (uint32_t)(uint16_t)~bits : (uint32_t)(bits | 0x8000u);
}
__device__ __forceinline__ float key_value(uint32_t key) {
    const uint16_t bits = (key & 0x8000u) ? (uint16_t)(key & 0x7fffu) : (uint16_t)~key;
    return __uint_as_float(((uint32_t)bits) << 16);
}

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ float philox_uniform(unsigned long long seed, uint32_t row, uint32_t counter) {
    uint32_t c[4] = {counter, row, 0x5eed5eedu, 0u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return (float)(c[0] >> 8) * (1.0f / 16777216.0f);       // 24 random bits: [0, 1)
}

// exclusive block scan of one double per thread, in thread order (deterministic); returns the total through *total
__device__ double block_excl_scan(double v, double* s_warp, double* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        double w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const double before = (warp ? s_warp[warp - 1] : 0.0) + (x - v);
    *total = s_warp[31];
    __syncthreads();
    return before;
}

__global__ void __launch_bounds__(SM_THREADS) sample_kernel(const bf16* __restrict__ logits, int64_t* __restrict__ ids_out, int V,
                                                            float inv_temp, int top_k, float top_p, unsigned long long seed,
                                                            int32_t* __restrict__ counters, uint32_t* __restrict__ hist_all) {
    const int row = blockIdx.x, tid = threadIdx.x;
    const uint16_t* lr = reinterpret_cast<const uint16_t*>(logits) + (size_t)row * V;
    uint32_t* hist = hist_all + (size_t)row * SM_BINS;
    __shared__ double s_warp[32];
    __shared__ int s_sel[4];
    __shared__ double s_val[2];
    for (int i = tid; i < SM_BINS; i += SM_THREADS) hist[i] = 0u;
    __syncthreads();
    for (int i = tid; i < V; i += SM_THREADS) atomicAdd(&hist[bf16_key(lr[i])], 1u);
    __syncthreads();
    // this thread owns bins [hi_key - SM_PER + 1, hi_key], walked downwards: thread 0 holds the largest values
    const int hi_key = SM_BINS - 1 - tid * SM_PER;
    uint32_t cnt[SM_PER];
    uint32_t local_n = 0;
#pragma unroll 8
    for (int j = 0; j < SM_PER; ++j) {
        cnt[j] = hist[hi_key - j];
        local_n += cnt[j];
    }
    // ---- max: the first occupied bin from the top
    double tot;
    const double n_before = block_excl_scan((double)local_n, s_warp, &tot);
    if (n_before == 0.0 && local_n > 0) {
        for (int j = 0; j < SM_PER; ++j)
            if (cnt[j]) { s_sel[0] = hi_key - j; break; }
    }
    __syncthreads();
    const int max_key = s_sel[0];
    const float x_max = key_value((uint32_t)max_key) * inv_temp;
    // ---- top-k: the bin holding the k-th largest logit (every logit of that bin is kept)
    int k_key = 0;                                   // keep everything by default
    if (top_k > 0 && top_k < V) {
        if (n_before < (double)top_k && n_before + (double)local_n >= (double)top_k) {
            double c = n_before;
            for (int j = 0; j < SM_PER; ++j) {
                c += (double)cnt[j];
                if (c >= (double)top_k) { s_sel[1] = hi_key - j; break; }
            }
        }
        __syncthreads();
        k_key = s_sel[1];
    }
    // ---- probability mass per thread over the bins that survive top-k (fixed order -> deterministic)
    double local_m = 0.0;
    for (int j = 0; j < SM_PER; ++j) {
        const int key = hi_key - j;
        if (cnt[j] && key >= k_key) local_m += (double)cnt[j] * (double)__expf(key_value((uint32_t)key) * inv_temp - x_max);
    }
    double Z;
    const double m_before = block_excl_scan(local_m, s_warp, &Z);
    // ---- top-p: keep a bin while the mass above it is < top_p * Z; the lowest kept bin is the threshold
    int p_key = k_key;
    double Z_kept = Z;
    if (top_p < 1.0f) {
        const double lim = (double)top_p * Z;
        if (tid == 0) { s_sel[2] = k_key; s_val[0] = Z; }
        __syncthreads();
        // the thread whose range contains the crossing point: mass before its range < lim <= mass through its range
        if (m_before < lim && m_before + local_m >= lim) {
            double c = m_before;
            for (int j = 0; j < SM_PER; ++j) {
                const int key = hi_key - j;
                if (!(cnt[j] && key >= k_key)) continue;
                const double w = (double)cnt[j] * (double)__expf(key_value((uint32_t)key) * inv_temp - x_max);
                if (c < lim) { s_sel[2] = key; s_val[0] = c + w; }     // kept: mass above it is still below the limit
                c += w;
            }
        }
        __syncthreads();
        p_key = s_sel[2];
        Z_kept = s_val[0];
    }
    // ---- draw and invert the CDF over the kept tokens in index order
    const uint32_t ctr = (uint32_t)counters[row];
    const double target = (double)philox_uniform(seed, (uint32_t)row, ctr) * Z_kept;
    const int per = (V + SM_THREADS - 1) / SM_THREADS;
    const int i0 = tid * per, i1 = min(V, i0 + per);
    double local_w = 0.0;
    for (int i = i0; i < i1; ++i) {
        const uint32_t key = bf16_key(lr[i]);
        if ((int)key >= p_key) local_w += (double)__expf(key_value(key) * inv_temp - x_max);
    }
    double W;
    const double w_before = block_excl_scan(local_w, s_warp, &W);
    if (tid == 0) s_sel[3] = -1;
    __syncthreads();
    if (local_w > 0.0 && w_before <= target && target < w_before + local_w) {
        double c = w_before;
        int pick = -1;
        for (int i = i0; i < i1; ++i) {
            const uint32_t key = bf16_key(lr[i]);
            if ((int)key < p_key) continue;
            pick = i;
            c += (double)__expf(key_value(key) * inv_temp - x_max);
            if (target < c) break;
        }
        s_sel[3] = pick;
    }
    __syncthreads();
    if (tid == 0) {
        int pick = s_sel[3];
        if (pick < 0) {                 // rounding left the target at / beyond the total: the last kept token
            for (int i = V - 1; i >= 0; --i)
                if ((int)bf16_key(lr[i]) >= p_key) { pick = i; break; }
        }
        ids_out[row] = (int64_t)pick;
        counters[row] = (int32_t)(ctr + 1u);
    }
}

}  // namespace tl

extern "C" {

size_t tl_sample_ws(int M) { return (size_t)(M > 0 ? M : 0) * tl::SM_BINS * sizeof(uint32_t); }

int tl_sample(const void* logits, int64_t* ids_out, int M, int V, float temperature, int top_k, float top_p,
              unsigned long long seed, int32_t* counters_dev, void* workspace, size_t ws_bytes, void* stream) {
    using namespace tl;
    TL_REQUIRE(logits && ids_out && counters_dev && workspace, TL_ERR_INVALID, "tl_sample: null argument");
    TL_REQUIRE(M >= 1 && V >= 1, TL_ERR_INVALID, "tl_sample: bad shape M=%d V=%d", M, V);
    TL_REQUIRE(temperature > 0.f && top_p > 0.f && top_p <= 1.f && top_k >= 0, TL_ERR_INVALID,
               "tl_sample: temperature must be > 0, 0 < top_p <= 1, top_k >= 0 (got %g, %g, %d)", temperature, top_p, top_k);
    TL_REQUIRE(ws_bytes >= tl_sample_ws(M), TL_ERR_INVALID, "tl_sample: workspace too small");
    sample_kernel<<<M, SM_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)logits, ids_out, V, 1.0f / temperature, top_k, top_p, seed,
                                                              counters_dev, (uint32_t*)workspace);
    return check_launch("tl_sample");
}

}  // extern "C"
