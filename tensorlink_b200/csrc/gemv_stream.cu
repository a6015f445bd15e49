// Decode-shaped Linear, v2: persistent weight-streaming GEMV with a bulk-copy (TMA) producer warp.
//
// One CTA per SM.  Each CTA owns a contiguous range of row PAIRS of W (pair = gate/up for the SwiGLU epilogue)
// sized so every SM streams the same number of bytes (+-1 pair).  Warp 8 is the producer: it walks the CTA's
// rows and issues cp.async.bulk copies of <= 16 KB "stages" into a deep shared-memory ring (mbarrier full/empty,
// ~190 KB in flight per SM), never waiting on the math.  A stage is either P whole consecutive pairs (K small:
// rows are contiguous in memory, ONE copy) or one K-chunk of one pair (K large: two row-segment copies).  Stages
// are issued round-robin over the 8 consumer warps; warp w owns every 8th unit for all of its K chunks, so a dot
// product finishes with one warp reduction and no cross-warp traffic.  x (optionally RMS-normalised with HF
// rounding) is staged once per CTA while the producer is already streaming.
// Algorithmic bytes per launch = 2*N*K.
#include <stdlib.h>

#include "common.cuh"

namespace tl {

constexpr int GS_CONSUMER_WARPS = 8;
constexpr int GS_THREADS = (GS_CONSUMER_WARPS + 1) * 32;
constexpr int GS_MAX_STAGES = 16;
constexpr int GS_STAGE_BYTES = 16 * 1024;
constexpr int GS_KC = 4096;            // K chunk (elements) when a pair does not fit one stage

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int M>
__global__ void __launch_bounds__(GS_THREADS, 1)
gemv_stream_kernel(const bf16* __restrict__ x, const bf16* __restrict__ W, bf16* __restrict__ y, int N, int K,
                   const bf16* __restrict__ bias, const bf16* __restrict__ residual, const bf16* __restrict__ norm_w,
                   float eps, int flags, int P, int n_stages, int NW, int stage_bytes, const unsigned char* __restrict__ pf_ptr,
                   unsigned long long pf_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;                                                    // [n_stages][stage_bytes]
    bf16* xs = reinterpret_cast<bf16*>(smem + (size_t)n_stages * stage_bytes);     // [M][K]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_stages * stage_bytes + (((size_t)M * K * 2 + 15) & ~(size_t)15));
    uint64_t* empty_bar = full_bar + GS_MAX_STAGES;
    __shared__ float s_part[GS_CONSUMER_WARPS][M];

    // programmatic dependent launch: let the next kernel's CTAs start (and prefetch ITS weights) as SMs free up
    asm volatile("griddepcontrol.launch_dependents;");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npairs = N >> 1;
    const int p_begin = (int)((long long)blockIdx.x * npairs / gridDim.x);
    const int p_end = (int)((long long)(blockIdx.x + 1) * npairs / gridDim.x);
    const int n_units = (p_end - p_begin + P - 1) / P;                 // unit = P consecutive pairs
    // NW consumer warps take units; n_stages %% NW == 0, so ring slot s is ALWAYS consumed by warp s %% NW and every
    // waiter observes every phase of the barriers it waits on (no mbarrier parity aliasing).
    const int n_groups = (n_units + NW - 1) / NW;
    const bool chunked = K > GS_KC || (size_t)K * 4 > GS_STAGE_BYTES;  // a pair does not fit one stage
    const int KC = chunked ? GS_KC : K;
    const int n_chunks = (K + KC - 1) / KC;

    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == GS_CONSUMER_WARPS) {
        // ================================================================= producer (one elected lane)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int g = 0; g < n_groups; ++g) {
                for (int c = 0; c < n_chunks; ++c) {
                    for (int w = 0; w < NW; ++w) {
                        const int unit = g * NW + w;
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        unsigned char* dst = ring + (size_t)stage * stage_bytes;
                        if (unit >= n_units) {
                            mbar_expect_tx(&full_bar[stage], 0);
                        } else {
                            const int pair0 = p_begin + unit * P;
                            const int np = min(P, p_end - pair0);
                            if (!chunked) {      // np pairs = 2*np whole rows, contiguous in memory: one copy
                                const uint32_t bytes = (uint32_t)(2 * np) * (uint32_t)K * 2u;
                                mbar_expect_tx(&full_bar[stage], bytes);
                                bulk_load_1d(dst, W + (size_t)(2 * pair0) * K, bytes, &full_bar[stage]);
                            } else {             // one K chunk of the two rows of one pair: two copies
                                const int k0 = c * KC;
                                const uint32_t bytes = (uint32_t)min(KC, K - k0) * 2u;
                                mbar_expect_tx(&full_bar[stage], 2 * bytes);
                                bulk_load_1d(dst, W + (size_t)(2 * pair0) * K + k0, bytes, &full_bar[stage]);
                                bulk_load_1d(dst + (size_t)KC * 2, W + (size_t)(2 * pair0 + 1) * K + k0, bytes, &full_bar[stage]);
                            }
                        }
                        if (++stage == n_stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
            // every load of this CTA is issued (the last ring-full is still in flight): queue L2 prefetches of this
            // CTA's slice of the NEXT kernel's weights behind them, so HBM keeps streaming through this kernel's tail,
            // the launch boundary and the next kernel's prologue; the next kernel then fills its ring from L2.
            if (pf_bytes) {
                const unsigned long long per = ((pf_bytes / gridDim.x) + 4095ull) & ~4095ull;
                unsigned long long off = (unsigned long long)blockIdx.x * per;
                const unsigned long long end = off + per < pf_bytes ? off + per : pf_bytes;
                for (; off < end; off += 16384ull) {
                    const uint32_t sz = (uint32_t)(end - off < 16384ull ? end - off : 16384ull);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pf_ptr + off), "r"(sz) : "memory");
                }
            }
        }
    } else {
        // ================================================================= consumers
        // x / residual are produced by the previous kernel: wait for it (no-op without the PDL launch attribute);
        // the producer warp above streams weights, which nobody writes, without waiting.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        const int nvec = K >> 3;
        if (norm_w) {
            float ss[M];
#pragma unroll
            for (int m = 0; m < M; ++m) ss[m] = 0.f;
            for (int v = tid; v < nvec; v += GS_CONSUMER_WARPS * 32) {
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const uint4 u = reinterpret_cast<const uint4*>(x + (size_t)m * K)[v];
                    const uint32_t* u32 = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float a = bf16_lo(u32[j]), b = bf16_hi(u32[j]);
                        ss[m] += a * a + b * b;
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float t = warp_sum(ss[m]);
                if (lane == 0) s_part[warp][m] = t;
            }
            named_bar_sync(1, GS_CONSUMER_WARPS * 32);
            float rstd[M];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < GS_CONSUMER_WARPS; ++w) t += s_part[w][m];
                rstd[m] = 1.0f / sqrtf(t / (float)K + eps);
            }
            for (int v = tid; v < nvec; v += GS_CONSUMER_WARPS * 32) {
                const uint4 g = reinterpret_cast<const uint4*>(norm_w)[v];
                const uint32_t* g32 = reinterpret_cast<const uint32_t*>(&g);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const uint4 u = reinterpret_cast<const uint4*>(x + (size_t)m * K)[v];
                    uint4 o;
                    const uint32_t* u32 = reinterpret_cast<const uint32_t*>(&u);
                    uint32_t* o32 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        o32[j] = pack_bf16(bf16_lo(g32[j]) * rbf(bf16_lo(u32[j]) * rstd[m]),
                                           bf16_hi(g32[j]) * rbf(bf16_hi(u32[j]) * rstd[m]));
                    reinterpret_cast<uint4*>(xs + (size_t)m * K)[v] = o;
                }
            }
        } else {
            for (int v = tid; v < nvec * M; v += GS_CONSUMER_WARPS * 32)
                reinterpret_cast<uint4*>(xs)[v] = reinterpret_cast<const uint4*>(x)[v];
        }
        named_bar_sync(1, GS_CONSUMER_WARPS * 32);

        const bool swiglu = flags & TL_EPI_SWIGLU;
        const int n_out = swiglu ? npairs : N;
        // epilogue for one finished pair (all lanes hold the reduced sums)
        auto finish = [&](int pair, const float (&a0)[M], const float (&a1)[M]) {
            if (lane != 0) return;
            const int r0 = 2 * pair;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float v0 = a0[m], v1 = a1[m];
                if (flags & TL_EPI_BIAS) {
                    v0 += bf2f(bias[r0]);
                    v1 += bf2f(bias[r0 + 1]);
                }
                if (swiglu) {
                    const float gate = rbf(v0), up = rbf(v1);
                    y[(size_t)m * n_out + pair] = f2bf(rbf(silu_f(gate)) * up);
                } else {
                    float t0 = rbf(v0), t1 = rbf(v1);
                    if (flags & TL_EPI_RESIDUAL) {
                        t0 += bf2f(residual[(size_t)m * N + r0]);
                        t1 += bf2f(residual[(size_t)m * N + r0 + 1]);
                    }
                    *reinterpret_cast<uint32_t*>(y + (size_t)m * N + r0) = pack_bf16(t0, t1);
                }
            }
        };
        // dot product of `vecs` 16-byte vectors of two rows against x[k0..]
        auto dot2 = [&](const uint4* r0, const uint4* r1, int k0, int vecs, float (&a0)[M], float (&a1)[M]) {
#pragma unroll 4
            for (int v = lane; v < vecs; v += 32) {
                const uint4 w0 = r0[v], w1 = r1[v];
                const uint32_t* a32 = reinterpret_cast<const uint32_t*>(&w0);
                const uint32_t* b32 = reinterpret_cast<const uint32_t*>(&w1);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const uint4 xv = reinterpret_cast<const uint4*>(xs + (size_t)m * K + k0)[v];
                    const uint32_t* x32 = reinterpret_cast<const uint32_t*>(&xv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float xl = bf16_lo(x32[j]), xh = bf16_hi(x32[j]);
                        a0[m] = fmaf(bf16_lo(a32[j]), xl, a0[m]);
                        a0[m] = fmaf(bf16_hi(a32[j]), xh, a0[m]);
                        a1[m] = fmaf(bf16_lo(b32[j]), xl, a1[m]);
                        a1[m] = fmaf(bf16_hi(b32[j]), xh, a1[m]);
                    }
                }
            }
        };
        // this warp's stages are sequence numbers warp, warp+8, warp+16, ... of the producer's order
        int seq = warp;
        for (int g = 0; g < (warp < NW ? n_groups : 0); ++g) {
            const int unit = g * NW + warp;
            const bool valid = unit < n_units;
            const int pair0 = p_begin + unit * P;
            float a0[M], a1[M];
#pragma unroll
            for (int m = 0; m < M; ++m) a0[m] = a1[m] = 0.f;
            for (int c = 0; c < n_chunks; ++c, seq += NW) {
                const int stage = seq % n_stages;
                const uint32_t phase = (uint32_t)(seq / n_stages) & 1u;
                mbar_wait(&full_bar[stage], phase);
                const unsigned char* src = ring + (size_t)stage * stage_bytes;
                if (valid) {
                    if (!chunked) {
                        const int np = min(P, p_end - pair0);
                        for (int pp = 0; pp < np; ++pp) {
                            float b0[M], b1[M];
#pragma unroll
                            for (int m = 0; m < M; ++m) b0[m] = b1[m] = 0.f;
                            dot2(reinterpret_cast<const uint4*>(src + (size_t)(2 * pp) * K * 2),
                                 reinterpret_cast<const uint4*>(src + (size_t)(2 * pp + 1) * K * 2), 0, nvec, b0, b1);
#pragma unroll
                            for (int m = 0; m < M; ++m) { b0[m] = warp_sum(b0[m]); b1[m] = warp_sum(b1[m]); }
                            finish(pair0 + pp, b0, b1);
                        }
                    } else {
                        const int k0 = c * KC;
                        dot2(reinterpret_cast<const uint4*>(src), reinterpret_cast<const uint4*>(src + (size_t)KC * 2), k0,
                             min(KC, K - k0) >> 3, a0, a1);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[stage]);
            }
            if (valid && chunked) {
#pragma unroll
                for (int m = 0; m < M; ++m) { a0[m] = warp_sum(a0[m]); a1[m] = warp_sum(a1[m]); }
                finish(pair0, a0, a1);
            }
        }
    }
}

template <int M>
static int launch_stream(const void* x, const void* W, void* y, int N, int K, const void* bias, const void* residual,
                         const void* norm_w, float eps, int flags, const void* pf_ptr, size_t pf_bytes, cudaStream_t st) {
    auto kern = gemv_stream_kernel<M>;
    // Two half-size rings per SM (16 consumer warps, finer work split, the next kernel's CTAs become resident as soon as
    // one of the two exits) when an SM's share of W is small — the latency-bound regime of small models: Qwen2.5-0.5B
    // decode 1210 -> 1343 tok/s (round 2); one deep ring per SM otherwise (7B: 355.1 vs 354.9 tok/s).
    // TL_GEMV_CTAS_PER_SM=1|2 forces either.
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("TL_GEMV_CTAS_PER_SM");
        forced = (e && e[0] == '2') ? 2 : ((e && e[0] == '1') ? 1 : 0);
    }
    const int per_sm = forced ? forced : (((size_t)N * K * 2 / (size_t)sm_count() <= (size_t)128 * 1024) ? 2 : 1);
    // TL_GEMV_RING_KB (default 220): shared memory per CTA.  <= 110 leaves room for the NEXT kernel's CTA on the same
    // SM, so under programmatic dependent launch its producer fills its ring while this kernel is still streaming.
    static int ring_kb = 0;
    if (ring_kb == 0) {
        const char* e = getenv("TL_GEMV_RING_KB");
        ring_kb = e ? atoi(e) : 220;
        if (ring_kb < 48 || ring_kb > 220) ring_kb = 220;
    }
    const int SMEM_CAP = per_sm == 2 ? 110 * 1024 : ring_kb * 1024;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess)
            return check_launch("tl_gemv_bf16/stream (smem attr)");
        attr_done = true;
    }
    const size_t xs_bytes = (((size_t)M * K * 2) + 15) & ~(size_t)15;
    const size_t fixed = xs_bytes + 2 * GS_MAX_STAGES * sizeof(uint64_t);
    const bool chunked = K > GS_KC || (size_t)K * 4 > GS_STAGE_BYTES;
    int P = chunked ? 1 : (int)(GS_STAGE_BYTES / ((size_t)K * 4));
    if (P < 1) P = 1;
    if (P > 8) P = 8;
    // a stage holds exactly one unit: P whole pairs, or one 4096-column chunk of one pair
    const int stage_bytes = chunked ? GS_STAGE_BYTES : (int)((((size_t)P * K * 4) + 127) & ~(size_t)127);
    int max_stages = (int)((SMEM_CAP - fixed) / stage_bytes);
    if (max_stages > GS_MAX_STAGES) max_stages = GS_MAX_STAGES;
    if (max_stages < 4) return 1;   // caller falls back to the register-streaming kernel
    // pick (n_stages, NW): n_stages a multiple of NW, as many bytes in flight as possible, then as many warps
    int n_stages = 0, NW = 0;
    for (int nw = GS_CONSUMER_WARPS; nw >= 4; --nw) {
        const int st_ = max_stages / nw * nw;
        if (st_ > n_stages) { n_stages = st_; NW = nw; }
    }
    const size_t smem = (size_t)n_stages * stage_bytes + fixed;
    const int npairs = N >> 1;
    int grid = sm_count() * per_sm;
    if (grid > npairs) grid = npairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static int use_pdl = -1;
    if (use_pdl < 0) {
        const char* e = getenv("TL_PDL");
        use_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, (const bf16*)x, (const bf16*)W, (bf16*)y, N, K, (const bf16*)bias, (const bf16*)residual,
                       (const bf16*)norm_w, eps, flags, P, n_stages, NW, stage_bytes, (const unsigned char*)pf_ptr,
                       (unsigned long long)(pf_bytes & ~(size_t)15));
    return check_launch("tl_gemv_bf16/stream");
}

// returns TL_OK, an error, or 1 = "not applicable, use the fallback kernel"
int gemv_stream_dispatch(const void* x, const void* W, void* y, int M, int N, int K, const void* bias,
                         const void* residual, const void* norm_w, float eps, int flags, const void* pf_ptr, size_t pf_bytes,
                         cudaStream_t st) {
    if (K % 8 != 0 || ((uintptr_t)W & 15)) return 1;
    if ((uintptr_t)pf_ptr & 15) pf_bytes = 0;
    switch (M) {
        case 1: return launch_stream<1>(x, W, y, N, K, bias, residual, norm_w, eps, flags, pf_ptr, pf_bytes, st);
        case 2: return launch_stream<2>(x, W, y, N, K, bias, residual, norm_w, eps, flags, pf_ptr, pf_bytes, st);
        case 3: return launch_stream<3>(x, W, y, N, K, bias, residual, norm_w, eps, flags, pf_ptr, pf_bytes, st);
        case 4: return launch_stream<4>(x, W, y, N, K, bias, residual, norm_w, eps, flags, pf_ptr, pf_bytes, st);
        default: return 1;
    }
}

}  // namespace tl
