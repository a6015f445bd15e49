// One decode step of a pipeline stage as ONE persistent kernel.
//
// The per-kernel decode path (gemv_stream.cu + attention.cu) pays a ramp-up and a tail per launch: the two small
// Linears of a layer (qkv 33 MB, o 26 MB at Qwen2.5-7B) run at ~2.3 TB/s for that reason.  Here one CTA per SM stays
// resident for the whole step and walks a job list in device memory:
//     [EMBED] { GEMV qkv(+norm,bias) | ATTN (RoPE + append + attention) | GEMV o(+res) | GEMV gate/up(+norm,SwiGLU) |
//     GEMV down(+res) } x layers [ GEMV lm_head(+norm) | ARGMAX ] [ADVANCE]
// The producer warp never stops: it streams the weight rows of job j+1 into the shared-memory ring while the
// consumer warps are still finishing job j and crossing the grid barrier that publishes job j's output vector, so the
// HBM pipe stays busy across job boundaries (ring = 176 KB/SM ~ 26 MB chip-wide ~ 4 us of streaming).
// Arithmetic, rounding points and per-lane accumulation order are those of gemv_stream_kernel / attn_decode_fused_kernel,
// so results are bit-identical to the per-kernel path (tests/test_decode_step_gpu.py).
#include "common.cuh"

namespace tl {

constexpr int DS_CONSUMER_WARPS = 8;
constexpr int DS_CTHREADS = DS_CONSUMER_WARPS * 32;
constexpr int DS_THREADS = DS_CTHREADS + 32;
constexpr int DS_STAGE_BYTES = 16 * 1024;
constexpr int DS_KC = 4096;
constexpr int DS_MAX_STAGES = 16;
constexpr int DS_ATTN_MAX_T = 2048;
constexpr int DS_MAX_M = 4;

__device__ __forceinline__ void ds_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldcg_v4(const void* p) {     // L2-coherent load of data written by other CTAs
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
    unsigned short v;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return __uint_as_float(((uint32_t)v) << 16);
}

struct GemvGeom {
    int npairs, p_begin, p_end, P, n_units, n_groups, KC, n_chunks;
    bool chunked;
};
__device__ __forceinline__ GemvGeom gemv_geom(int N, int K, int NW) {
    GemvGeom g;
    g.npairs = N >> 1;
    g.p_begin = (int)((long long)blockIdx.x * g.npairs / gridDim.x);
    g.p_end = (int)((long long)(blockIdx.x + 1) * g.npairs / gridDim.x);
    g.chunked = K > DS_KC || (size_t)K * 4 > DS_STAGE_BYTES;
    g.P = g.chunked ? 1 : min(8, (int)(DS_STAGE_BYTES / ((size_t)K * 4)));
    g.n_units = (g.p_end - g.p_begin + g.P - 1) / g.P;
    g.n_groups = (g.n_units + NW - 1) / NW;
    g.KC = g.chunked ? DS_KC : K;
    g.n_chunks = (K + g.KC - 1) / g.KC;
    return g;
}

template <int M>
__global__ void __launch_bounds__(DS_THREADS, 1)
decode_step_kernel(const tl_decode_job* __restrict__ jobs, int n_jobs, unsigned* __restrict__ sync_ws, int n_stages, int NW,
                   int xs_elems) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ring = smem;
    bf16* xs = reinterpret_cast<bf16*>(smem + (size_t)n_stages * DS_STAGE_BYTES);                 // [M][K_max]
    float* attn_s = reinterpret_cast<float*>(smem + (size_t)n_stages * DS_STAGE_BYTES + (size_t)xs_elems * 2);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(attn_s + (3 * 128 + DS_ATTN_MAX_T + 8 + 16 + 4 * 128));
    uint64_t* empty_bar = full_bar + DS_MAX_STAGES;
    __shared__ float s_part[DS_CONSUMER_WARPS][M];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == DS_CONSUMER_WARPS) {
        // ================================================================= producer: weights of every GEMV job, in order
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < n_jobs; ++j) {
                const tl_decode_job& jb = jobs[j];
                if (jb.type != TL_JOB_GEMV) continue;
                const GemvGeom g = gemv_geom(jb.N, jb.K, NW);
                const bf16* W = reinterpret_cast<const bf16*>(jb.W);
                const int K = jb.K;
                for (int gi = 0; gi < g.n_groups; ++gi)
                    for (int c = 0; c < g.n_chunks; ++c)
                        for (int w = 0; w < NW; ++w) {
                            const int unit = gi * NW + w;
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            unsigned char* dst = ring + (size_t)stage * DS_STAGE_BYTES;
                            if (unit >= g.n_units) {
                                mbar_expect_tx(&full_bar[stage], 0);
                            } else {
                                const int pair0 = g.p_begin + unit * g.P;
                                const int np = min(g.P, g.p_end - pair0);
                                if (!g.chunked) {
                                    const uint32_t bytes = (uint32_t)(2 * np) * (uint32_t)K * 2u;
                                    mbar_expect_tx(&full_bar[stage], bytes);
                                    bulk_load_1d(dst, W + (size_t)(2 * pair0) * K, bytes, &full_bar[stage]);
                                } else {
                                    const int k0 = c * g.KC;
                                    const uint32_t bytes = (uint32_t)min(g.KC, K - k0) * 2u;
                                    mbar_expect_tx(&full_bar[stage], 2 * bytes);
                                    bulk_load_1d(dst, W + (size_t)(2 * pair0) * K + k0, bytes, &full_bar[stage]);
                                    bulk_load_1d(dst + (size_t)g.KC * 2, W + (size_t)(2 * pair0 + 1) * K + k0, bytes, &full_bar[stage]);
                                }
                            }
                            if (++stage == n_stages) { stage = 0; phase ^= 1; }
                        }
            }
        }
        return;
    }

    // ===================================================================== consumers (256 threads)
    unsigned bar_target = 0;
    auto grid_barrier = [&]() {       // publishes everything written so far to every CTA
        bar_target += gridDim.x;
        __threadfence();
        ds_bar(1, DS_CTHREADS);
        if (tid == 0) {
            atomicAdd(&sync_ws[0], 1u);
            while (ld_acquire_u32(&sync_ws[0]) < bar_target) {}
        }
        ds_bar(1, DS_CTHREADS);
    };
    int seq = warp;     // this warp's next ring sequence number (advances by NW per stage, only for warp < NW)

    for (int j = 0; j < n_jobs; ++j) {
        const tl_decode_job jb = jobs[j];
        if (jb.type == TL_JOB_GEMV) {
            const int N = jb.N, K = jb.K, flags = jb.flags;
            const bf16* x = reinterpret_cast<const bf16*>(jb.x);
            bf16* y = reinterpret_cast<bf16*>(jb.y);
            const bf16* bias = reinterpret_cast<const bf16*>(jb.bias);
            const bf16* residual = reinterpret_cast<const bf16*>(jb.residual);
            const bf16* norm_w = reinterpret_cast<const bf16*>(jb.norm_w);
            const float eps = jb.eps;
            const int nvec = K >> 3;
            // ---- stage x (written by other CTAs in the previous job: L2-coherent loads)
            if (norm_w) {
                float ss[M];
#pragma unroll
                for (int m = 0; m < M; ++m) ss[m] = 0.f;
                for (int v = tid; v < nvec; v += DS_CTHREADS) {
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const uint4 u = ldcg_v4(x + (size_t)m * K + (size_t)v * 8);
                        const uint32_t* u32 = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float a = bf16_lo(u32[q]), b = bf16_hi(u32[q]);
                            ss[m] += a * a + b * b;
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const float t = warp_sum(ss[m]);
                    if (lane == 0) s_part[warp][m] = t;
                }
                ds_bar(1, DS_CTHREADS);
                float rstd[M];
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    float t = 0.f;
#pragma unroll
                    for (int w = 0; w < DS_CONSUMER_WARPS; ++w) t += s_part[w][m];
                    rstd[m] = 1.0f / sqrtf(t / (float)K + eps);
                }
                for (int v = tid; v < nvec; v += DS_CTHREADS) {
                    const uint4 g = reinterpret_cast<const uint4*>(norm_w)[v];
                    const uint32_t* g32 = reinterpret_cast<const uint32_t*>(&g);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        const uint4 u = ldcg_v4(x + (size_t)m * K + (size_t)v * 8);
                        uint4 o;
                        const uint32_t* u32 = reinterpret_cast<const uint32_t*>(&u);
                        uint32_t* o32 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            o32[q] = pack_bf16(bf16_lo(g32[q]) * rbf(bf16_lo(u32[q]) * rstd[m]),
                                               bf16_hi(g32[q]) * rbf(bf16_hi(u32[q]) * rstd[m]));
                        reinterpret_cast<uint4*>(xs + (size_t)m * K)[v] = o;
                    }
                }
            } else {
                for (int v = tid; v < nvec * M; v += DS_CTHREADS)
                    reinterpret_cast<uint4*>(xs)[v] = ldcg_v4(x + (size_t)v * 8);
            }
            ds_bar(1, DS_CTHREADS);

            const GemvGeom g = gemv_geom(N, K, NW);
            const bool swiglu = flags & TL_EPI_SWIGLU;
            const int n_out = swiglu ? g.npairs : N;
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
                            t0 += ldcg_bf16(residual + (size_t)m * N + r0);
                            t1 += ldcg_bf16(residual + (size_t)m * N + r0 + 1);
                        }
                        *reinterpret_cast<uint32_t*>(y + (size_t)m * N + r0) = pack_bf16(t0, t1);
                    }
                }
            };
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
                        for (int q = 0; q < 4; ++q) {
                            const float xl = bf16_lo(x32[q]), xh = bf16_hi(x32[q]);
                            a0[m] = fmaf(bf16_lo(a32[q]), xl, a0[m]);
                            a0[m] = fmaf(bf16_hi(a32[q]), xh, a0[m]);
                            a1[m] = fmaf(bf16_lo(b32[q]), xl, a1[m]);
                            a1[m] = fmaf(bf16_hi(b32[q]), xh, a1[m]);
                        }
                    }
                }
            };
            if (warp < NW) {
                for (int gi = 0; gi < g.n_groups; ++gi) {
                    const int unit = gi * NW + warp;
                    const bool valid = unit < g.n_units;
                    const int pair0 = g.p_begin + unit * g.P;
                    float a0[M], a1[M];
#pragma unroll
                    for (int m = 0; m < M; ++m) a0[m] = a1[m] = 0.f;
                    for (int c = 0; c < g.n_chunks; ++c, seq += NW) {
                        const int stage = seq % n_stages;
                        const uint32_t phase = (uint32_t)(seq / n_stages) & 1u;
                        mbar_wait(&full_bar[stage], phase);
                        const unsigned char* src = ring + (size_t)stage * DS_STAGE_BYTES;
                        if (valid) {
                            if (!g.chunked) {
                                const int np = min(g.P, g.p_end - pair0);
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
                                const int k0 = c * g.KC;
                                dot2(reinterpret_cast<const uint4*>(src), reinterpret_cast<const uint4*>(src + (size_t)g.KC * 2), k0,
                                     min(g.KC, K - k0) >> 3, a0, a1);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[stage]);
                    }
                    if (valid && g.chunked) {
#pragma unroll
                        for (int m = 0; m < M; ++m) { a0[m] = warp_sum(a0[m]); a1[m] = warp_sum(a1[m]); }
                        finish(pair0, a0, a1);
                    }
                }
            }
            grid_barrier();
        } else if (jb.type == TL_JOB_ATTN) {
            // ---- RoPE + append + single-pass attention; CTA c < n_h*M handles (head c % n_h, row c / n_h) with warps 0..3
            const int n_h = jb.n_h, n_kv = jb.n_kv, D = jb.d, T_max = jb.T_max;
            const int HALF = D >> 1, n_rep = n_h / n_kv;
            const int cta = blockIdx.x;
            const bool active = cta < n_h * M;
            if (active && tid < 128) {
                const int h = cta % n_h, b = cta / n_h, kvh = h / n_rep;
                const int pos = *reinterpret_cast<const int32_t*>(jb.pos_dev);
                const int heads = n_h + 2 * n_kv;
                float* sq = attn_s;
                float* sk = sq + 128;
                float* sv = sk + 128;
                float* sscore = sv + 128;                      // [DS_ATTN_MAX_T + 8]
                float* sred = sscore + DS_ATTN_MAX_T + 8;      // [16]
                float* so = sred + 16;                         // [4][128]
                const bf16* row = reinterpret_cast<const bf16*>(jb.x) + (size_t)b * heads * D;
                const bf16* cos_tab = reinterpret_cast<const bf16*>(jb.cos_tab);
                const bf16* sin_tab = reinterpret_cast<const bf16*>(jb.sin_tab);
                const bf16* qn = reinterpret_cast<const bf16*>(jb.q_norm_w);
                const bf16* kn = reinterpret_cast<const bf16*>(jb.k_norm_w);
                bf16* k_cache = reinterpret_cast<bf16*>(jb.k_cache);
                bf16* v_cache = reinterpret_cast<bf16*>(jb.v_cache);
                float q1 = 0.f, q2 = 0.f, k1 = 0.f, k2 = 0.f;
                if (tid < HALF) {
                    q1 = ldcg_bf16(row + (size_t)h * D + tid);
                    q2 = ldcg_bf16(row + (size_t)h * D + tid + HALF);
                    k1 = ldcg_bf16(row + (size_t)(n_h + kvh) * D + tid);
                    k2 = ldcg_bf16(row + (size_t)(n_h + kvh) * D + tid + HALF);
                    sv[tid] = ldcg_bf16(row + (size_t)(n_h + n_kv + kvh) * D + tid);
                    sv[tid + HALF] = ldcg_bf16(row + (size_t)(n_h + n_kv + kvh) * D + tid + HALF);
                }
                if (qn) {
                    float a = warp_sum(q1 * q1 + q2 * q2), c = warp_sum(k1 * k1 + k2 * k2);
                    if (lane == 0) { sred[warp] = a; sred[4 + warp] = c; }
                    ds_bar(2, 128);
                    a = sred[0] + sred[1] + sred[2] + sred[3];
                    c = sred[4] + sred[5] + sred[6] + sred[7];
                    ds_bar(2, 128);
                    const float rq = 1.0f / sqrtf(a / (float)D + jb.eps), rk = 1.0f / sqrtf(c / (float)D + jb.eps);
                    if (tid < HALF) {
                        q1 = rbf(bf2f(qn[tid]) * rbf(q1 * rq));
                        q2 = rbf(bf2f(qn[tid + HALF]) * rbf(q2 * rq));
                        k1 = rbf(bf2f(kn[tid]) * rbf(k1 * rk));
                        k2 = rbf(bf2f(kn[tid + HALF]) * rbf(k2 * rk));
                    }
                }
                if (tid < HALF) {
                    const float c = bf2f(cos_tab[(size_t)pos * HALF + tid]), s = bf2f(sin_tab[(size_t)pos * HALF + tid]);
                    sq[tid] = rbf(rbf(q1 * c) + rbf(-q2 * s));
                    sq[tid + HALF] = rbf(rbf(q2 * c) + rbf(q1 * s));
                    sk[tid] = rbf(rbf(k1 * c) + rbf(-k2 * s));
                    sk[tid + HALF] = rbf(rbf(k2 * c) + rbf(k1 * s));
                }
                ds_bar(2, 128);
                if (h % n_rep == 0 && tid < D) {
                    const size_t off = (((size_t)b * n_kv + kvh) * T_max + pos) * D + tid;
                    k_cache[off] = f2bf(sk[tid]);
                    v_cache[off] = f2bf(sv[tid]);
                }
                const bf16* kb = k_cache + ((size_t)b * n_kv + kvh) * T_max * D;
                const bf16* vb = v_cache + ((size_t)b * n_kv + kvh) * T_max * D;
                const float scale_log2 = jb.scale * 1.4426950408889634f;
                float mx = -INFINITY;
                for (int key = tid; key < pos; key += 128) {
                    const uint4* kr = reinterpret_cast<const uint4*>(kb + (size_t)key * D);
                    float acc = 0.f;
                    for (int c = 0; c < (D >> 3); ++c) {
                        const uint4 kv4 = kr[c];
                        const uint32_t* k32 = reinterpret_cast<const uint32_t*>(&kv4);
                        const float4 qa = *reinterpret_cast<const float4*>(&sq[c * 8]);
                        const float4 qb = *reinterpret_cast<const float4*>(&sq[c * 8 + 4]);
                        acc += bf16_lo(k32[0]) * qa.x + bf16_hi(k32[0]) * qa.y + bf16_lo(k32[1]) * qa.z + bf16_hi(k32[1]) * qa.w +
                               bf16_lo(k32[2]) * qb.x + bf16_hi(k32[2]) * qb.y + bf16_lo(k32[3]) * qb.z + bf16_hi(k32[3]) * qb.w;
                    }
                    acc *= scale_log2;
                    sscore[key] = acc;
                    mx = fmaxf(mx, acc);
                }
                if (warp == 0) {
                    float acc = 0.f;
                    for (int i = lane; i < D; i += 32) acc += sq[i] * sk[i];
                    acc = warp_sum(acc) * scale_log2;
                    if (lane == 0) sscore[pos] = acc;
                    mx = fmaxf(mx, acc);
                }
                mx = warp_max(mx);
                if (lane == 0) sred[warp] = mx;
                ds_bar(2, 128);
                const float m_all = fmaxf(fmaxf(sred[0], sred[1]), fmaxf(sred[2], sred[3]));
                ds_bar(2, 128);
                float lsum = 0.f;
                for (int key = tid; key <= pos; key += 128) {
                    const float p = exp2f(sscore[key] - m_all);
                    lsum += p;
                    sscore[key] = rbf(p);
                }
                lsum = warp_sum(lsum);
                if (lane == 0) sred[8 + warp] = lsum;
                ds_bar(2, 128);
                const int TPR = D >> 3, KG = 128 / TPR;
                const int dd0 = (tid % TPR) * 8, kgi = tid / TPR;
                float acc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                for (int key = kgi; key < pos; key += KG) {
                    const uint4 vv = *reinterpret_cast<const uint4*>(vb + (size_t)key * D + dd0);
                    const uint32_t* v32 = reinterpret_cast<const uint32_t*>(&vv);
                    const float p = sscore[key];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[2 * q] = fmaf(p, bf16_lo(v32[q]), acc[2 * q]);
                        acc[2 * q + 1] = fmaf(p, bf16_hi(v32[q]), acc[2 * q + 1]);
                    }
                }
                if (kgi == 0) {
                    const float p = sscore[pos];
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[q] = fmaf(p, sv[dd0 + q], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float v = acc[q];
                    for (int off = TPR; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    acc[q] = v;
                }
                if (lane < TPR) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) so[warp * 128 + dd0 + q] = acc[q];
                }
                ds_bar(2, 128);
                if (tid < D) {
                    const float l = sred[8] + sred[9] + sred[10] + sred[11];
                    const float o = so[tid] + so[128 + tid] + so[256 + tid] + so[384 + tid];
                    reinterpret_cast<bf16*>(jb.y)[((size_t)b * n_h + h) * D + tid] = f2bf(o / l);
                }
            }
            grid_barrier();
        } else if (jb.type == TL_JOB_EMBED) {
            // y[m, :] = table[ids[m], :]  (K = hidden).  Rows are split over the first M CTAs.
            if (blockIdx.x < M) {
                const int m = blockIdx.x;
                long long id = reinterpret_cast<const int64_t*>(jb.x)[m];
                if (id < 0 || id >= jb.N) id = 0;
                const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(jb.W) + (size_t)id * jb.K);
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(jb.y) + (size_t)m * jb.K);
                for (int i = tid; i < (jb.K >> 3); i += DS_CTHREADS) dst[i] = src[i];
            }
            grid_barrier();
        } else if (jb.type == TL_JOB_ARGMAX) {
            // ids[m] = argmax_v logits[m, v] (lowest index on ties); logits = jb.x [M, N] bf16 written by the lm_head job
            const int V = jb.N;
            float* pval = reinterpret_cast<float*>(const_cast<void*>(jb.W));   // [M][grid] partial maxima (workspace)
            int* pidx = reinterpret_cast<int*>(pval + (size_t)M * gridDim.x);
            const int per = (V + gridDim.x - 1) / gridDim.x;
            const int lo = blockIdx.x * per, hi = min(V, lo + per);
            for (int m = 0; m < M; ++m) {
                const bf16* rowp = reinterpret_cast<const bf16*>(jb.x) + (size_t)m * V;
                float best = -INFINITY;
                int bi = 0x7fffffff;
                for (int i = lo + tid; i < hi; i += DS_CTHREADS) {
                    const float v = ldcg_bf16(rowp + i);
                    if (v > best) { best = v; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                float* sv_ = attn_s;
                int* si_ = reinterpret_cast<int*>(attn_s + 16);
                if (lane == 0) { sv_[warp] = best; si_[warp] = bi; }
                ds_bar(1, DS_CTHREADS);
                if (tid == 0) {
                    for (int w = 1; w < DS_CONSUMER_WARPS; ++w)
                        if (sv_[w] > best || (sv_[w] == best && si_[w] < bi)) { best = sv_[w]; bi = si_[w]; }
                    pval[(size_t)m * gridDim.x + blockIdx.x] = best;
                    pidx[(size_t)m * gridDim.x + blockIdx.x] = bi;
                }
                ds_bar(1, DS_CTHREADS);
            }
            grid_barrier();
            if (blockIdx.x == 0 && warp < M) {
                const int m = warp;
                float best = -INFINITY;
                int bi = 0x7fffffff;
                for (int i = lane; i < (int)gridDim.x; i += 32) {
                    const float v = __ldcg(&pval[(size_t)m * gridDim.x + i]);
                    const int ix = __ldcg(&pidx[(size_t)m * gridDim.x + i]);
                    if (v > best || (v == best && ix < bi)) { best = v; bi = ix; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (lane == 0) reinterpret_cast<int64_t*>(jb.y)[m] = (bi == 0x7fffffff) ? 0 : (int64_t)bi;
            }
        } else if (jb.type == TL_JOB_ADVANCE) {
            // pos += 1, kv_len = pos (after every CTA has finished reading pos: this job follows a grid barrier)
            if (blockIdx.x == 0 && tid == 0) {
                int32_t* pos = reinterpret_cast<int32_t*>(const_cast<void*>(jb.pos_dev));
                const int p = *pos + 1;
                *pos = p;
                if (jb.y) *reinterpret_cast<int32_t*>(jb.y) = p;
            }
        }
    }
    // ---- self-cleaning: the last CTA to leave resets the barrier counter for the next launch
    ds_bar(1, DS_CTHREADS);
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(&sync_ws[1], 1u);
        if (prev == gridDim.x - 1) {
            sync_ws[0] = 0;
            sync_ws[1] = 0;
            __threadfence();
        }
    }
}

}  // namespace tl

extern "C" {

size_t tl_decode_step_ws(int M) { return 64 + (size_t)M * 1024 * (sizeof(float) + sizeof(int)); }

int tl_decode_step(const tl_decode_job* jobs_dev, const tl_decode_job* jobs_host, int n_jobs, int M, void* sync_ws,
                   void* stream) {
    using namespace tl;
    TL_REQUIRE(M >= 1 && M <= DS_MAX_M, TL_ERR_INVALID, "tl_decode_step: M=%d outside 1..%d", M, DS_MAX_M);
    TL_REQUIRE(jobs_dev && jobs_host && n_jobs > 0 && sync_ws, TL_ERR_INVALID, "tl_decode_step: null argument");
    int k_max = 0;
    for (int j = 0; j < n_jobs; ++j) {
        const tl_decode_job& jb = jobs_host[j];
        if (jb.type == TL_JOB_GEMV) {
            TL_REQUIRE(jb.K % 8 == 0 && jb.N % 2 == 0 && jb.N > 0, TL_ERR_INVALID, "tl_decode_step: job %d bad GEMV shape", j);
            if (jb.K > k_max) k_max = jb.K;
        } else if (jb.type == TL_JOB_ATTN) {
            TL_REQUIRE((jb.d == 64 || jb.d == 128) && jb.T_max <= DS_ATTN_MAX_T && jb.n_h * M <= sm_count(), TL_ERR_INVALID,
                       "tl_decode_step: job %d attention shape unsupported (d=%d T_max=%d)", j, jb.d, jb.T_max);
        }
    }
    constexpr int SMEM_CAP = 226 * 1024;
    const size_t xs_bytes = (((size_t)M * k_max * 2) + 127) & ~(size_t)127;
    const size_t attn_bytes = (size_t)(3 * 128 + DS_ATTN_MAX_T + 8 + 16 + 4 * 128) * sizeof(float);
    const size_t fixed = xs_bytes + attn_bytes + 2 * DS_MAX_STAGES * sizeof(uint64_t);
    TL_REQUIRE(fixed + 4 * DS_STAGE_BYTES <= (size_t)SMEM_CAP, TL_ERR_INVALID, "tl_decode_step: M*K_max too large (%zu B)", fixed);
    int max_stages = (int)((SMEM_CAP - fixed) / DS_STAGE_BYTES);
    if (max_stages > DS_MAX_STAGES) max_stages = DS_MAX_STAGES;
    int n_stages = 0, NW = 0;
    for (int nw = DS_CONSUMER_WARPS; nw >= 4; --nw) {
        const int s = max_stages / nw * nw;
        if (s > n_stages) { n_stages = s; NW = nw; }
    }
    const size_t smem = (size_t)n_stages * DS_STAGE_BYTES + fixed;
    const int grid = sm_count();
    cudaStream_t st = (cudaStream_t)stream;
    const int xs_elems = (int)(xs_bytes / 2);
#define TL_DS_LAUNCH(MM)                                                                                              \
    {                                                                                                                 \
        static bool done = false;                                                                                     \
        if (!done) {                                                                                                  \
            if (cudaFuncSetAttribute(decode_step_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CAP) != cudaSuccess) \
                return check_launch("tl_decode_step (smem attr)");                                                  \
            done = true;                                                                                              \
        }                                                                                                             \
        decode_step_kernel<MM><<<grid, DS_THREADS, smem, st>>>(jobs_dev, n_jobs, (unsigned*)sync_ws, n_stages, NW, xs_elems); \
    }
    switch (M) {
        case 1: TL_DS_LAUNCH(1) break;
        case 2: TL_DS_LAUNCH(2) break;
        case 3: TL_DS_LAUNCH(3) break;
        default: TL_DS_LAUNCH(4) break;
    }
#undef TL_DS_LAUNCH
    return check_launch("tl_decode_step");
}

}  // extern "C"
