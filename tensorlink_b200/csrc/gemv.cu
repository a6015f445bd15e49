// Decode-shaped Linear: y[M,N] = f(norm(x)[M,K] · W[N,K]^T), M <= 8.  HBM-bound weight streaming.
//
// Work item = G consecutive row PAIRS of W (a pair is gate/up for the SwiGLU epilogue).  WPI warps share
// an item and split K between them with a CTA-coalesced stride; 8/WPI items are in flight per CTA.
// Every lane keeps 2*G independent 16-byte loads in flight per k-iteration (ld.global.nc, L1 no-allocate).
// Algorithmic bytes per launch = 2*N*K (weights) + O(M*(K+N)) activations.
#include <stdlib.h>

#include "common.cuh"

namespace tl {

constexpr int GEMV_THREADS = 256;
constexpr int GEMV_WARPS = GEMV_THREADS / 32;

template <int M, int G, int WPI>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_kernel(const bf16* __restrict__ x, const bf16* __restrict__ W,
                                                            bf16* __restrict__ y, int N, int K,
                                                            const bf16* __restrict__ bias,
                                                            const bf16* __restrict__ residual,
                                                            const bf16* __restrict__ norm_w, float eps, int flags) {
    constexpr int SLOTS = GEMV_WARPS / WPI;
    constexpr int ROWS = 2 * G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* xs = reinterpret_cast<bf16*>(smem_raw);                                  // [M][K]
    float* red = reinterpret_cast<float*>(smem_raw + (size_t)M * K * sizeof(bf16));  // [WARPS][ROWS*M]
    __shared__ float s_part[GEMV_WARPS][M];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nvec = K >> 3;

    // ---- prologue: stage x (optionally RMS-normalised with HF rounding) in shared memory
    if (norm_w) {
        float ss[M];
#pragma unroll
        for (int m = 0; m < M; ++m) ss[m] = 0.f;
        for (int v = tid; v < nvec; v += GEMV_THREADS) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                uint4 u = reinterpret_cast<const uint4*>(x + (size_t)m * K)[v];
                const uint32_t* w32 = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = bf16_lo(w32[j]), b = bf16_hi(w32[j]);
                    ss[m] += a * a + b * b;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float t = warp_sum(ss[m]);
            if (lane == 0) s_part[warp][m] = t;
        }
        __syncthreads();
        float rstd[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < GEMV_WARPS; ++w) t += s_part[w][m];
            rstd[m] = 1.0f / sqrtf(t / (float)K + eps);
        }
        for (int v = tid; v < nvec; v += GEMV_THREADS) {
            uint4 g = reinterpret_cast<const uint4*>(norm_w)[v];
            const uint32_t* g32 = reinterpret_cast<const uint32_t*>(&g);
#pragma unroll
            for (int m = 0; m < M; ++m) {
                uint4 u = reinterpret_cast<const uint4*>(x + (size_t)m * K)[v], o;
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
        for (int v = tid; v < nvec * M; v += GEMV_THREADS)
            reinterpret_cast<uint4*>(xs)[v] = reinterpret_cast<const uint4*>(x)[v];
    }
    __syncthreads();

    const int npairs = N >> 1;
    const int n_items = (npairs + G - 1) / G;
    const int slot = warp / WPI, wi = warp % WPI;
    const bool swiglu = flags & TL_EPI_SWIGLU;
    const int n_out = swiglu ? npairs : N;

    for (int base = blockIdx.x * SLOTS; base < n_items; base += gridDim.x * SLOTS) {
        const int item = base + slot;
        float acc[ROWS][M];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int m = 0; m < M; ++m) acc[r][m] = 0.f;
        if (item < n_items) {
            const int row0 = item * ROWS;
            const uint4* wrow[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                int row = row0 + r;
                if (row >= N) row = N - 1;   // clamp (result discarded)
                wrow[r] = reinterpret_cast<const uint4*>(W + (size_t)row * K);
            }
#pragma unroll 2
            for (int v = wi * 32 + lane; v < nvec; v += WPI * 32) {
                uint4 wv[ROWS];
#pragma unroll
                for (int r = 0; r < ROWS; ++r) wv[r] = ldg_nc_v4(wrow[r] + v);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    const uint4 xv = reinterpret_cast<const uint4*>(xs + (size_t)m * K)[v];
                    const uint32_t* x32 = reinterpret_cast<const uint32_t*>(&xv);
                    float xf[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        xf[2 * j] = bf16_lo(x32[j]);
                        xf[2 * j + 1] = bf16_hi(x32[j]);
                    }
#pragma unroll
                    for (int r = 0; r < ROWS; ++r) {
                        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(&wv[r]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            acc[r][m] = fmaf(bf16_lo(w32[j]), xf[2 * j], acc[r][m]);
                            acc[r][m] = fmaf(bf16_hi(w32[j]), xf[2 * j + 1], acc[r][m]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float t = warp_sum(acc[r][m]);
                if (lane == 0) red[warp * (ROWS * M) + r * M + m] = t;
            }
        __syncthreads();
        // ---- epilogue: one thread per (slot, pair g, token m)
        if (tid < SLOTS * G * M) {
            const int s = tid / (G * M), rem = tid - s * (G * M);
            const int g = rem / M, m = rem - g * M;
            const int it = base + s;
            const int r0 = (it * G + g) * 2;
            if (it < n_items && r0 < N) {
                float v0 = 0.f, v1 = 0.f;
#pragma unroll
                for (int w = 0; w < WPI; ++w) {
                    v0 += red[(s * WPI + w) * (ROWS * M) + (2 * g) * M + m];
                    v1 += red[(s * WPI + w) * (ROWS * M) + (2 * g + 1) * M + m];
                }
                if (flags & TL_EPI_BIAS) {
                    v0 += bf2f(bias[r0]);
                    v1 += bf2f(bias[r0 + 1]);
                }
                if (swiglu) {
                    const float gate = rbf(v0), up = rbf(v1);
                    y[(size_t)m * n_out + (r0 >> 1)] = f2bf(rbf(silu_f(gate)) * up);
                } else {
                    float t0 = rbf(v0), t1 = rbf(v1);
                    if (flags & TL_EPI_RESIDUAL) {
                        t0 += bf2f(residual[(size_t)m * N + r0]);
                        t1 += bf2f(residual[(size_t)m * N + r0 + 1]);
                    }
                    y[(size_t)m * N + r0] = f2bf(t0);
                    if (r0 + 1 < N) y[(size_t)m * N + r0 + 1] = f2bf(t1);
                }
            }
        }
        __syncthreads();
    }
}

template <int M, int G, int WPI>
static int launch_gemv(const void* x, const void* W, void* y, int N, int K, const void* bias, const void* residual,
                       const void* norm_w, float eps, int flags, cudaStream_t st) {
    constexpr int SLOTS = GEMV_WARPS / WPI;
    auto kern = gemv_kernel<M, G, WPI>;
    const size_t smem = (size_t)M * K * sizeof(bf16) + (size_t)GEMV_WARPS * 2 * G * M * sizeof(float);
    static bool attr_done = false;   // per instantiation
    if (!attr_done) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return check_launch("tl_gemv_bf16 (smem attr)");
        attr_done = true;
    }
    TL_REQUIRE(smem <= 200 * 1024, TL_ERR_INVALID, "tl_gemv_bf16: M*K too large for shared memory (%zu B)", smem);
    const int n_items = ((N >> 1) + G - 1) / G;
    const int per_sm = smem > 100 * 1024 ? 1 : (smem > 64 * 1024 ? 2 : 4);
    int grid = (n_items + SLOTS - 1) / SLOTS;
    const int cap = sm_count() * per_sm;
    if (grid > cap) grid = cap;
    kern<<<grid, GEMV_THREADS, smem, st>>>((const bf16*)x, (const bf16*)W, (bf16*)y, N, K, (const bf16*)bias,
                                           (const bf16*)residual, (const bf16*)norm_w, eps, flags);
    return check_launch("tl_gemv_bf16");
}

template <int M, int G>
static int dispatch_wpi(int wpi, const void* x, const void* W, void* y, int N, int K, const void* bias,
                        const void* residual, const void* norm_w, float eps, int flags, cudaStream_t st) {
    switch (wpi) {
        case 1: return launch_gemv<M, G, 1>(x, W, y, N, K, bias, residual, norm_w, eps, flags, st);
        case 2: return launch_gemv<M, G, 2>(x, W, y, N, K, bias, residual, norm_w, eps, flags, st);
        case 4: return launch_gemv<M, G, 4>(x, W, y, N, K, bias, residual, norm_w, eps, flags, st);
        default: return launch_gemv<M, G, 8>(x, W, y, N, K, bias, residual, norm_w, eps, flags, st);
    }
}

template <int M>
static int dispatch_g(int g, int wpi, const void* x, const void* W, void* y, int N, int K, const void* bias,
                      const void* residual, const void* norm_w, float eps, int flags, cudaStream_t st) {
    switch (g) {
        case 4: return dispatch_wpi<M, 4>(wpi, x, W, y, N, K, bias, residual, norm_w, eps, flags, st);
        case 2: return dispatch_wpi<M, 2>(wpi, x, W, y, N, K, bias, residual, norm_w, eps, flags, st);
        default: return dispatch_wpi<M, 1>(wpi, x, W, y, N, K, bias, residual, norm_w, eps, flags, st);
    }
}

int gemv_stream_dispatch(const void* x, const void* W, void* y, int M, int N, int K, const void* bias,
                         const void* residual, const void* norm_w, float eps, int flags, const void* pf_ptr, size_t pf_bytes,
                         cudaStream_t st);
int gemv_mma_dispatch(const void* x, const void* W, void* y, int M, int N, int K, const void* bias, const void* residual,
                      const void* norm_w, float eps, int flags, cudaStream_t st);

static bool use_stream_kernel() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("TL_GEMV_IMPL");
        v = (e && e[0] == 'r') ? 0 : 1;     // TL_GEMV_IMPL=reg forces the register-streaming kernel (A/B tests)
    }
    return v == 1;
}

}  // namespace tl

extern "C" int tl_gemv_bf16_pf(const void* x, const void* W, void* y, int M, int N, int K, const void* bias,
                               const void* residual, const void* norm_w, float eps, int flags, const void* next_W,
                               size_t next_bytes, void* stream);

extern "C" int tl_gemv_bf16(const void* x, const void* W, void* y, int M, int N, int K, const void* bias,
                            const void* residual, const void* norm_w, float eps, int flags, void* stream) {
    return tl_gemv_bf16_pf(x, W, y, M, N, K, bias, residual, norm_w, eps, flags, nullptr, 0, stream);
}

extern "C" int tl_gemv_bf16_pf(const void* x, const void* W, void* y, int M, int N, int K, const void* bias,
                               const void* residual, const void* norm_w, float eps, int flags, const void* next_W,
                               size_t next_bytes, void* stream) {
    using namespace tl;
    TL_REQUIRE(M >= 1 && M <= 8, TL_ERR_INVALID, "tl_gemv_bf16: M=%d outside 1..8 (use tl_gemm_bf16)", M);
    TL_REQUIRE(K % 8 == 0 && N % 2 == 0 && N > 0 && K > 0, TL_ERR_INVALID,
               "tl_gemv_bf16: need K %% 8 == 0 and N even (N=%d K=%d)", N, K);
    TL_REQUIRE(!(flags & ~(TL_EPI_BIAS | TL_EPI_RESIDUAL | TL_EPI_SWIGLU)), TL_ERR_INVALID,
               "tl_gemv_bf16: unsupported flags 0x%x", flags);
    TL_REQUIRE(!((flags & TL_EPI_SWIGLU) && (flags & TL_EPI_RESIDUAL)), TL_ERR_INVALID,
               "tl_gemv_bf16: SWIGLU and RESIDUAL are exclusive");
    TL_REQUIRE(!(flags & TL_EPI_BIAS) || bias, TL_ERR_INVALID, "tl_gemv_bf16: BIAS flag without bias pointer");
    TL_REQUIRE(!(flags & TL_EPI_RESIDUAL) || residual, TL_ERR_INVALID, "tl_gemv_bf16: RESIDUAL flag without pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (M >= 2 && use_stream_kernel()) {
        // 2..8 rows on mma.sync (gemv_mma.cu): correct, but measured SLOWER than the CUDA-core stream kernel in
        // round 1 (its per-row 1-2 KB bulk copies cap an SM at ~1 copy / 50 ns: 2.7 TB/s vs 3.5-5.9) -> opt-in only.
        const char* e = getenv("TL_GEMV_MMA");
        if (e && e[0] == '1') {
            const int rc = gemv_mma_dispatch(x, W, y, M, N, K, bias, residual, norm_w, eps, flags, st);
            if (rc != 1) return rc;
        }
    }
    const int nvec = K >> 3;
    int wpi = 1;
    while (wpi < 8 && (nvec + 32 * wpi - 1) / (32 * wpi) > 4) wpi <<= 1;
    const int slots = GEMV_WARPS / wpi;
    const int npairs = N >> 1;
    // larger G = more loads in flight per lane; shrink it until there are >= 4 rounds of items per CTA wave
    int g = 4;
    const int wave = sm_count() * 2 * slots;
    while (g > 1 && (npairs / g) < 4 * wave) g >>= 1;
    const int iters = (nvec + 32 * wpi - 1) / (32 * wpi);
    if (iters <= 4 && g < 2 && npairs >= 2 * wave) g = 2;
    // the M template is rounded up to 1/2/4/8 and the surplus rows are never written because the epilogue
    // indexes only m < M... (rows beyond M would read x out of bounds), so dispatch exactly for 1..4 and
    // split larger M into two calls.
    auto run = [&](int m, const bf16* xx, bf16* yy, const bf16* rr, bool last) -> int {
        if (use_stream_kernel()) {
            const int rc = gemv_stream_dispatch(xx, W, yy, m, N, K, bias, rr, norm_w, eps, flags, last ? next_W : nullptr,
                                                last && next_W ? next_bytes : 0, st);
            if (rc != 1) return rc;
        }
        switch (m) {
            case 1: return dispatch_g<1>(g, wpi, xx, W, yy, N, K, bias, rr, norm_w, eps, flags, st);
            case 2: return dispatch_g<2>(g, wpi, xx, W, yy, N, K, bias, rr, norm_w, eps, flags, st);
            case 3: return dispatch_g<3>(g, wpi, xx, W, yy, N, K, bias, rr, norm_w, eps, flags, st);
            default: return dispatch_g<4>(g, wpi, xx, W, yy, N, K, bias, rr, norm_w, eps, flags, st);
        }
    };
    const int n_out = (flags & TL_EPI_SWIGLU) ? N / 2 : N;
    int done = 0;
    while (done < M) {
        const int m = (M - done) > 4 ? 4 : (M - done);
        int rc = run(m, (const bf16*)x + (size_t)done * K, (bf16*)y + (size_t)done * n_out,
                     residual ? (const bf16*)residual + (size_t)done * N : nullptr, done + m == M);
        if (rc != TL_OK) return rc;
        done += m;
    }
    return TL_OK;
}
