// Shared pieces of the tcgen05 GEMM kernels (1-CTA gemm.cu, 2-CTA gemm2.cu): tensor-map construction and the
// TMEM -> registers -> HBM epilogue for one 32-column chunk of one accumulator row.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tl {

constexpr int BM = 128;
constexpr int BK = 64;

// 2-D bf16 row-major tensor [outer, inner] with leading dimension ld (elements); 128B-swizzled boxes (cached)
int make_tensor_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                    uint32_t box_outer);

// r[32]: fp32 accumulators of row `row`, columns col0..col0+31 (as loaded by tcgen05.ld 32x32b.x32)
__device__ __forceinline__ void gemm_epilogue_chunk(const uint32_t (&r)[32], void* Cv, int row, int col0, int M, int N, int ldc,
                                                    const bf16* __restrict__ bias, const bf16* __restrict__ residual, int ldr,
                                                    int flags) {
    if (!(row < M && col0 < N)) return;
    const bool swiglu = flags & TL_EPI_SWIGLU;
    const bool out_f32 = flags & TL_EPI_OUT_F32;
        float v[32];
    #pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int ncols = min(32, N - col0);
        if (flags & TL_EPI_BIAS) {
    #pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < ncols) v[j] += bf2f(bias[col0 + j]);
        }
        if (swiglu) {
            // interleaved (gate, up) column pairs -> 16 outputs
            bf16* dst = reinterpret_cast<bf16*>(Cv) + (size_t)row * ldc + (col0 >> 1);
            uint32_t o[8];
    #pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float g0 = rbf(v[4 * j]), u0 = rbf(v[4 * j + 1]);
                const float g1 = rbf(v[4 * j + 2]), u1 = rbf(v[4 * j + 3]);
                o[j] = pack_bf16(rbf(silu_f(g0)) * u0, rbf(silu_f(g1)) * u1);
            }
            if (ncols == 32) {
                reinterpret_cast<uint4*>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
                reinterpret_cast<uint4*>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
            } else {
                for (int j = 0; j < ncols / 2; ++j)
                    dst[j] = reinterpret_cast<bf16*>(o)[j];
            }
        } else if (out_f32) {
            float* dst = reinterpret_cast<float*>(Cv) + (size_t)row * ldc + col0;
            if (flags & TL_EPI_ACCUM) {
    #pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < ncols) v[j] += dst[j];
            }
            if (ncols == 32) {
    #pragma unroll
                for (int j = 0; j < 8; ++j)
                    reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
                for (int j = 0; j < ncols; ++j) dst[j] = v[j];
            }
        } else {
            bf16* dst = reinterpret_cast<bf16*>(Cv) + (size_t)row * ldc + col0;
    #pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = rbf(v[j]);     // the Linear's own bf16 output
            if (flags & TL_EPI_RESIDUAL) {
                const bf16* rs = residual + (size_t)row * ldr + col0;
                if (ncols == 32) {
    #pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 u = reinterpret_cast<const uint4*>(rs)[q];
                        const uint32_t* u32 = reinterpret_cast<const uint32_t*>(&u);
    #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[8 * q + 2 * j] += bf16_lo(u32[j]);
                            v[8 * q + 2 * j + 1] += bf16_hi(u32[j]);
                        }
                    }
                } else {
                    for (int j = 0; j < ncols; ++j) v[j] += bf2f(rs[j]);
                }
            }
            if (flags & TL_EPI_ACCUM) {
                if (ncols == 32) {
    #pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 u = reinterpret_cast<const uint4*>(dst)[q];
                        const uint32_t* u32 = reinterpret_cast<const uint32_t*>(&u);
    #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            v[8 * q + 2 * j] += bf16_lo(u32[j]);
                            v[8 * q + 2 * j + 1] += bf16_hi(u32[j]);
                        }
                    }
                } else {
                    for (int j = 0; j < ncols; ++j) v[j] += bf2f(dst[j]);
                }
            }
            if (ncols == 32) {
    #pragma unroll
                for (int q = 0; q < 4; ++q)
                    reinterpret_cast<uint4*>(dst)[q] =
                        make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                                   pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
            } else {
                for (int j = 0; j < ncols; ++j) dst[j] = f2bf(v[j]);
            }
        }
}

}  // namespace tl
