// Shared pieces of the tcgen05 GEMM kernels (1-CTA gemm.cu, 2-CTA gemm2.cu): tensor-map construction and the
// TMEM -> registers -> shared-memory staging -> HBM epilogue.
//
// Epilogue layout.  tcgen05.ld (32x32b) hands every thread of an epilogue warp ONE accumulator row, so a direct store
// makes each warp instruction touch 32 different rows (32 partial sectors, row pitch = N*2 bytes apart); measured in
// round 1 this cost ~15 us per 128x256 tile and capped the GEMM at ~0.8x cuBLAS (and 0.45x at K = 896).  Here a warp
// stages its 32 rows x 64 columns through a private 32 x 144-byte shared-memory tile and then writes (and, for the
// residual / accumulate operands, reads) global memory with 8 lanes per row: every instruction moves four complete
// 128-byte lines.  Both shared-memory access patterns (row-per-lane and 8-lanes-per-row) are bank-conflict free with
// the 144-byte pitch.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace tl {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_PITCH = 144;                       // bytes per staged row (128 + 16 pad)
constexpr int EPI_STAGE_BYTES = 32 * EPI_PITCH;      // per epilogue warp
constexpr int EPI_SMEM_BYTES = 4 * EPI_STAGE_BYTES;  // four epilogue warps

// 2-D bf16 row-major tensor [outer, inner] with leading dimension ld (elements); 128B-swizzled boxes (cached)
int make_tensor_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                    uint32_t box_outer);

// coalesced [32 rows x 128 bytes] global -> staging tile (rows beyond M / pieces beyond `valid_bytes` read as zero)
__device__ __forceinline__ void epi_load_tile(unsigned char* stg, const unsigned char* gbase, size_t row_pitch_bytes, int row0,
                                              int M, int valid_bytes, int lane) {
    const int piece = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row0 + rr < M && piece * 16 < valid_bytes)
            v = *reinterpret_cast<const uint4*>(gbase + (size_t)rr * row_pitch_bytes + piece * 16);
        *reinterpret_cast<uint4*>(stg + rr * EPI_PITCH + piece * 16) = v;
    }
    __syncwarp();
}
// staging tile -> coalesced global store of `row_bytes` (64 or 128) per row
__device__ __forceinline__ void epi_store_tile(const unsigned char* stg, unsigned char* gbase, size_t row_pitch_bytes, int row0,
                                               int M, int row_bytes, int valid_bytes, int lane) {
    __syncwarp();
    if (row_bytes == 128) {
        const int piece = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3);
            if (row0 + rr < M && piece * 16 < valid_bytes)
                *reinterpret_cast<uint4*>(gbase + (size_t)rr * row_pitch_bytes + piece * 16) =
                    *reinterpret_cast<const uint4*>(stg + rr * EPI_PITCH + piece * 16);
        }
    } else {   // 64 bytes per row: 4 lanes per row, 8 rows per instruction
        const int piece = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + (lane >> 2);
            if (row0 + rr < M && piece * 16 < valid_bytes)
                *reinterpret_cast<uint4*>(gbase + (size_t)rr * row_pitch_bytes + piece * 16) =
                    *reinterpret_cast<const uint4*>(stg + rr * EPI_PITCH + piece * 16);
        }
    }
    __syncwarp();
}

// One 64-column chunk of one epilogue warp: r0 / r1 = fp32 accumulators of this thread's row for columns
// [col0, col0+32) and [col0+32, col0+64).  row0 = first row of the warp (thread `lane` owns row0 + lane).
__device__ __forceinline__ void gemm_epilogue_chunk64(const uint32_t (&r0)[32], const uint32_t (&r1)[32], unsigned char* stg,
                                                      void* Cv, int row0, int lane, int col0, int M, int N, int ldc,
                                                      const bf16* __restrict__ bias, const bf16* __restrict__ residual,
                                                      int ldr, int flags, int max_cols = 64) {
    if (row0 >= M || col0 >= N) return;                 // warp-uniform
    const int ncols = min(max_cols, N - col0);          // multiple of 8 (max_cols = 32 for 32-wide tiles)
    float v[64];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        v[j] = __uint_as_float(r0[j]);
        v[32 + j] = __uint_as_float(r1[j]);
    }
    if (flags & TL_EPI_BIAS) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q * 8 < ncols) {
                const uint4 b4 = *reinterpret_cast<const uint4*>(bias + col0 + q * 8);
                const uint32_t* b32 = reinterpret_cast<const uint32_t*>(&b4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[q * 8 + 2 * j] += bf16_lo(b32[j]);
                    v[q * 8 + 2 * j + 1] += bf16_hi(b32[j]);
                }
            }
        }
    }
    unsigned char* myrow = stg + lane * EPI_PITCH;
    if (flags & TL_EPI_SWIGLU) {
        // interleaved (gate, up) column pairs -> 32 bf16 outputs (64 bytes) per row
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float g0 = rbf(v[4 * j]), u0 = rbf(v[4 * j + 1]);
            const float g1 = rbf(v[4 * j + 2]), u1 = rbf(v[4 * j + 3]);
            o[j] = pack_bf16(rbf(silu_f(g0)) * u0, rbf(silu_f(g1)) * u1);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(myrow + q * 16) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        unsigned char* g = reinterpret_cast<unsigned char*>(Cv) + ((size_t)row0 * ldc + (col0 >> 1)) * 2;
        epi_store_tile(stg, g, (size_t)ldc * 2, row0, M, 64, ncols, lane);
        return;
    }
    if (flags & TL_EPI_OUT_F32) {
        // fp32 output: two half-chunks of 32 columns (128 bytes per row each)
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            const int c0 = col0 + hb * 32;
            if (c0 >= N || hb * 32 >= max_cols) break;
            const int vb = min(32, N - c0) * 4;
            unsigned char* g = reinterpret_cast<unsigned char*>(Cv) + ((size_t)row0 * ldc + c0) * 4;
            if (flags & TL_EPI_ACCUM) {
                epi_load_tile(stg, g, (size_t)ldc * 4, row0, M, vb, lane);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 o = *reinterpret_cast<const float4*>(myrow + q * 16);
                    v[hb * 32 + 4 * q] += o.x; v[hb * 32 + 4 * q + 1] += o.y; v[hb * 32 + 4 * q + 2] += o.z; v[hb * 32 + 4 * q + 3] += o.w;
                }
                __syncwarp();
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(myrow + q * 16) =
                    make_float4(v[hb * 32 + 4 * q], v[hb * 32 + 4 * q + 1], v[hb * 32 + 4 * q + 2], v[hb * 32 + 4 * q + 3]);
            epi_store_tile(stg, g, (size_t)ldc * 4, row0, M, 128, vb, lane);
        }
        return;
    }
    // bf16 output: the Linear's own rounding, then the optional residual / accumulate adds (each rounded like torch)
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = rbf(v[j]);
    unsigned char* g = reinterpret_cast<unsigned char*>(Cv) + ((size_t)row0 * ldc + col0) * 2;
    if (flags & TL_EPI_RESIDUAL) {
        const unsigned char* rg = reinterpret_cast<const unsigned char*>(residual) + ((size_t)row0 * ldr + col0) * 2;
        epi_load_tile(stg, rg, (size_t)ldr * 2, row0, M, ncols * 2, lane);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 u = *reinterpret_cast<const uint4*>(myrow + q * 16);
            const uint32_t* u32 = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[8 * q + 2 * j] += bf16_lo(u32[j]);
                v[8 * q + 2 * j + 1] += bf16_hi(u32[j]);
            }
        }
        __syncwarp();
    }
    if (flags & TL_EPI_ACCUM) {
        epi_load_tile(stg, g, (size_t)ldc * 2, row0, M, ncols * 2, lane);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint4 u = *reinterpret_cast<const uint4*>(myrow + q * 16);
            const uint32_t* u32 = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[8 * q + 2 * j] += bf16_lo(u32[j]);
                v[8 * q + 2 * j + 1] += bf16_hi(u32[j]);
            }
        }
        __syncwarp();
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(myrow + q * 16) =
            make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                       pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
    epi_store_tile(stg, g, (size_t)ldc * 2, row0, M, 128, ncols * 2, lane);
}

}  // namespace tl
