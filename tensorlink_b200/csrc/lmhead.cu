// Final RMSNorm + lm_head + greedy argmax for decode-shaped inputs (M <= 8 rows).
// logits are produced by the weight-streaming GEMV (bf16, as the reference's lm_head emits them), then a
// two-stage argmax picks the lowest index among equal maxima, which is what torch.argmax returns.
#include "common.cuh"

namespace tl {

constexpr int AM_PARTS = 64, AM_THREADS = 256;

__global__ void __launch_bounds__(AM_THREADS) argmax_part_kernel(const bf16* __restrict__ logits, float* __restrict__ pval,
                                                                 int* __restrict__ pidx, int V) {
    const int m = blockIdx.y, part = blockIdx.x;
    const int per = (V + AM_PARTS - 1) / AM_PARTS;
    const int lo = part * per, hi = min(V, lo + per);
    const bf16* row = logits + (size_t)m * V;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lo + threadIdx.x; i < hi; i += AM_THREADS) {
        const float v = bf2f(row[i]);
        if (v > best) { best = v; bi = i; }     // ascending i per thread: first max kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __shared__ float sv[AM_THREADS / 32];
    __shared__ int si[AM_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < AM_THREADS / 32; ++w)
            if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
        pval[m * AM_PARTS + part] = best;
        pidx[m * AM_PARTS + part] = bi;
    }
}

__global__ void argmax_final_kernel(const float* __restrict__ pval, const int* __restrict__ pidx,
                                    int64_t* __restrict__ ids_out) {
    const int m = blockIdx.x, lane = threadIdx.x;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < AM_PARTS; i += 32) {
        const float v = pval[m * AM_PARTS + i];
        const int ix = pidx[m * AM_PARTS + i];
        if (v > best || (v == best && ix < bi)) { best = v; bi = ix; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) ids_out[m] = (bi == 0x7fffffff) ? 0 : (int64_t)bi;
}

}  // namespace tl

extern "C" {

size_t tl_lmhead_ws(int M, int V) {
    return (size_t)M * V * sizeof(tl::bf16) + (size_t)M * tl::AM_PARTS * (sizeof(float) + sizeof(int)) + 256;
}

int tl_argmax_bf16(const void* logits, int64_t* ids_out, void* workspace, size_t ws_bytes, int M, int V, void* stream) {
    using namespace tl;
    const size_t need = (size_t)M * AM_PARTS * (sizeof(float) + sizeof(int));
    TL_REQUIRE(ws_bytes >= need, TL_ERR_WORKSPACE, "tl_argmax_bf16: workspace %zu < %zu", ws_bytes, need);
    if (M == 0) return TL_OK;
    float* pval = (float*)workspace;
    int* pidx = (int*)(pval + (size_t)M * AM_PARTS);
    cudaStream_t st = (cudaStream_t)stream;
    argmax_part_kernel<<<dim3(AM_PARTS, M), AM_THREADS, 0, st>>>((const bf16*)logits, pval, pidx, V);
    argmax_final_kernel<<<M, 32, 0, st>>>(pval, pidx, ids_out);
    return check_launch("tl_argmax_bf16");
}

int tl_lmhead_argmax(const void* x, const void* W, const void* norm_w, float eps, int64_t* ids_out, void* logits_out,
                     void* workspace, size_t ws_bytes, int M, int V, int H, void* stream) {
    using namespace tl;
    TL_REQUIRE(M >= 1 && M <= 8, TL_ERR_INVALID, "tl_lmhead_argmax: M=%d outside 1..8", M);
    TL_REQUIRE(ws_bytes >= tl_lmhead_ws(M, V), TL_ERR_WORKSPACE, "tl_lmhead_argmax: workspace %zu < %zu", ws_bytes,
               tl_lmhead_ws(M, V));
    unsigned char* ws = (unsigned char*)workspace;
    void* logits = logits_out ? logits_out : (void*)ws;
    size_t off = ((size_t)M * V * sizeof(bf16) + 255) & ~(size_t)255;
    int rc = tl_gemv_bf16(x, W, logits, M, V, H, nullptr, nullptr, norm_w, eps, 0, stream);
    if (rc != TL_OK) return rc;
    return tl_argmax_bf16(logits, ids_out, ws + off, ws_bytes - off, M, V, stream);
}

}  // extern "C"
