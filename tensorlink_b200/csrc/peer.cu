// Peer-memory mailboxes: the inter-shard hop of a decode step without a host round trip.
//
// The reference moves every activation between shards as a pickled tensor through its p2p node process
// (tensorlink/ml/module.py:438-462 -> p2p/torch_node.py:195-214 -> worker.py:297-357).  On one NVSwitch node the
// next stage's input buffer is simply mapped into this process (CUDA IPC): the last kernel of a stage stores its
// output rows straight into the neighbour's HBM over NVLink, a one-thread kernel then publishes a sequence number
// (release, system scope), and the neighbour's captured decode graph starts with a one-thread kernel that waits for
// that number (acquire, system scope).  Both counters live in device memory and advance inside the kernels, so the
// same CUDA graph replays for every token and the host enqueues the whole generation without synchronising.
//
// A waiting kernel never blocks the GPU for good: after `timeout_ns` it raises the mailbox error word and returns.
#include "common.cuh"

namespace tl {

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ++*want; spin until *flag >= *want.  wait_ns (optional) accumulates the time spent spinning.
__global__ void peer_wait_kernel(const uint32_t* __restrict__ flag, uint32_t* __restrict__ want, uint32_t* __restrict__ err,
                                 unsigned long long* __restrict__ wait_ns, uint64_t timeout_ns, int32_t* __restrict__ bump) {
    if (threadIdx.x != 0) return;
    const uint32_t w = *want + 1u;
    *want = w;
    if (bump) *bump += 1;             // e.g. the KV length of the step this wait opens (saves a launch per step)
    if (*err) return;                 // an earlier wait already gave up: do not stall the rest of the queue
    const uint64_t t0 = global_ns();
    uint64_t t = t0;
    // sequence numbers compare modulo 2^32
    while ((int32_t)(ld_acquire_sys(flag) - w) < 0) {
        __nanosleep(64);
        t = global_ns();
        if (t - t0 > timeout_ns) {
            *err = 1u;
            break;
        }
    }
    if (wait_ns) *wait_ns += (unsigned long long)(global_ns() - t0);
}

// ++*sent; publish it in the peer's flag after everything this stream wrote before
__global__ void peer_signal_kernel(uint32_t* __restrict__ flag_peer, uint32_t* __restrict__ sent, int32_t* __restrict__ bump) {
    if (threadIdx.x != 0) return;
    const uint32_t s = *sent + 1u;
    *sent = s;
    if (bump) *bump += 1;             // e.g. the cache write position, advanced once the step's kernels are done
    __threadfence_system();
    st_release_sys(flag_peer, s);
}

// copy `n16` 16-byte words into the peer buffer, then signal (one CTA; payloads are a few KB)
__global__ void peer_put_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16, uint32_t* __restrict__ flag_peer,
                                uint32_t* __restrict__ sent) {
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = *sent + 1u;
        *sent = s;
        __threadfence_system();
        st_release_sys(flag_peer, s);
    }
}

}  // namespace tl

using namespace tl;

extern "C" {

int tl_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
    TL_REQUIRE(bytes > 0 && ptr && handle64, TL_ERR_INVALID, "tl_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return check_launch("tl_peer_alloc (cudaMalloc)");
    if (cudaMemset(p, 0, bytes) != cudaSuccess) return check_launch("tl_peer_alloc (cudaMemset)");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
        cudaFree(p);
        return check_launch("tl_peer_alloc (cudaIpcGetMemHandle)");
    }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return TL_OK;
}

int tl_peer_open(const unsigned char* handle64, void** ptr) {
    TL_REQUIRE(handle64 && ptr, TL_ERR_INVALID, "tl_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return check_launch("tl_peer_open");
    *ptr = p;
    return TL_OK;
}

int tl_peer_close(void* ptr) {
    if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) return check_launch("tl_peer_close");
    return TL_OK;
}

int tl_peer_free(void* ptr) {
    if (ptr && cudaFree(ptr) != cudaSuccess) return check_launch("tl_peer_free");
    return TL_OK;
}

int tl_peer_wait(const uint32_t* flag_local, uint32_t* want_dev, uint32_t* err_dev, uint64_t* wait_ns_dev, uint64_t timeout_ns,
                 int32_t* bump_dev, void* stream) {
    TL_REQUIRE(flag_local && want_dev && err_dev, TL_ERR_INVALID, "tl_peer_wait: null pointer");
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flag_local, want_dev, err_dev, (unsigned long long*)wait_ns_dev,
                                                         timeout_ns ? timeout_ns : 10000000000ull, bump_dev);
    return check_launch("tl_peer_wait");
}

int tl_peer_signal(uint32_t* flag_peer, uint32_t* sent_dev, int32_t* bump_dev, void* stream) {
    TL_REQUIRE(flag_peer && sent_dev, TL_ERR_INVALID, "tl_peer_signal: null pointer");
    peer_signal_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flag_peer, sent_dev, bump_dev);
    return check_launch("tl_peer_signal");
}

int tl_peer_put(void* dst_peer, const void* src, size_t bytes, uint32_t* flag_peer, uint32_t* sent_dev, void* stream) {
    TL_REQUIRE(dst_peer && src && flag_peer && sent_dev, TL_ERR_INVALID, "tl_peer_put: null pointer");
    TL_REQUIRE(bytes % 16 == 0 && (((uintptr_t)dst_peer | (uintptr_t)src) & 15) == 0, TL_ERR_INVALID,
               "tl_peer_put: payload must be 16-byte aligned and a multiple of 16 bytes");
    peer_put_kernel<<<1, 256, 0, (cudaStream_t)stream>>>((uint4*)dst_peer, (const uint4*)src, bytes / 16, flag_peer, sent_dev);
    return check_launch("tl_peer_put");
}

}  // extern "C"
