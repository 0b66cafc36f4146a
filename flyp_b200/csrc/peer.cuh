// peer.cuh — device-side readiness flags for data that another GPU of the node writes into this GPU's memory over
// NVLink (peer-mapped exchange segments, comm.cu).  A producer rank writes its rows and then a monotonically increasing
// sequence number into flags[rank]; consumers poll the flag of the rank whose rows they are about to read.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace flyp {

// Rows [k * rows_per_flag, (k + 1) * rows_per_flag) of the operand are valid once (int)(flags[k] - seq) >= 0.
// flags == nullptr: nothing to wait for (single rank, or data already known to be complete).
struct PeerWait {
    const uint32_t* flags;  // producer k is done when its words flags[k * stride + j], j < sub, have all reached seq
    uint32_t seq;
    int n_flags;
    int rows_per_flag;
    int sub;                // flag words per producer (one per CTA of the pushing kernel); 0 is read as 1
    int stride;             // distance in words between the flag groups of consecutive producers; 0 is read as sub
    uint32_t* err;          // optional (host-mapped) word set to 1 + k when waiting for flags[k] timed out
};

__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Bounded spin (4 s): a peer that never arrives must not hang the GPU; the error word tells the host.
__device__ __forceinline__ void peer_wait_flag(const PeerWait& w, int k) {
    const int sub = w.sub > 0 ? w.sub : 1, stride = w.stride > 0 ? w.stride : sub;
    for (int j = 0; j < sub; ++j) {
        const uint32_t* f = w.flags + k * stride + j;
        if ((int)(ld_acquire_sys_u32(f) - w.seq) >= 0) continue;
        const unsigned long long t0 = global_timer_ns();
        while ((int)(ld_acquire_sys_u32(f) - w.seq) < 0) {
            __nanosleep(100);
            if (global_timer_ns() - t0 > 4000000000ull) {
                if (w.err != nullptr) *reinterpret_cast<volatile uint32_t*>(w.err) = 1u + (uint32_t)k;
                return;
            }
        }
    }
}
// wait for every producer whose rows intersect [row0, row1); then order the following TMA (async proxy) reads
__device__ __forceinline__ void peer_wait_rows(const PeerWait& w, int row0, int row1) {
    if (w.flags == nullptr) return;
    int k0 = row0 / w.rows_per_flag, k1 = (row1 - 1) / w.rows_per_flag;
    if (k1 >= w.n_flags) k1 = w.n_flags - 1;
    for (int k = k0; k <= k1; ++k) peer_wait_flag(w, k);
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void peer_wait_all(const PeerWait& w) {
    if (w.flags == nullptr) return;
    for (int k = 0; k < w.n_flags; ++k) peer_wait_flag(w, k);
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

}  // namespace flyp
