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
    uint32_t timeout_ms;    // 0: wait without bound
};

// Where a kernel publishes one fp32 value of this rank to every rank (d(logit_scale) partial of the row block): the
// value goes to dst[0 .. n_dst) (each already pointing at this rank's slot; one multicast address, or one address per
// rank), then - after a system-scope fence - the sequence number is released into every rank's flag word through the
// UNICAST mappings (ordering of a flag behind multicast data is only guaranteed this way).  n_dst == 0: nothing to do.
constexpr int PEER_MAXW = 16;
struct PeerPush {
    float* dst[PEER_MAXW];
    uint32_t* flag[PEER_MAXW];
    int n_dst, n_flag;
    uint32_t seq;
};


__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void peer_push_value(const PeerPush& w, float v) {
    for (int q = 0; q < w.n_dst; ++q) *reinterpret_cast<volatile float*>(w.dst[q]) = v;
    __threadfence_system();
    for (int q = 0; q < w.n_flag; ++q) st_release_sys_u32(w.flag[q], w.seq);
}
// the same by a whole (converged) warp: lane 0 stores the value, then one lane per rank releases that rank's flag - the
// W release stores (each a round trip over NVLink) overlap instead of queueing behind one another
__device__ __forceinline__ void peer_push_value_warp(const PeerPush& w, float v, int lane) {
    if (w.n_dst <= 0) return;
    if (lane == 0) {
        for (int q = 0; q < w.n_dst; ++q) *reinterpret_cast<volatile float*>(w.dst[q]) = v;
        __threadfence_system();
    }
    __syncwarp();
    if (lane < w.n_flag) st_release_sys_u32(w.flag[lane], w.seq);
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// A peer that never arrives must not hang the GPU for ever, and a kernel must NEVER continue on rows that have not
// arrived: after timeout_ms (0: wait without bound, hang detection is then the host's business) the error word tells the
// host which rank was missing and the kernel traps - the context is lost like after an NCCL watchdog abort, no wrong
// loss or gradient is ever produced.
__device__ __forceinline__ void peer_wait_flag(const PeerWait& w, int k) {
    const int sub = w.sub > 0 ? w.sub : 1, stride = w.stride > 0 ? w.stride : sub;
    for (int j = 0; j < sub; ++j) {
        const uint32_t* f = w.flags + k * stride + j;
        if ((int)(ld_acquire_sys_u32(f) - w.seq) >= 0) continue;
        const unsigned long long t0 = global_timer_ns();
        while ((int)(ld_acquire_sys_u32(f) - w.seq) < 0) {
            __nanosleep(100);
            if (w.timeout_ms != 0u && global_timer_ns() - t0 > (unsigned long long)w.timeout_ms * 1000000ull) {
                if (w.err != nullptr) {
                    *reinterpret_cast<volatile uint32_t*>(w.err) = 1u + (uint32_t)k;
                    __threadfence_system();
                }
                __trap();
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
