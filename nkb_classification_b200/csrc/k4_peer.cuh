// Device-side protocol of the NVLink peer-memory exchange (K4'), shared by the standalone kernel (k4_peer.cu) and by
// the fused heads step (k2_fused.cu), whose last phase is this exchange.  See k4_peer.cu for the protocol itself.
#pragma once
#include "k2_common.cuh"

namespace nkbk {

constexpr int PEER_MAX_WORLD = 16;
constexpr int PEER_MAX_SLICES = 148;                        // payload slices == flag words per (parity, source rank)
constexpr int PEER_THREADS = 256;
constexpr int PEER_MIN_VECS_PER_SLICE = 32;
constexpr int PEER_TAIL_VECS = (2 * K2_MAX_TASKS + 3) / 4;  // per-slice copy of [loss_sum T | denom T]
constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

// What a kernel needs to reach every rank's inbox.
struct PeerLinks {
    int4* data[PEER_MAX_WORLD];          // every rank's inbox area (data[rank] is local)
    unsigned int* flags[PEER_MAX_WORLD]; // every rank's flag area
    unsigned int* ctl;                   // local: {step, ticket, status, -}
    int rank, world;
    long long slot_vecs, cap_vecs;       // slot = [cap_vecs payload | PEER_MAX_SLICES * PEER_TAIL_VECS tails]
};

// The payload (fp32 reduce buffer, then the int64 step confusion counts, both padded to 16-byte vectors) is cut into
// at most PEER_MAX_SLICES slices.  The cut depends on the payload size ONLY, so ranks whose local batches (and hence
// grids) differ still agree on which flag word guards which vectors.
__host__ __device__ inline void peer_slicing(long long total_vecs, int& per_slice, int& n_slices) {
    long long per = (total_vecs + PEER_MAX_SLICES - 1) / PEER_MAX_SLICES;
    if (per < PEER_MIN_VECS_PER_SLICE) per = PEER_MIN_VECS_PER_SLICE;
    per_slice = (int)per;
    n_slices = (int)((total_vecs + per - 1) / per);
}

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ int peer_task_of(const K2Seg& seg, int c) {
    int t = 0;
    while (t + 1 < seg.T && c >= seg.off[t + 1]) ++t;
    return t;
}

// Lane `r` (r < world, r != rank) of the calling CTA publishes "my slice `slice` of step `step` is in your inbox" to
// rank r and then waits for rank r's counterpart.  The caller has joined the CTA's stores with a __syncthreads()
// before, and joins again afterwards.  A wait that exceeds PEER_TIMEOUT_NS raises the status word (ctl[2] = r + 1)
// instead of hanging; the step's results are then invalid (the kernels poison the loss with NaN).
__device__ __forceinline__ void peer_publish(const PeerLinks& L, int r, long long par, int slice, unsigned int step) {
    // st.release.sys is itself the (cumulative) system-scope release: it orders the whole CTA's payload stores -- joined
    // to this thread by the barrier -- before the flag; a separate __threadfence_system() in front only added a second
    // round trip to the peer
    st_release_sys(L.flags[r] + (par * L.world + L.rank) * PEER_MAX_SLICES + slice, step);
}
__device__ __forceinline__ void peer_wait(const PeerLinks& L, int r, long long par, int slice, unsigned int step) {
    const unsigned int* f = L.flags[L.rank] + (par * L.world + r) * PEER_MAX_SLICES + slice;
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while (ld_acquire_sys(f) != step) {
        if ((++spins & 0x3ffu) == 0u && global_timer_ns() - t0 > PEER_TIMEOUT_NS) {
            atomicExch(L.ctl + 2, 1u + (unsigned int)r);  // status: peer r never arrived
            break;
        }
    }
}

// The same for a CTA that owns SEVERAL slices (first, first + stride, ... < n_slices): one system-scope fence covers all
// of them -- a release per flag would serialise one round trip per slice -- then relaxed flag stores; on the way in,
// relaxed polling of every flag and one fence at the end.
__device__ __forceinline__ void peer_publish_many(const PeerLinks& L, int r, long long par, int first, int stride,
                                                  int n_slices, unsigned int step) {
    __threadfence_system();
    unsigned int* f = L.flags[r] + (par * L.world + L.rank) * PEER_MAX_SLICES;
    for (int sl = first; sl < n_slices; sl += stride) st_relaxed_sys(f + sl, step);
}
__device__ __forceinline__ void peer_wait_many(const PeerLinks& L, int r, long long par, int first, int stride,
                                               int n_slices, unsigned int step) {
    const unsigned int* f = L.flags[L.rank] + (par * L.world + r) * PEER_MAX_SLICES;
    const unsigned long long t0 = global_timer_ns();
    for (int sl = first; sl < n_slices; sl += stride) {
        unsigned int spins = 0;
        while (ld_relaxed_sys(f + sl) != step) {
            if ((++spins & 0x3ffu) == 0u && global_timer_ns() - t0 > PEER_TIMEOUT_NS) {
                atomicExch(L.ctl + 2, 1u + (unsigned int)r);  // status: peer r never arrived
                break;
            }
        }
    }
    __threadfence_system();   // acquire side: the payload reads that follow see what the flags announce
}

// Host side (k4_peer.cu): fills `L` from the connected state; NKBK_E_NCCL when the transport is not connected,
// NKBK_E_SHAPE when the payload exceeds the capacity given to nkbk_peer_init.
int peer_links(PeerLinks& L, long long n_f32, long long n_i64, const char* who);

}  // namespace nkbk
