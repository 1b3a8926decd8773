// K4': the exchange step of the path fused with K2's finalize, over NVLink peer memory.
//
// One kernel per training step replaces { ncclAllReduce(fp32 reduce buffer), ncclAllReduce(int64 confusion
// counts), k2_heads_finalize }:
//
//   push     every CTA copies its slice of the local payload (head-gradient sums | loss sums | denominators |
//            step confusion counts) straight into the inbox slot [parity][my rank] of EVERY peer with 16-byte
//            stores over NVLink/NVSwitch (cudaIpc-mapped peer memory), plus a private copy of the 2T-float
//            loss/denominator tail, then publishes one flag word per (source rank, CTA) with st.release.sys;
//   wait     lane r of the CTA spins (ld.acquire.sys, bounded by a timeout) on the flag rank r wrote for this CTA:
//            no grid-wide or device-wide barrier, CTAs only ever wait for their own counterparts on other GPUs;
//   reduce   the CTA sums its slice over the `world` inbox slots in RANK ORDER (own slice read from the local
//            buffer), so every rank obtains bit-identical sums, deterministically;
//   finalize divides dW / db by the global denominators, writes the per-task mean losses, folds the summed step
//            confusion counts into the epoch totals and clears the step counts (what k2_heads_finalize does).
//
// Slots are double-buffered by step parity and flags carry the step number, which is enough to make reuse safe: a
// rank can only be pushing step s+2 into a slot once every peer has LAUNCHED its step s+1 kernel, i.e. finished
// reading step s (stream order).  The step counter lives in device memory and is advanced by the last CTA to
// finish, so the launch is a constant and can be captured in a CUDA graph.
//
// Memory is allocated by the library (cudaMalloc) and exchanged as cudaIpcMemHandle_t; torch.distributed only
// carries the 64-byte handles.  NCCL (k4_comm.cu) remains available as the plain baseline for the same step.
#include "k4_peer.cuh"

#include <string.h>

namespace nkbk {

struct PeerParams {
    PeerLinks L;
    float* reduce_buf;
    long long n_f32, vf;                 // floats / 16-byte vectors of the fp32 payload
    long long* cm_step;
    long long* cm_total;
    long long n_cm, vi;
    float* out_loss;
    int NC, D, vecs_per_cta;
    K2Seg seg;
};

// 16-byte vector `v` of the local payload: fp32 part first (zero padded), then the int64 counts (zero padded).
__device__ __forceinline__ int4 peer_load_local(const PeerParams& p, long long v) {
    if (v < p.vf) {
        const long long i = 4 * v;
        if (i + 3 < p.n_f32) return *reinterpret_cast<const int4*>(p.reduce_buf + i);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i + 0 < p.n_f32) x.x = p.reduce_buf[i + 0];
        if (i + 1 < p.n_f32) x.y = p.reduce_buf[i + 1];
        if (i + 2 < p.n_f32) x.z = p.reduce_buf[i + 2];
        return *reinterpret_cast<int4*>(&x);
    }
    const long long j = 2 * (v - p.vf);
    longlong2 x = make_longlong2(0, 0);
    if (j + 0 < p.n_cm) x.x = p.cm_step[j + 0];
    if (j + 1 < p.n_cm) x.y = p.cm_step[j + 1];
    return *reinterpret_cast<int4*>(&x);
}

__global__ void __launch_bounds__(PEER_THREADS) k4_peer_allreduce_finalize(const PeerParams p) {
    __shared__ float tail_s[2 * K2_MAX_TASKS];  // reduced [loss_sum T | denom T]
    __shared__ unsigned int step_s;
    __shared__ bool last_s;
    const PeerLinks& L = p.L;
    const int tid = threadIdx.x, cta = blockIdx.x;   // one payload slice per CTA (peer_slicing)
    const int T = p.seg.T, world = L.world, rank = L.rank;
    if (tid == 0) step_s = *reinterpret_cast<volatile unsigned int*>(L.ctl) + 1u;
    __syncthreads();
    const unsigned int step = step_s;
    const long long par = step & 1u;
    const long long v0 = (long long)cta * p.vecs_per_cta;
    const long long v1 = min(v0 + (long long)p.vecs_per_cta, p.vf + p.vi);
    const long long nW = (long long)p.NC * p.D, tail0 = nW + p.NC;
    const int tail_vecs = (2 * T + 3) / 4;
    const long long my_slot = (par * world + rank) * L.slot_vecs;  // where my data lands in every peer's inbox
    const long long tail_off = L.cap_vecs + (long long)cta * PEER_TAIL_VECS;

    // ---- push: my slice (and my copy of the tail) into every peer's inbox ----
    if (world > 1) {
        for (long long v = v0 + tid; v < v1; v += PEER_THREADS) {
            const int4 x = peer_load_local(p, v);
            for (int k = 1; k < world; ++k) {
                const int q = (rank + k) % world;  // staggered so the ranks do not all hit the same peer first
                L.data[q][my_slot + v] = x;
            }
        }
        if (tid < tail_vecs) {
            float4 x;
            float* xf = reinterpret_cast<float*>(&x);
#pragma unroll
            for (int k = 0; k < 4; ++k) xf[k] = (4 * tid + k < 2 * T) ? p.reduce_buf[tail0 + 4 * tid + k] : 0.f;
            for (int k = 1; k < world; ++k) {
                const int q = (rank + k) % world;
                L.data[q][my_slot + tail_off + tid] = *reinterpret_cast<int4*>(&x);
            }
        }
        __syncthreads();
        if (tid < world && tid != rank) {
            peer_publish(L, tid, par, cta, step);
            peer_wait(L, tid, par, cta, step);   // no grid-wide barrier: only this slice's counterparts
        }
        __syncthreads();
    }

    // ---- the global loss sums / denominators, identical in every CTA of every rank (rank order) ----
    if (tid < 2 * T) {
        float s = 0.f;
        for (int r = 0; r < world; ++r) {
            float x;
            if (r == rank) x = p.reduce_buf[tail0 + tid];
            else x = __ldcg(reinterpret_cast<const float*>(L.data[rank] + (par * world + r) * L.slot_vecs + tail_off) + tid);
            s = (r == 0) ? x : s + x;
        }
        tail_s[tid] = s;
    }
    __syncthreads();

    // ---- reduce my slice in rank order, finalize in place ----
    for (long long v = v0 + tid; v < v1; v += PEER_THREADS) {
        if (v < p.vf) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < world; ++r) {
                int4 raw;
                if (r == rank) raw = peer_load_local(p, v);
                else raw = __ldcg(L.data[rank] + (par * world + r) * L.slot_vecs + v);
                const float4 x = *reinterpret_cast<float4*>(&raw);
                if (r == 0) acc = x;
                else { acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
            }
            float* a = reinterpret_cast<float*>(&acc);
            const long long i0 = 4 * v;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long i = i0 + k;
                if (i < tail0) {
                    const int c = (i < nW) ? (int)(i / p.D) : (int)(i - nW);
                    const float dn = tail_s[T + peer_task_of(p.seg, c)];
                    a[k] = dn > 0.f ? __fdiv_rn(a[k], dn) : 0.f;
                }
            }
            // The [loss_sum | denom] tail is NOT written here: other CTAs of this rank may still have to read the
            // LOCAL values (for their push and for their rank-ordered sum); the last CTA to finish writes it below.
            if (i0 + 3 < tail0) *reinterpret_cast<float4*>(p.reduce_buf + i0) = acc;
            else
                for (int k = 0; k < 4; ++k)
                    if (i0 + k < tail0) p.reduce_buf[i0 + k] = a[k];
        } else {
            longlong2 acc = make_longlong2(0, 0);
            for (int r = 0; r < world; ++r) {
                int4 raw;
                if (r == rank) raw = peer_load_local(p, v);
                else raw = __ldcg(L.data[rank] + (par * world + r) * L.slot_vecs + v);
                const longlong2 x = *reinterpret_cast<longlong2*>(&raw);
                acc.x += x.x;
                acc.y += x.y;
            }
            const long long j = 2 * (v - p.vf);
            if (j < p.n_cm) { p.cm_total[j] += acc.x; p.cm_step[j] = 0; }
            if (j + 1 < p.n_cm) { p.cm_total[j + 1] += acc.y; p.cm_step[j + 1] = 0; }
        }
    }

    // ---- the last CTA to finish advances the step counter (every CTA read it before taking a ticket) ----
    // It also publishes the global loss sums / denominators (left unnormalised: k2_heads_demb reads the
    // denominators) and the mean losses -- only now, when no CTA of this rank needs the local values any more.
    // A peer wait that timed out (status word set) poisons the losses with NaN: the step's sums are incomplete.
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        last_s = atomicAdd(L.ctl + 1, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (last_s) {
        if (tid < 2 * T) p.reduce_buf[tail0 + tid] = tail_s[tid];
        if (tid == 0) {
            if (p.out_loss != nullptr) {
                const bool bad = *reinterpret_cast<volatile unsigned int*>(L.ctl + 2) != 0u;
                float total = 0.f;
                for (int t = 0; t < T; ++t) {
                    const float l = tail_s[T + t] > 0.f ? __fdiv_rn(tail_s[t], tail_s[T + t]) : 0.f;
                    p.out_loss[t] = bad ? __int_as_float(0x7fc00000) : l;
                    total += l;
                }
                p.out_loss[T] = bad ? __int_as_float(0x7fc00000) : total;
            }
            L.ctl[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile unsigned int*>(L.ctl) = step;
        }
    }
}

struct PeerState {
    bool on = false, connected = false;
    int rank = 0, world = 0, device = 0;
    long long max_f32 = 0, max_i64 = 0, cap_vecs = 0, slot_vecs = 0;
    size_t data_bytes = 0, flags_off = 0, ctl_off = 0, bytes = 0;
    char* local = nullptr;
    char* mapped[PEER_MAX_WORLD] = {};
};
static PeerState g_peer;

static int fill_seg_peer(K2Seg& seg, const int32_t* seg_offsets, int T) {
    if (!seg_offsets || T < 1 || T > K2_MAX_TASKS) {
        set_error("nkbk_peer_allreduce_finalize: bad seg_offsets / T=%d", T);
        return NKBK_E_ARG;
    }
    seg.T = T;
    for (int t = 0; t <= T; ++t) seg.off[t] = seg_offsets[t];
    if (seg.off[0] != 0) { set_error("nkbk_peer_allreduce_finalize: seg_offsets[0] != 0"); return NKBK_E_ARG; }
    for (int t = 0; t < T; ++t)
        if (seg.off[t + 1] <= seg.off[t]) { set_error("nkbk_peer_allreduce_finalize: task %d has no classes", t); return NKBK_E_ARG; }
    if (seg.off[T] > K2_MAX_NC) { set_error("nkbk_peer_allreduce_finalize: %d classes > %d", seg.off[T], K2_MAX_NC); return NKBK_E_SHAPE; }
    return NKBK_OK;
}

}  // namespace nkbk

using namespace nkbk;

extern "C" int nkbk_peer_init(int rank, int world, int device, int64_t max_f32, int64_t max_i64, void* out_handle_host) {
    NKBK_CHECK_ARG(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "nkbk_peer_init: rank=%d world=%d (max %d)",
                   rank, world, PEER_MAX_WORLD);
    NKBK_CHECK_ARG(max_f32 >= 1 && max_i64 >= 0 && out_handle_host, "nkbk_peer_init: max_f32=%lld max_i64=%lld",
                   (long long)max_f32, (long long)max_i64);
    if (g_peer.on) {
        set_error("nkbk_peer_init: already initialised (world %d); call nkbk_peer_shutdown first", g_peer.world);
        return NKBK_E_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == NKBK_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    NKBK_CHECK_CUDA(cudaSetDevice(device));
    PeerState s;
    s.rank = rank; s.world = world; s.device = device; s.max_f32 = max_f32; s.max_i64 = max_i64;
    s.cap_vecs = (max_f32 + 3) / 4 + (max_i64 + 1) / 2;
    s.slot_vecs = s.cap_vecs + (long long)PEER_MAX_SLICES * PEER_TAIL_VECS;
    s.data_bytes = (size_t)2 * world * s.slot_vecs * 16;
    s.flags_off = (s.data_bytes + 255) & ~size_t(255);
    s.ctl_off = (s.flags_off + (size_t)2 * world * PEER_MAX_SLICES * 4 + 255) & ~size_t(255);
    s.bytes = s.ctl_off + 256;
    void* ptr = nullptr;
    NKBK_CHECK_CUDA(cudaMalloc(&ptr, s.bytes));
    s.local = static_cast<char*>(ptr);
    cudaError_t e = cudaMemset(s.local, 0, s.bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();  // zeroed before any peer can learn the handle
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (e == cudaSuccess && world > 1) e = cudaIpcGetMemHandle(&h, s.local);
    if (e != cudaSuccess) {
        cudaFree(s.local);
        set_error("nkbk_peer_init: %s", cudaGetErrorString(e));
        return NKBK_E_CUDA;
    }
    memcpy(out_handle_host, &h, sizeof(h));
    s.mapped[rank] = s.local;
    s.on = true;
    s.connected = (world == 1);
    g_peer = s;
    return NKBK_OK;
}

extern "C" int nkbk_peer_connect(const void* handles_host) {
    if (!g_peer.on) { set_error("nkbk_peer_connect: call nkbk_peer_init first"); return NKBK_E_ARG; }
    if (g_peer.connected) return NKBK_OK;
    NKBK_CHECK_ARG(handles_host != nullptr, "nkbk_peer_connect: NULL handles");
    NKBK_CHECK_CUDA(cudaSetDevice(g_peer.device));
    for (int r = 0; r < g_peer.world; ++r) {
        if (r == g_peer.rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles_host) + (size_t)r * NKBK_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("nkbk_peer_connect: cudaIpcOpenMemHandle(rank %d) failed: %s (no P2P / NVLink path between the "
                      "devices, or a different IPC namespace)", r, cudaGetErrorString(e));
            for (int q = 0; q < r; ++q)
                if (q != g_peer.rank && g_peer.mapped[q]) { cudaIpcCloseMemHandle(g_peer.mapped[q]); g_peer.mapped[q] = nullptr; }
            return NKBK_E_UNSUPPORTED;
        }
        g_peer.mapped[r] = static_cast<char*>(ptr);
    }
    g_peer.connected = true;
    return NKBK_OK;
}

extern "C" int nkbk_peer_world(void) { return (g_peer.on && g_peer.connected) ? g_peer.world : 0; }

int nkbk::peer_links(PeerLinks& L, long long n_f32, long long n_i64, const char* who) {
    if (!g_peer.on || !g_peer.connected) {
        set_error("%s: peer memory not connected (nkbk_peer_init + nkbk_peer_connect)", who);
        return NKBK_E_NCCL;
    }
    if (n_f32 > g_peer.max_f32 || n_i64 > g_peer.max_i64) {
        set_error("%s: payload %lld f32 + %lld i64 exceeds the capacity given to nkbk_peer_init (%lld, %lld)", who, n_f32,
                  n_i64, g_peer.max_f32, g_peer.max_i64);
        return NKBK_E_SHAPE;
    }
    memset(&L, 0, sizeof(L));
    for (int r = 0; r < g_peer.world; ++r) {
        L.data[r] = reinterpret_cast<int4*>(g_peer.mapped[r]);
        L.flags[r] = reinterpret_cast<unsigned int*>(g_peer.mapped[r] + g_peer.flags_off);
    }
    L.ctl = reinterpret_cast<unsigned int*>(g_peer.local + g_peer.ctl_off);
    L.rank = g_peer.rank; L.world = g_peer.world;
    L.slot_vecs = g_peer.slot_vecs; L.cap_vecs = g_peer.cap_vecs;
    return NKBK_OK;
}

extern "C" int nkbk_peer_allreduce_finalize(float* reduce_buf, int D, const int32_t* seg_offsets, int T, float* out_loss,
                                            int64_t* cm_total, int64_t* cm_step, int64_t n_cm, void* stream) {
    if (!g_peer.on || !g_peer.connected) {
        set_error("nkbk_peer_allreduce_finalize: peer memory not connected (nkbk_peer_init + nkbk_peer_connect)");
        return NKBK_E_NCCL;
    }
    K2Seg seg;
    int rc = fill_seg_peer(seg, seg_offsets, T);
    if (rc) return rc;
    NKBK_CHECK_ARG(reduce_buf && D >= 1, "nkbk_peer_allreduce_finalize: NULL reduce_buf or D=%d", D);
    NKBK_CHECK_ARG(n_cm >= 0 && (n_cm == 0 || (cm_total && cm_step)), "nkbk_peer_allreduce_finalize: bad confusion buffers");
    const int NC = seg.off[T];
    const long long n_f32 = (long long)NC * D + NC + 2LL * T;
    PeerParams p;
    memset(&p, 0, sizeof(p));
    rc = peer_links(p.L, n_f32, n_cm, "nkbk_peer_allreduce_finalize");
    if (rc) return rc;
    NKBK_CHECK_ARG((reinterpret_cast<uintptr_t>(reduce_buf) & 15) == 0, "nkbk_peer_allreduce_finalize: reduce_buf must be 16-byte aligned");
    p.reduce_buf = reduce_buf; p.n_f32 = n_f32; p.vf = (n_f32 + 3) / 4;
    p.cm_step = reinterpret_cast<long long*>(cm_step); p.cm_total = reinterpret_cast<long long*>(cm_total);
    p.n_cm = n_cm; p.vi = (n_cm + 1) / 2;
    p.out_loss = out_loss; p.NC = NC; p.D = D; p.seg = seg;
    int per = 0, blocks = 0;
    peer_slicing(p.vf + p.vi, per, blocks);   // one slice per CTA: the same cut the fused heads step uses
    p.vecs_per_cta = per;
    k4_peer_allreduce_finalize<<<blocks, PEER_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
    NKBK_CHECK_LAUNCH("k4_peer_allreduce_finalize");
    return NKBK_OK;
}

extern "C" int nkbk_peer_status(int32_t* out_status_host) {
    NKBK_CHECK_ARG(out_status_host != nullptr, "nkbk_peer_status: NULL output");
    if (!g_peer.on) { *out_status_host = 0; return NKBK_OK; }
    unsigned int ctl[4];
    NKBK_CHECK_CUDA(cudaMemcpy(ctl, g_peer.local + g_peer.ctl_off, sizeof(ctl), cudaMemcpyDeviceToHost));
    *out_status_host = (int32_t)ctl[2];
    return NKBK_OK;
}

extern "C" int nkbk_peer_disconnect(void) {
    if (!g_peer.on) return NKBK_OK;
    cudaSetDevice(g_peer.device);
    cudaDeviceSynchronize();
    for (int r = 0; r < g_peer.world; ++r)
        if (r != g_peer.rank && g_peer.mapped[r]) {
            cudaIpcCloseMemHandle(g_peer.mapped[r]);
            g_peer.mapped[r] = nullptr;
        }
    g_peer.connected = false;
    return NKBK_OK;
}

extern "C" int nkbk_peer_shutdown(void) {
    if (!g_peer.on) return NKBK_OK;
    nkbk_peer_disconnect();
    cudaFree(g_peer.local);
    g_peer = PeerState();
    return NKBK_OK;
}
