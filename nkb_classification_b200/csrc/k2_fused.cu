// K2 fused: the whole heads step of a training batch as ONE persistent kernel --
//
//   forward (segmented GEMM, all task heads) -> per-task softmax / CE | focal loss / dlogits -> K3 (argmax + confusion
//   counts, privatised per CTA in shared memory) -> dW / db -> cross-CTA sum -> [K4' exchange over NVLink peer memory]
//   -> finalize (divide by the global denominators, mean losses, fold the step confusion counts into the epoch totals)
//
// replacing { k2_heads_forward_v3, k2_heads_dw, k4_peer_allreduce_finalize | k2_heads_finalize } (3 launches, the
// embeddings read from HBM twice).  A CTA (one per SM, 16 warps, cooperative launch) owns a contiguous range of rows and
// brings it in with TMA bulk copies (cp.async.bulk + mbarrier complete_tx) through a two-stage shared-memory ring of
// <= 8-row tiles; the head weights (fp32 [NC][D]) are staged once per CTA, also by a bulk copy.  Two passes over the
// CTA's rows, so that the forward accumulators and the dW accumulators are never live together (16 warps fit):
//
//   pass 1    forward, tile by tile: 16 warps = (row group of 4 rows) x (K slice): a lane owns 4 consecutive columns of
//             a 128-column chunk, accumulates 4 rows x <= 16 classes as FFMA2 pairs (two rows per packed fp32 FMA, the
//             weight as the broadcast operand), folds the 64 partial sums with the transposing shuffle reduction; K
//             slices are added in a fixed order through shared memory.  The logits of ALL rows stay in shared memory.
//   epilogue  one warp per row, all rows of the CTA at once: max / exp / sum / log-sum-exp, loss term, dlogit scale,
//             argmax (first maximum, NaN maximal), confusion count into the CTA's shared-memory histogram (global 64-bit
//             atomics only when the matrix has more than KF_MAX_SMEM_HIST bins), logits / probabilities / dlogits to
//             global memory, the dlogits also to shared memory.
//   pass 2    dW / db, NEWEST tile first: the last two tiles of pass 1 are still in the ring (up to 16 rows per CTA,
//             i.e. B <= 2368, never leave shared memory); older tiles come back from L2 (they were loaded with
//             evict_last) into whichever slot is free -- the dead weight area serves as a third slot, and its first
//             re-fetch is issued before the epilogue.  Embedding bytes cross the HBM interface ONCE.  Thread = KQ
//             column quads x NCP classes in registers (dlogit broadcast from shared memory, FFMA2 over adjacent columns).
//   reduce    every CTA writes its [dW | db] partial; after one grid-wide barrier the payload (the reduce buffer of
//             nkbk.h, then the int64 step confusion counts) is cut into <= 148 slices (peer_slicing) and a CTA sums its
//             slices over the CTAs in CTA order, S-way split per vector with a fixed combine tree -- deterministic.
//   exchange  (world > 1, mode PEER) the slice goes straight from registers into every peer's inbox; flags, waits and
//             the rank-ordered sum are those of k4_peer.cu (same inbox, same slicing, same step counter): the all-reduce
//             is the epilogue of K2, not a launch of its own.
//
// Exact fp32 (FFMA) like k2_heads.cu; bf16 embeddings are widened on load.  Shapes that do not fit (NC * D too large
// for the register accumulators, weights + ring + per-row logits larger than shared memory, unaligned rows) return 0
// from launch_k2_fused and take the three-kernel path.
#include "k4_peer.cuh"

#include <cooperative_groups.h>
#include <algorithm>
#include <stdlib.h>
#include <string.h>

namespace cg = cooperative_groups;

namespace nkbk {

constexpr int KF_THREADS = 512;
constexpr int KF_WARPS = KF_THREADS / 32;
constexpr int KF_TILE_ROWS = 8;         // rows per ring stage (4 when shared memory is short)
constexpr int KF_STAGES = 2;
constexpr int KF_NCB = 16;              // classes per forward pass
constexpr int KF_MAX_SMEM_HIST = 4096;  // confusion bins privatised per CTA in shared memory
constexpr int KF_TIMING_SLOTS = 16;
constexpr int KF_Y_IGN = 0x40000000;    // shared-memory label code: the label equals ignore_index (K3 still counts it)
constexpr int KF_MAX_ACC = 64;          // dW accumulator registers per thread: 4 * NCP * KQ

struct KFParams {
    K2FwdParams f;          // emb, W, bias, labels, outputs, cm_step (loss_part / counters of the 3-kernel path unused)
    float* reduce_buf;      // [dW NC*D | db NC | loss_sum T | denom T]
    float* part;            // [grid][part_stride]: each CTA's dW | db partial sums
    float* loss_part;       // [grid][2T]
    float* out_loss;        // [T+1] (modes FINALIZE / PEER), may be NULL
    long long* cm_total;    // epoch confusion totals (modes FINALIZE / PEER), may be NULL
    long long n_cm;         // confusion bins folded into cm_total / exchanged (0 in mode SUMS)
    long long n_bins;       // confusion bins K3 counts into (0 = no counts)
    long long part_stride;  // floats, multiple of 4
    int rows_per_cta, tile_rows, mode;
    int G, QTHR;            // dW thread map: class groups, threads per class group (multiple of 32)
    int NCD;                // G * NCP: padded class count of the shared-memory dlogit tile
    int hist_smem;          // confusion counts privatised in shared memory
    int w_slot;             // the weight area is large enough to serve as a third tile slot in the dW pass
    PeerLinks peer;         // mode PEER only
    unsigned long long* timing;   // debug: [grid][KF_TIMING_SLOTS] clock64 stamps per phase (NULL = off)
};

struct KFSmem {  // byte offsets into dynamic shared memory
    size_t ring, w, zs, dls, zpart, eps, lacc, bias, cw, ylab, task_of, cm_off, hist, tail, bars, total;
};

// `rows` = rows per CTA (logits, dlogits and labels of ALL rows of the CTA stay in shared memory between the passes)
__host__ __device__ inline KFSmem kf_smem(int D, int NC, int T, int NCD, int tile_rows, int rows, int es, int hist_bins) {
    KFSmem s;
    size_t o = 0;
    s.ring = o;    o += (size_t)KF_STAGES * tile_rows * D * es;
    s.w = o;       o += (size_t)NC * D * 4;
    s.zs = o;      o += (size_t)rows * NC * 4;
    o = (o + 15) & ~size_t(15);
    s.dls = o;     o += (size_t)rows * NCD * 4;
    s.zpart = o;   o += (size_t)2 * KF_WARPS * 4 * KF_NCB * 4;   // [2][K slices * rows <= 64][16], double buffered
    s.eps = o;     o += (size_t)2 * KF_WARPS * (NC + 4 * T) * 4;            // per half-warp
    s.lacc = o;    o += (size_t)2 * KF_WARPS * 2 * T * 4;                   // per half-warp
    s.bias = o;    o += (size_t)NC * 4;
    s.cw = o;      o += (size_t)NC * 4;
    s.ylab = o;    o += (size_t)rows * T * 4;
    s.task_of = o; o += (size_t)NC * 4;
    s.cm_off = o;  o += (size_t)T * 4;
    s.hist = o;    o += (size_t)hist_bins * 4;
    s.tail = o;    o += (size_t)4 * T * 4;   // local and global [loss_sum | denom]
    o = (o + 15) & ~size_t(15);
    s.bars = o;    o += 8 * (KF_STAGES + 1);
    s.total = o;
    return s;
}

// ---- PTX helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t kf_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kf_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void kf_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void kf_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "KF_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KF_DONE_%=;\n\t"
        "bra KF_WAIT_%=;\n\t"
        "KF_DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void kf_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t kf_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t kf_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ unsigned long long kf_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long kf_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void kf_unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// acc.lo = fma(a.lo, s, acc.lo); acc.hi = fma(a.hi, s, acc.hi)  -- FFMA2, the scalar is a broadcast operand in SASS
__device__ __forceinline__ void kf_ffma2(unsigned long long& acc, unsigned long long a, float s) {
    unsigned long long s2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(s2) : "f"(s));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(s2));
}

// IEEE division, kept out of line: the finalize divisions sit in run-once code whose size is instruction-fetch time
__device__ __noinline__ float kf_div(float a, float b) { return __fdiv_rn(a, b); }

// four consecutive elements of a shared-memory embedding row as fp32
__device__ __forceinline__ float4 kf_lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 kf_lds4(const __nv_bfloat16* p) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
}

// Transposing warp reduction (as in k2_heads.cu): every lane holds 32 partial sums; lane L returns the warp-wide total
// of partial sum number L.
__device__ __forceinline__ float kf_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// Sum `n` floats spaced `stride` apart in a fixed order with one warp (lane-strided partial sums, xor tree).
__device__ __forceinline__ float kf_warp_fixed_sum(const float* base, int n, int64_t stride, int lane) {
    float s = 0.f;
    for (int i0 = lane; i0 < n; i0 += 32 * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i0 + 32 * u < n) v[u] = __ldcg(base + (int64_t)(i0 + 32 * u) * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i0 + 32 * u < n) s += v[u];
    }
    return warp_sum(s);
}

// ---- the kernel ------------------------------------------------------------------------------------------------------
template <typename ET, int NCP, int KQ>
__global__ void __launch_bounds__(KF_THREADS, 1) k2_fused_step(const KFParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ unsigned int step_s;
    __shared__ bool last_s;
    // This CTA is resident: a kernel enqueued behind this one as a programmatic dependent (K1 of the next batch under
    // nkbk_k1_overlap_previous) may start on the SMs this grid does not occupy.  No effect on ordinary launches.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const K2FwdParams& f = p.f;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = f.seg.T, NC = f.NC, D = f.D, NCD = p.NCD, TR = p.tile_rows;
    const int n_hist = p.hist_smem ? (int)p.n_bins : 0;
    const KFSmem SM = kf_smem(D, NC, T, NCD, TR, p.rows_per_cta, (int)sizeof(ET), n_hist);
    ET* ring = reinterpret_cast<ET*>(smem_raw + SM.ring);
    const float* wsm = reinterpret_cast<const float*>(smem_raw + SM.w);
    float* zs = reinterpret_cast<float*>(smem_raw + SM.zs);          // [rows][NC] logits of every row of this CTA
    float* dls = reinterpret_cast<float*>(smem_raw + SM.dls);        // [rows][NCD] dlogits (zero padded)
    float* zpart = reinterpret_cast<float*>(smem_raw + SM.zpart);    // [2][KS][RT][16] K-slice partial logits
    float* lacc = reinterpret_cast<float*>(smem_raw + SM.lacc);      // [2 * WARPS][2T] loss / denominator sums
    float* bias_s = reinterpret_cast<float*>(smem_raw + SM.bias);    // [NC]
    float* cw_s = reinterpret_cast<float*>(smem_raw + SM.cw);        // [NC] class weights (1 when there are none)
    int* ylab = reinterpret_cast<int*>(smem_raw + SM.ylab);          // [rows][T] label | KF_Y_IGN, or -1 (out of range)
    int* task_of = reinterpret_cast<int*>(smem_raw + SM.task_of);    // [NC]
    int* cm_off = reinterpret_cast<int*>(smem_raw + SM.cm_off);      // [T] first confusion bin of a task
    unsigned int* hist = reinterpret_cast<unsigned int*>(smem_raw + SM.hist);
    float* tail_l = reinterpret_cast<float*>(smem_raw + SM.tail);    // [2T] local loss sums / denominators
    float* tail_g = tail_l + 2 * T;                                  // [2T] the same over all ranks
    const uint32_t bars = kf_smem_u32(smem_raw + SM.bars);           // one per slot: ring stages, then the weight area
    if (p.timing != nullptr && tid == 0) {
        p.timing[blockIdx.x * KF_TIMING_SLOTS + 0] = kf_globaltimer();
        p.timing[blockIdx.x * KF_TIMING_SLOTS + 1] = (unsigned long long)clock64();
    }

    const int row_begin = blockIdx.x * p.rows_per_cta;
    const int row_end = min(f.B, row_begin + p.rows_per_cta);
    const int nrows_cta = row_end - row_begin;
    const int ntiles = (nrows_cta + TR - 1) / TR;
    const size_t stage_elems = (size_t)TR * D;
    const uint32_t row_bytes = (uint32_t)D * (uint32_t)sizeof(ET);

    // Tile slots: 0 / 1 = the ring stages, 2 = the weight area (dW pass only).  One mbarrier per slot; a slot's k-th
    // load completes phase k, so one parity bit per slot (uniform over the CTA) says what to wait for.
    auto slot_ptr = [&](int slot) -> ET* {
        return slot < KF_STAGES ? ring + (size_t)slot * stage_elems : reinterpret_cast<ET*>(smem_raw + SM.w);
    };
    auto issue_tile = [&](int i, int slot, bool keep) {   // one thread; the rows of a tile are contiguous: ONE bulk copy
        const int r0 = row_begin + i * TR;
        const int nr = min(TR, row_end - r0);
        const uint32_t bar = bars + 8u * slot;
        // forward pass: the rows come back for the dW pass -> keep them in L2; dW pass: last use
        const uint64_t pol = keep ? kf_policy_evict_last() : kf_policy_evict_first();
        kf_mbar_expect_tx(bar, (uint32_t)nr * row_bytes);
        kf_bulk_g2s(kf_smem_u32(slot_ptr(slot)), static_cast<const ET*>(f.emb) + (int64_t)r0 * D, (uint32_t)nr * row_bytes,
                    bar, pol);
    };
    uint32_t phases = 0u;   // bit `slot` = parity of the slot's next load
    auto wait_slot = [&](int slot) {
        kf_mbar_wait(bars + 8u * slot, (phases >> slot) & 1u);
        phases ^= 1u << slot;
    };

    // ---- prologue: barriers; bulk copies of the weights (one thread) and of the first two tiles (another); tables ----
    if (tid == 0) {
        for (int s = 0; s <= KF_STAGES; ++s) kf_mbar_init(bars + 8u * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < min(ntiles, KF_STAGES); ++i) issue_tile(i, i, i <= ntiles - 3);
    } else if (tid == 32) {
        const uint64_t polw = kf_policy_evict_last();   // every CTA stages the same weights: keep them in L2
        const uint32_t wbytes = (uint32_t)NC * (uint32_t)D * 4u;
        kf_mbar_expect_tx(bars + 8u * KF_STAGES, wbytes);
        const uint32_t wdst = kf_smem_u32(wsm);
        for (uint32_t o = 0; o < wbytes; o += 65536u)
            kf_bulk_g2s(wdst + o, reinterpret_cast<const char*>(f.W) + o, min(65536u, wbytes - o), bars + 8u * KF_STAGES, polw);
    }
    for (int c = tid; c < NC; c += KF_THREADS) {
        task_of[c] = peer_task_of(f.seg, c);
        bias_s[c] = __ldg(f.bias + c);
        cw_s[c] = f.class_weight ? __ldg(f.class_weight + c) : 1.f;
    }
    for (int t = tid; t < T; t += KF_THREADS) {
        int off = 0;
        for (int s = 0; s < t; ++s) {
            const int Cs = f.seg.off[s + 1] - f.seg.off[s];
            off += Cs * Cs;
        }
        cm_off[t] = off;
    }
    // labels of this CTA's rows (contiguous in global memory): in flight while the first tile arrives
#pragma unroll 1
    for (int i = tid; i < nrows_cta * T; i += KF_THREADS) {
        int code = -1;
        if (f.labels != nullptr) {
            const int64_t y = __ldg(f.labels + (int64_t)row_begin * T + i);
            const int t = T == 1 ? 0 : i % T;
            const int C = f.seg.off[t + 1] - f.seg.off[t];
            if (y >= 0 && y < C) code = (int)y | (y == f.ignore_index ? KF_Y_IGN : 0);
        }
        ylab[i] = code;
    }
#pragma unroll 1
    for (int i = tid; i < n_hist; i += KF_THREADS) hist[i] = 0u;
#pragma unroll 1
    for (int i = tid; i < 2 * KF_WARPS * 2 * T; i += KF_THREADS) lacc[i] = 0.f;
    // (no barrier here: every table is first read after the barrier of the first class pass)
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 2] = (unsigned long long)clock64();
    const int nchunks = (D + 127) >> 7;

    // ===================== pass 1: forward, tile by tile; logits of all rows stay in shared memory =====================
    wait_slot(KF_STAGES);   // the weights
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 3] = (unsigned long long)clock64();
    int zbuf = 0;
    for (int i = 0; i < ntiles; ++i) {
        const int s = i & (KF_STAGES - 1);
        const int nrows = min(TR, nrows_cta - i * TR);
        const ET* tile = ring + s * stage_elems;
        float* zt = zs + (size_t)i * TR * NC;
        wait_slot(s);
        const int KS = nrows > 4 ? KF_WARPS / 2 : KF_WARPS;   // K slices; row groups = WARPS / KS
        const int rg = warp / KS, ks = warp - rg * KS;
        const bool fwd_on = rg * 4 < nrows;
        for (int cb = 0; cb < NC; cb += KF_NCB) {
            float* zp_all = zpart + (size_t)zbuf * (KF_WARPS * 4 * KF_NCB);
            const int RT = KS == KF_WARPS ? 4 : 8;   // rows a K slice covers: KS * RT = 64 entries of 16 classes
            if (fwd_on) {
                unsigned long long a2[2][KF_NCB];
#pragma unroll
                for (int rp = 0; rp < 2; ++rp)
#pragma unroll
                    for (int c = 0; c < KF_NCB; ++c) a2[rp][c] = 0ull;
                const ET* t0 = tile + (size_t)(rg * 4) * D;
                const int ncls = min(KF_NCB, NC - cb);
                auto load_e = [&](float4 (&e)[4], int chunk) {
                    const int k = chunk * 128 + lane * 4;
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        e[r] = (chunk < nchunks && k < D) ? kf_lds4(t0 + (size_t)r * D + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                };
                float4 e[4];
                load_e(e, ks);
                for (int chunk = ks; chunk < nchunks; chunk += KS) {
                    const int k = chunk * 128 + lane * 4;
                    const int kk = k < D ? k : 0;   // lanes past D hold zeros in e: any valid W address
                    unsigned long long ex[2], ey[2], ez[2], ew[2];
#pragma unroll
                    for (int rp = 0; rp < 2; ++rp) {
                        ex[rp] = kf_pack(e[2 * rp].x, e[2 * rp + 1].x);
                        ey[rp] = kf_pack(e[2 * rp].y, e[2 * rp + 1].y);
                        ez[rp] = kf_pack(e[2 * rp].z, e[2 * rp + 1].z);
                        ew[rp] = kf_pack(e[2 * rp].w, e[2 * rp + 1].w);
                    }
                    load_e(e, chunk + KS);          // the next chunk's rows are in flight under this chunk's FFMAs
                    const float* wk = wsm + (size_t)cb * D + kk;
#pragma unroll
                    for (int c2 = 0; c2 < KF_NCB; c2 += 2) {
                        if (c2 >= ncls) break;  // warp-uniform
                        const float4 w0 = *reinterpret_cast<const float4*>(wk + (size_t)c2 * D);
                        const float4 w1 = *reinterpret_cast<const float4*>(wk + (size_t)min(c2 + 1, ncls - 1) * D);
#pragma unroll
                        for (int rp = 0; rp < 2; ++rp) {
                            kf_ffma2(a2[rp][c2], ex[rp], w0.x);
                            kf_ffma2(a2[rp][c2], ey[rp], w0.y);
                            kf_ffma2(a2[rp][c2], ez[rp], w0.z);
                            kf_ffma2(a2[rp][c2], ew[rp], w0.w);
                        }
                        if (c2 + 1 < ncls) {  // warp-uniform: a padded class costs no FFMAs
#pragma unroll
                            for (int rp = 0; rp < 2; ++rp) {
                                kf_ffma2(a2[rp][c2 + 1], ex[rp], w1.x);
                                kf_ffma2(a2[rp][c2 + 1], ey[rp], w1.y);
                                kf_ffma2(a2[rp][c2 + 1], ez[rp], w1.z);
                                kf_ffma2(a2[rp][c2 + 1], ew[rp], w1.w);
                            }
                        }
                    }
                }
                // 64 partial sums per lane (entry = row * 16 + class) -> lane L owns entries L and 32 + L
                float lo[32], hi[32];
#pragma unroll
                for (int c = 0; c < KF_NCB; ++c) {
                    kf_unpack(a2[0][c], lo[c], lo[16 + c]);
                    kf_unpack(a2[1][c], hi[c], hi[16 + c]);
                }
                const float s_lo = kf_reduce_scatter32(lo, lane);
                const float s_hi = kf_reduce_scatter32(hi, lane);
                float* zp = zp_all + ((size_t)ks * RT + rg * 4) * KF_NCB;
                zp[(lane >> 4) * KF_NCB + (lane & 15)] = s_lo;
                zp[(2 + (lane >> 4)) * KF_NCB + (lane & 15)] = s_hi;
            }
            __syncthreads();   // the only barrier of a class pass: zpart is double buffered (see below)
            if (tid < KF_TILE_ROWS * KF_NCB) {   // K slices in slice order, then the bias
                const int r = tid >> 4, c = tid & 15;
                if (r < nrows && cb + c < NC) {
                    float sum = zp_all[(size_t)r * KF_NCB + c];
                    for (int q = 1; q < KS; ++q) sum += zp_all[((size_t)q * RT + r) * KF_NCB + c];
                    zt[r * NC + cb + c] = sum + bias_s[cb + c];
                }
            }
            // These 128 threads read buffer `zbuf` while everybody moves on and fills the other one; `zbuf` itself is
            // written again two passes later, i.e. after the next pass's barrier, which the readers reach only when done.
            zbuf ^= 1;
        }
        // Everyone is past the barrier of the tile's last class pass = done reading the stage.  Refill it -- only when
        // a later tile needs it: the last two tiles stay resident for the dW pass.
        if (tid == 0 && i + KF_STAGES < ntiles) issue_tile(i + KF_STAGES, s, i + KF_STAGES <= ntiles - 3);
    }
    // The weights are dead once every warp has left pass 1: their area becomes a third tile slot, and the newest tile
    // that is no longer resident starts coming back (from L2) while the epilogue runs.
    __syncthreads();
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 4] = (unsigned long long)clock64();
    const int n_refetch = max(0, ntiles - KF_STAGES);
    const int nslots = p.w_slot ? KF_STAGES + 1 : KF_STAGES;
    const int sA = (ntiles - 1) & (KF_STAGES - 1), sB = (ntiles - 2) & (KF_STAGES - 1);
    auto refetch_slot = [&](int k) -> int {   // slot of the k-th re-fetched tile (tile ntiles - 3 - k)
        const int m = k % nslots;
        if (p.w_slot) return m == 0 ? KF_STAGES : (m == 1 ? sA : sB);
        return m == 0 ? sA : sB;
    };
    if (tid == 0 && p.w_slot && n_refetch > 0) issue_tile(ntiles - 3, KF_STAGES, false);

    // ===================== epilogue: one HALF-warp per row, all rows of the CTA =====================
    {
        const int hw = 2 * warp + (lane >> 4), sl = lane & 15;     // half-warp id, lane within it
        float* es = reinterpret_cast<float*>(smem_raw + SM.eps) + (size_t)hw * (NC + 4 * T);   // [NC] exp(z - max)
        float* mxs = es + NC;       // [T]
        float* lses = mxs + T;      // [T]
        float* qs = lses + T;       // [T] dlogit scale (0 when the row is ignored)
        float* ys = qs + T;         // [T] label as float (-1 when ignored)
        float* la = lacc + (size_t)hw * 2 * T;
        const int rounds = (nrows_cta + 2 * KF_WARPS - 1) / (2 * KF_WARPS);
        for (int rd = 0; rd < rounds; ++rd) {   // both halves of a warp walk the same number of rounds (__syncwarp)
            const int r = rd * 2 * KF_WARPS + hw;
            const bool on = r < nrows_cta;
            const int row = row_begin + r;
            const float* z = zs + (size_t)(on ? r : 0) * NC;
            if (on)
                for (int t = sl; t < T; t += 16) {
                    const int c0 = f.seg.off[t], C = f.seg.off[t + 1] - c0;
                    float mx = z[c0];
                    for (int j = 1; j < C; ++j) mx = fmaxf(mx, z[c0 + j]);
                    mxs[t] = mx;
                }
            __syncwarp();
            if (on)
                for (int c = sl; c < NC; c += 16) es[c] = expf(z[c] - mxs[task_of[c]]);
            __syncwarp();
            if (on)
                for (int t = sl; t < T; t += 16) {
                    const int c0 = f.seg.off[t], C = f.seg.off[t + 1] - c0;
                    const int code = ylab[r * T + t];
                    float se = 0.f;
                    for (int j = 0; j < C; ++j) se += es[c0 + j];
                    const float lse = mxs[t] + logf(se);
                    lses[t] = lse;
                    // K3: first maximum, NaN maximal (torch.argmax)
                    float best = z[c0];
                    int bi = 0;
                    for (int j = 1; j < C; ++j) {
                        const float v = z[c0 + j];
                        if (!(best != best) && (v > best || v != v)) { best = v; bi = j; }
                    }
                    float q = 0.f, loss_i = 0.f, den_i = 0.f, yf = -1.f;
                    if (code >= 0 && !(code & KF_Y_IGN)) {
                        const int y = code;
                        const float logpt = z[c0 + y] - lse;
                        const float a = cw_s[c0 + y];
                        if (f.loss_kind == NKBK_LOSS_FOCAL) {
                            const float pt = expf(logpt);
                            const float om = 1.f - pt;
                            const float gm = f.gamma;
                            float ft, dterm;  // ft = om^g ; dterm = g * pt * om^(g-1) * logpt
                            if (gm == 0.f) { ft = 1.f; dterm = 0.f; }
                            else {
                                const float pw1 = (gm == 1.f) ? 1.f : ((gm == 2.f) ? om : powf(om, gm - 1.f));
                                ft = pw1 * om;
                                dterm = gm * pt * pw1 * logpt;
                            }
                            loss_i = -a * ft * logpt;
                            q = a * (dterm - ft);
                            den_i = 1.f;
                        } else {
                            loss_i = -a * logpt;
                            q = -a;
                            den_i = a;
                        }
                        yf = (float)y;
                    }
                    qs[t] = q;
                    ys[t] = yf;
                    la[t] += loss_i;          // rows of this half-warp in row order: fixed
                    la[T + t] += den_i;
                    if (f.out_pred) f.out_pred[(int64_t)row * T + t] = bi;
                    if (f.cm_step != nullptr && code >= 0) {   // counts every label inside [0, C), ignored or not
                        const int y = code & ~KF_Y_IGN;
                        if (p.hist_smem) atomicAdd(hist + cm_off[t] + y * C + bi, 1u);
                        else {
                            int64_t off = 0;   // bins beyond int range: in 64 bits
                            for (int s2 = 0; s2 < t; ++s2) {
                                const int64_t Cs = f.seg.off[s2 + 1] - f.seg.off[s2];
                                off += Cs * Cs;
                            }
                            atomicAdd(f.cm_step + off + (int64_t)y * C + bi, 1ull);
                        }
                    }
                }
            __syncwarp();
            if (on)
                for (int c = sl; c < NCD; c += 16) {
                    float dl = 0.f;
                    if (c < NC) {
                        const int t = task_of[c];
                        const float zj = z[c];
                        const float pj = expf(zj - lses[t]);
                        const int64_t o = (int64_t)row * NC + c;
                        if (f.out_logits) f.out_logits[o] = zj;
                        if (f.out_probs) f.out_probs[o] = pj;
                        const float yf = ys[t];
                        dl = yf >= 0.f ? qs[t] * (((float)(c - f.seg.off[t]) == yf ? 1.f : 0.f) - pj) : 0.f;
                        f.dlogits[o] = dl;
                    }
                    dls[(size_t)r * NCD + c] = dl;
                }
            __syncwarp();   // the scratch is reused by the half-warp's next row
        }
    }
    __syncthreads();
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 5] = (unsigned long long)clock64();

    // ===================== pass 2: dW / db, newest tile first (the last two are still resident) =====================
    // thread map: class group g (NCP classes), column quads tq + j * QTHR
    const int g = tid / p.QTHR, tq = tid - g * p.QTHR;
    const bool dw_on = g < p.G;
    const int nquads = D >> 2;
    unsigned long long acc2[NCP][KQ][2];
#pragma unroll
    for (int c = 0; c < NCP; ++c)
#pragma unroll
        for (int j = 0; j < KQ; ++j) acc2[c][j][0] = acc2[c][j][1] = 0ull;
    float dbacc = 0.f;
    for (int step = 0; step < ntiles; ++step) {
        const int i = ntiles - 1 - step;                       // tile
        const int nrows = min(TR, nrows_cta - i * TR);
        int slot;
        if (step < KF_STAGES) slot = i & (KF_STAGES - 1);      // resident since pass 1 (waited for there)
        else {
            slot = refetch_slot(step - KF_STAGES);
            wait_slot(slot);
        }
        const ET* tile = slot_ptr(slot);
        const float* dlt = dls + (size_t)i * TR * NCD;
        if (dw_on) {
            // one-row software pipeline: row r + 1 (embedding quads + dlogits) is in flight under the FFMAs of row r
            float4 en[KQ], gn[NCP / 4];
            auto load_row = [&](int r) {
                const ET* er = tile + (size_t)r * D;
#pragma unroll
                for (int j = 0; j < KQ; ++j) {
                    const int q = tq + j * p.QTHR;
                    en[j] = q < nquads ? kf_lds4(er + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const float* dr = dlt + r * NCD + g * NCP;
#pragma unroll
                for (int c4 = 0; c4 < NCP / 4; ++c4) gn[c4] = *reinterpret_cast<const float4*>(dr + 4 * c4);
            };
            load_row(0);
            for (int r = 0; r < nrows; ++r) {
                unsigned long long e01[KQ], e23[KQ];
                float gg[NCP];
#pragma unroll
                for (int j = 0; j < KQ; ++j) {
                    e01[j] = kf_pack(en[j].x, en[j].y);
                    e23[j] = kf_pack(en[j].z, en[j].w);
                }
#pragma unroll
                for (int c4 = 0; c4 < NCP / 4; ++c4) {
                    gg[4 * c4] = gn[c4].x; gg[4 * c4 + 1] = gn[c4].y; gg[4 * c4 + 2] = gn[c4].z; gg[4 * c4 + 3] = gn[c4].w;
                }
                if (r + 1 < nrows) load_row(r + 1);
#pragma unroll
                for (int c = 0; c < NCP; ++c)
#pragma unroll
                    for (int j = 0; j < KQ; ++j) {
                        kf_ffma2(acc2[c][j][0], e01[j], gg[c]);
                        kf_ffma2(acc2[c][j][1], e23[j], gg[c]);
                    }
            }
        }
        if (tid < NC)
            for (int r = 0; r < nrows; ++r) dbacc += dlt[r * NCD + tid];
        // the slot is free: bring back the next tile that needs one (re-fetch k goes out after step k - w_slot)
        const int k = step + (p.w_slot ? 1 : 0);
        if (k < n_refetch) {
            __syncthreads();
            if (tid == 0) issue_tile(ntiles - 3 - k, refetch_slot(k), false);
        }
    }
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 6] = (unsigned long long)clock64();

    // ---- this CTA's partial sums ----
    const int64_t nW = (int64_t)NC * D, tail0 = nW + NC;
    {
        float* mypart = p.part + (int64_t)blockIdx.x * p.part_stride;
        if (dw_on) {
#pragma unroll
            for (int c = 0; c < NCP; ++c) {
                const int cls = g * NCP + c;
                if (cls < NC) {
#pragma unroll
                    for (int j = 0; j < KQ; ++j) {
                        const int q = tq + j * p.QTHR;
                        if (q < nquads) {
                            float4 v;
                            kf_unpack(acc2[c][j][0], v.x, v.y);
                            kf_unpack(acc2[c][j][1], v.z, v.w);
                            *reinterpret_cast<float4*>(mypart + (int64_t)cls * D + 4 * q) = v;
                        }
                    }
                }
            }
        }
        if (tid < NC) mypart[nW + tid] = dbacc;
        if (tid < 2 * T) {   // the epilogue's __syncthreads made lacc visible; half-warps in a fixed order
            float sum = 0.f;
#pragma unroll 4
            for (int w = 0; w < 2 * KF_WARPS; ++w) sum += lacc[w * 2 * T + tid];
            p.loss_part[(int64_t)blockIdx.x * 2 * T + tid] = sum;
        }
        if (f.cm_step != nullptr)
#pragma unroll 1
            for (int b = tid; b < n_hist; b += KF_THREADS) {
                const unsigned int h = hist[b];
                if (h != 0u) atomicAdd(f.cm_step + b, (unsigned long long)h);
            }
    }
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 7] = (unsigned long long)clock64();
    // (grid.sync() orders the partial stores: its master thread fences after the CTA barrier, which is cumulative)
    cg::this_grid().sync();
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 8] = (unsigned long long)clock64();

    // ---- after the barrier: the payload (reduce buffer, then the step confusion counts) in <= 148 slices ----
    const int G_ctas = gridDim.x;
    const int n_f32 = (int)tail0 + 2 * T, itail0 = (int)tail0, inW = (int)nW;   // NC * D < 2^31 on this path
    const int vf = (n_f32 + 3) / 4, vi = (int)((p.n_cm + 1) / 2), VT = vf + vi;
    int per = 0, nsl = 0;
    peer_slicing(VT, per, nsl);
    // This CTA owns slices blockIdx.x, blockIdx.x + G, ...: their vectors form one flat list, walked KF_THREADS / S at a
    // time.  S-way split of the CTA sum per vector (S a power of two, lanes of one vector adjacent): thread = (vector,
    // split); 8 .. 19 partials per thread: a small grid (few partials, many slices per CTA) takes S = 1, a full grid
    // (one slice per CTA, 148 partials) S = 8.
    const int n_my = (int)blockIdx.x < nsl ? (nsl - (int)blockIdx.x + G_ctas - 1) / G_ctas : 0;
    const int n_my_vec = n_my * per;
    int S = 1;
    while (S < 8 && 16 * S <= G_ctas) S *= 2;   // a function of the grid only: the summation order (hence every bit of
                                                // the result) does not depend on the mode or on which CTA owns a slice
    auto flat_vec = [&](int i, int& v, bool& on) {   // i-th vector of this CTA's list
        const int m = i / per, sl = (int)blockIdx.x + m * G_ctas;
        v = sl * per + (i - m * per);
        on = i < n_my_vec && v < min((sl + 1) * per, VT);
        if (!on) v = 0;
    };
    const int lv = tid / S, sp = tid - lv * S;
    const int cper = (G_ctas + S - 1) / S;
    const int c_lo = sp * cper, c_hi = min(G_ctas, c_lo + cper);
    const int vstep = KF_THREADS / S;
    const long long* cm_step_ll = reinterpret_cast<const long long*>(f.cm_step);

    // CTA-ordered sum of the float4 at element 4v of every partial (this thread's split), combined over the S lanes
    // of the vector by a fixed tree.  Called by whole warps; `ld` false = contribute zeros.  Elements past the end of
    // [dW | db] (padding of the partial buffers) are garbage here and are replaced by the caller.
    auto part_sum = [&](int v, bool ld) -> float4 {
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ld) {
            const float* base = p.part + 4 * (int64_t)v;
            for (int c0 = c_lo; c0 < c_hi; c0 += 20) {   // 20 partials in flight per thread, added in CTA order
                float4 x[20];
#pragma unroll
                for (int u = 0; u < 20; ++u)
                    if (c0 + u < c_hi) x[u] = __ldcg(reinterpret_cast<const float4*>(base + (int64_t)(c0 + u) * p.part_stride));
#pragma unroll
                for (int u = 0; u < 20; ++u)
                    if (c0 + u < c_hi) { sum.x += x[u].x; sum.y += x[u].y; sum.z += x[u].z; sum.w += x[u].w; }
            }
        }
        for (int o = 1; o < S; o <<= 1) {
            sum.x += __shfl_xor_sync(0xffffffffu, sum.x, o);
            sum.y += __shfl_xor_sync(0xffffffffu, sum.y, o);
            sum.z += __shfl_xor_sync(0xffffffffu, sum.z, o);
            sum.w += __shfl_xor_sync(0xffffffffu, sum.w, o);
        }
        return sum;
    };
    // [loss_sum | denom] over the CTAs, one warp per entry starting from the LAST warp (those usually hold no vector)
    auto local_tail = [&]() {
        for (int j = 0; j < 2 * T; ++j)
            if (warp == KF_WARPS - 1 - (j % KF_WARPS)) {
                const float sum = kf_warp_fixed_sum(p.loss_part + j, G_ctas, 2 * T, lane);
                if (lane == 0) tail_l[j] = sum;
            }
        __syncthreads();
    };
    // the local value of payload vector v as it travels / is finalized: partial sums, tail entries, or step counts
    auto local_vec = [&](int v, float4 a4) -> int4 {
        if (v >= vf) {
            const int j = 2 * (v - vf);
            longlong2 x = make_longlong2(0, 0);
            if (j < p.n_cm) x.x = __ldcg(cm_step_ll + j);
            if (j + 1 < p.n_cm) x.y = __ldcg(cm_step_ll + j + 1);
            return *reinterpret_cast<int4*>(&x);
        }
        float* a = reinterpret_cast<float*>(&a4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = 4 * v + k;
            if (e >= itail0) a[k] = e < n_f32 ? tail_l[e - itail0] : 0.f;
        }
        return *reinterpret_cast<int4*>(&a4);
    };
    // divide the [dW | db] elements of vector v by the denominator of their task (D % 4 == 0: a vector inside dW
    // never crosses a class row, so one integer division serves its four elements)
    auto finalize_vec = [&](int v, float4& a4, const float* tail) {
        float* a = reinterpret_cast<float*>(&a4);
        const int e0 = 4 * v;
        const int crow = e0 < inW ? (int)((unsigned)e0 / (unsigned)D) : 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = e0 + k;
            if (e < itail0) {
                const int c = (e < inW) ? crow : e - inW;
                const float dn = tail[T + task_of[c]];
                a[k] = dn > 0.f ? kf_div(a[k], dn) : 0.f;
            }
        }
    };
    auto store_vec = [&](int v, const float4& a4) {
        const float* a = reinterpret_cast<const float*>(&a4);
        const int i0 = 4 * v;
        if (i0 + 3 < n_f32) *reinterpret_cast<float4*>(p.reduce_buf + i0) = a4;
        else
            for (int k = 0; k < 4; ++k)
                if (i0 + k < n_f32) p.reduce_buf[i0 + k] = a[k];
    };
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 9] = (unsigned long long)clock64();

    if (p.mode != KF_MODE_PEER || p.peer.world == 1) {
        bool have_tail = false;
        {
            for (int ib = 0; ib < n_my_vec; ib += vstep) {      // whole warps iterate together (shuffles inside)
                int v;
                bool on;
                flat_vec(ib + lv, v, on);
                const float4 a4s = part_sum(v, on && 4 * v < itail0);   // loads first ...
                if (!have_tail) { local_tail(); have_tail = true; }   // ... the tail sums ride under them
                if (!on || sp != 0) continue;
                const int4 raw = local_vec(v, a4s);
                if (v < vf) {
                    float4 a4 = *reinterpret_cast<const float4*>(&raw);
                    if (p.mode != KF_MODE_SUMS) finalize_vec(v, a4, tail_l);
                    store_vec(v, a4);
                } else if (p.mode != KF_MODE_SUMS && p.cm_total != nullptr) {
                    const longlong2 x = *reinterpret_cast<const longlong2*>(&raw);
                    const int j = 2 * (v - vf);
                    long long* cs = reinterpret_cast<long long*>(f.cm_step);
                    if (j < p.n_cm) { p.cm_total[j] += x.x; cs[j] = 0; }
                    if (j + 1 < p.n_cm) { p.cm_total[j + 1] += x.y; cs[j + 1] = 0; }
                }
            }
        }
        if (blockIdx.x == 0 && tid == 0 && p.mode != KF_MODE_SUMS && p.out_loss != nullptr) {   // CTA 0 owns slice 0
            float total = 0.f;
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                const float l = tail_l[T + t] > 0.f ? kf_div(tail_l[t], tail_l[T + t]) : 0.f;
                p.out_loss[t] = l;
                total += l;
            }
            p.out_loss[T] = total;
        }
        if (p.timing != nullptr && tid == 0) {
            p.timing[blockIdx.x * KF_TIMING_SLOTS + 10] = (unsigned long long)clock64();
            p.timing[blockIdx.x * KF_TIMING_SLOTS + 11] = kf_globaltimer();
        }
        return;
    }

    // ================= exchange over NVLink peer memory (protocol of k4_peer.cu) =================
    const PeerLinks& L = p.peer;
    const int world = L.world, rank = L.rank;
    if (tid == 0) step_s = *reinterpret_cast<volatile unsigned int*>(L.ctl) + 1u;
    local_tail();   // (also the barrier that publishes step_s)
    const unsigned int step = step_s;
    const long long par = step & 1u;
    const long long my_slot = (par * world + rank) * L.slot_vecs;
    const int tail_vecs = (2 * T + 3) / 4;
    // push: every vector of this CTA's slices into the slot [parity][my rank] of every inbox (mine included), a copy of
    // the tail per slice, then the flags
    for (int ib = 0; ib < n_my_vec; ib += vstep) {
        int v;
        bool on;
        flat_vec(ib + lv, v, on);
        const float4 a4s = part_sum(v, on && 4 * v < itail0);
        if (!on || sp != 0) continue;
        const int4 raw = local_vec(v, a4s);
        for (int k = 0; k < world; ++k) {
            int q = rank + k;                   // own inbox first, peers in staggered order
            if (q >= world) q -= world;
            L.data[q][my_slot + v] = raw;
        }
    }
    for (int sl = blockIdx.x; sl < nsl; sl += G_ctas) {
        if (tid < tail_vecs) {
            float4 x;
            float* xf = reinterpret_cast<float*>(&x);
#pragma unroll
            for (int k = 0; k < 4; ++k) xf[k] = (4 * tid + k < 2 * T) ? tail_l[4 * tid + k] : 0.f;
            const long long tail_off = L.cap_vecs + (long long)sl * PEER_TAIL_VECS;
            for (int k = 1; k < world; ++k) {
                int q = rank + k;
                if (q >= world) q -= world;
                L.data[q][my_slot + tail_off + tid] = *reinterpret_cast<int4*>(&x);
            }
        }
    }
    __syncthreads();
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 12] = (unsigned long long)clock64();
    if (tid < world && tid != rank) peer_publish_many(L, tid, par, blockIdx.x, G_ctas, nsl, step);
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 13] = (unsigned long long)clock64();
    // wait for every peer's copy of all my slices (all pushes of this CTA are out: no wait can block a push), then the
    // rank-ordered sum + finalize over the same flat list
    if (tid < world && tid != rank) peer_wait_many(L, tid, par, blockIdx.x, G_ctas, nsl, step);
    __syncthreads();
    if (p.timing != nullptr && tid == 0) p.timing[blockIdx.x * KF_TIMING_SLOTS + 14] = (unsigned long long)clock64();
    if (n_my > 0) {   // global loss sums / denominators in rank order: identical in every CTA of every rank
        const long long tail_off = L.cap_vecs + (long long)blockIdx.x * PEER_TAIL_VECS;
        if (tid < 2 * T) {
            float sum = 0.f;
            for (int r = 0; r < world; ++r) {
                float x;
                if (r == rank) x = tail_l[tid];
                else x = __ldcg(reinterpret_cast<const float*>(L.data[rank] + (par * world + r) * L.slot_vecs + tail_off) + tid);
                sum = (r == 0) ? x : sum + x;
            }
            tail_g[tid] = sum;
        }
        __syncthreads();
        for (int i = tid; i < n_my_vec; i += KF_THREADS) {
            int v;
            bool on;
            flat_vec(i, v, on);
            if (!on) continue;
            if (v < vf) {
                float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = 0; r < world; ++r) {
                    const int4 raw = __ldcg(L.data[rank] + (par * world + r) * L.slot_vecs + v);
                    const float4 x = *reinterpret_cast<const float4*>(&raw);
                    if (r == 0) a4 = x;
                    else { a4.x += x.x; a4.y += x.y; a4.z += x.z; a4.w += x.w; }
                }
                finalize_vec(v, a4, tail_g);
                // tail elements of a straddling vector: the global (unnormalised) sums, as finalize leaves them
                float* a = reinterpret_cast<float*>(&a4);
                for (int k = 0; k < 4; ++k)
                    if (4 * v + k >= itail0 && 4 * v + k < n_f32) a[k] = tail_g[4 * v + k - itail0];
                store_vec(v, a4);
            } else {
                longlong2 a2 = make_longlong2(0, 0);
                for (int r = 0; r < world; ++r) {
                    const int4 raw = __ldcg(L.data[rank] + (par * world + r) * L.slot_vecs + v);
                    const longlong2 x = *reinterpret_cast<const longlong2*>(&raw);
                    a2.x += x.x;
                    a2.y += x.y;
                }
                const int j = 2 * (v - vf);
                long long* cs = reinterpret_cast<long long*>(f.cm_step);
                if (p.cm_total != nullptr) {
                    if (j < p.n_cm) { p.cm_total[j] += a2.x; cs[j] = 0; }
                    if (j + 1 < p.n_cm) { p.cm_total[j + 1] += a2.y; cs[j + 1] = 0; }
                }
            }
        }
    }
    // the last CTA to finish publishes the losses (NaN when a peer wait timed out) and advances the step counter
    __syncthreads();
    if (p.timing != nullptr && tid == 0) {
        p.timing[blockIdx.x * KF_TIMING_SLOTS + 10] = (unsigned long long)clock64();
        p.timing[blockIdx.x * KF_TIMING_SLOTS + 11] = kf_globaltimer();
    }
    if (tid == 0) {
        __threadfence();
        last_s = atomicAdd(L.ctl + 1, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (last_s && tid == 0) {
        if (p.out_loss != nullptr) {
            // (a CTA without slices never built tail_g and the last CTA may be one of those: sum slice 0's tails again)
            const long long tail_off = L.cap_vecs;   // slice 0 exists whenever the payload is not empty
            auto gsum = [&](int j) {
                float sum = 0.f;
                for (int r = 0; r < world; ++r) {
                    float x;
                    if (r == rank) x = tail_l[j];
                    else x = __ldcg(reinterpret_cast<const float*>(L.data[rank] + (par * world + r) * L.slot_vecs + tail_off) + j);
                    sum = (r == 0) ? x : sum + x;
                }
                return sum;
            };
            const bool bad = *reinterpret_cast<volatile unsigned int*>(L.ctl + 2) != 0u;
            const float qnan = __int_as_float(0x7fc00000);
            float total = 0.f;
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                const float ls = gsum(t), dn = gsum(T + t);
                const float l = dn > 0.f ? kf_div(ls, dn) : 0.f;
                p.out_loss[t] = bad ? qnan : l;
                total += l;
            }
            p.out_loss[T] = bad ? qnan : total;
        }
        L.ctl[1] = 0u;
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(L.ctl) = step;
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
struct KFPlan {
    int NCP, KQ, G, QTHR, tile_rows, rows_per_cta, grid, hist_smem, w_slot;
    size_t smem;
};

static bool kf_plan(int B, int D, int NC, int T, int es, long long n_cm, int n_sm, size_t smem_max, KFPlan& P) {
    if (B < 1 || D % 4 != 0 || ((size_t)D * es) % 16 != 0 || NC < 1) return false;
    const int nquads = D / 4;
    int best = -1, bNCP = 0, bKQ = 0, bG = 0, bQ = 0;
    const int ncps[4] = {16, 12, 8, 4};
    for (int i = 0; i < 4; ++i) {
        const int NCP = ncps[i];
        const int G = (NC + NCP - 1) / NCP;
        const int QTHR = (KF_THREADS / G) / 32 * 32;
        if (QTHR < 32) continue;
        const int KQ = (nquads + QTHR - 1) / QTHR;
        if (4 * NCP * KQ > KF_MAX_ACC) continue;   // instantiated: (16,1) (12,1) (8,1) (8,2) (4,1) (4,2) (4,3) (4,4)
        const int active = G * std::min(QTHR, (nquads + KQ - 1) / KQ);   // threads with work in the dW pass
        // prefer more active threads, then fewer wasted FMAs on padded classes
        const int score = active * 64 - (G * NCP - NC) * KQ * 16;
        if (score > best) { best = score; bNCP = NCP; bKQ = KQ; bG = G; bQ = QTHR; }
    }
    if (best < 0) return false;
    P.NCP = bNCP; P.KQ = bKQ; P.G = bG; P.QTHR = bQ;
    P.hist_smem = (n_cm > 0 && n_cm <= KF_MAX_SMEM_HIST) ? 1 : 0;
    const int hist_bins = P.hist_smem ? (int)n_cm : 0;
    int rpc = (B + n_sm - 1) / n_sm;
    // Small batches: fewer, fuller CTAs (three tiles each).  Fewer partials to sum and fewer flags to exchange; above all
    // the kernel then leaves most SMs to whatever runs next to it -- K1 of the next batch in a pipelined loop: with
    // 512 rows, 8 -> 24 rows per CTA costs 3 us when the kernel runs alone and saves 7.5 us when K1 overlaps it.
    int min_rows = 3 * KF_TILE_ROWS;
    if (const char* mr = getenv("NKBK_FUSED_MIN_ROWS")) {   // experiment knob (8 | 16 | 24 | 32)
        const int v = atoi(mr);
        if (v >= 1 && v <= 1024) min_rows = v;
    }
    if (rpc < min_rows) rpc = min_rows;
    P.rows_per_cta = rpc;
    P.grid = (B + rpc - 1) / rpc;
    P.tile_rows = 0;
    for (int tr = KF_TILE_ROWS; tr >= 4; tr >>= 1) {
        const KFSmem S = kf_smem(D, NC, T, bG * bNCP, tr, rpc, es, hist_bins);
        if (S.total + 2048 <= smem_max) { P.tile_rows = tr; P.smem = S.total; break; }   // static + driver-reserved
    }
    if (P.tile_rows == 0) return false;
    P.w_slot = ((size_t)NC * D * 4 >= (size_t)P.tile_rows * D * es) ? 1 : 0;
    return true;
}

struct KFDevice {
    int n_sm = 0, coop = 0;
    size_t smem_max = 0;
    bool ready = false;
    unsigned long long* timing = nullptr;   // NKBK_FUSED_TIMING=1: per-CTA phase stamps of the last launch
    int timing_grid = 0;
};
static KFDevice g_kfdev[64];

int64_t k2_fused_workspace_floats(int B, int D, int NC, int T) {
    // [148][round4(NC * D + NC)] partials + [148][2T] loss partials
    const int64_t stride = ((int64_t)NC * D + NC + 3) & ~int64_t(3);
    return PEER_MAX_SLICES * stride + (((int64_t)PEER_MAX_SLICES * 2 * T + 3) & ~int64_t(3));
}

template <typename ET, int NCP, int KQ>
static int kf_launch(const KFParams& p, const KFPlan& P, int dev, cudaStream_t st) {
    auto kern = k2_fused_step<ET, NCP, KQ>;
    static size_t set_smem[64] = {};
    if (P.smem > set_smem[dev]) {
        NKBK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.smem));
        set_smem[dev] = P.smem;
    }
    void* args[] = {const_cast<KFParams*>(&p)};
    NKBK_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(P.grid), dim3(KF_THREADS), args, P.smem, st));
    count_launch();
    return 1;
}

// Returns 1 when the fused kernel was launched, 0 when the shape / device does not qualify (caller takes the
// three-kernel path), < 0 on error.  `mode`: KF_MODE_SUMS leaves unnormalised sums in reduce_buf (nkbk_heads_step
// contract), KF_MODE_FINALIZE also applies nkbk_heads_finalize, KF_MODE_PEER also the K4' exchange.
int launch_k2_fused(const K2FwdParams& f, int emb_dtype, float* reduce_buf, float* ws_floats, float* out_loss,
                    int64_t* cm_total, int64_t n_cm, int mode, cudaStream_t st) {
    if (f.dlogits == nullptr) return 0;
    const char* off = getenv("NKBK_DISABLE_FUSED_HEADS");
    if ((off != nullptr && off[0] == '1') || !heads_one_launch()) return 0;
    int dev = 0;
    NKBK_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return 0;
    KFDevice& dv = g_kfdev[dev];
    if (!dv.ready) {
        int v = 0;
        NKBK_CHECK_CUDA(cudaDeviceGetAttribute(&dv.n_sm, cudaDevAttrMultiProcessorCount, dev));
        NKBK_CHECK_CUDA(cudaDeviceGetAttribute(&dv.coop, cudaDevAttrCooperativeLaunch, dev));
        NKBK_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        dv.smem_max = (size_t)v;
        dv.ready = true;
    }
    if (!dv.coop || dv.n_sm < 1) return 0;
    const int es = emb_dtype == NKBK_F32 ? 4 : 2;
    const int T = f.seg.T;
    if ((reinterpret_cast<uintptr_t>(f.emb) & 15) != 0 || (reinterpret_cast<uintptr_t>(f.W) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(reduce_buf) & 15) != 0 || (reinterpret_cast<uintptr_t>(ws_floats) & 15) != 0)
        return 0;
    KFPlan P;
    const int n_sm = std::min(dv.n_sm, PEER_MAX_SLICES);
    long long n_bins = 0;
    if (f.cm_step != nullptr)
        for (int t = 0; t < T; ++t) {
            const long long C = f.seg.off[t + 1] - f.seg.off[t];
            n_bins += C * C;
        }
    if (!kf_plan(f.B, f.D, f.NC, T, es, n_bins, n_sm, dv.smem_max, P)) return 0;

    KFParams p;
    memset(&p, 0, sizeof(p));
    p.f = f;
    p.reduce_buf = reduce_buf;
    p.part_stride = ((int64_t)f.NC * f.D + f.NC + 3) & ~int64_t(3);
    p.part = ws_floats;
    p.loss_part = ws_floats + PEER_MAX_SLICES * p.part_stride;
    p.out_loss = out_loss;
    p.cm_total = reinterpret_cast<long long*>(cm_total);
    p.n_cm = (f.cm_step && mode != KF_MODE_SUMS) ? n_cm : 0;
    p.n_bins = n_bins;
    p.rows_per_cta = P.rows_per_cta; p.tile_rows = P.tile_rows; p.mode = mode;
    p.G = P.G; p.QTHR = P.QTHR; p.NCD = P.G * P.NCP; p.hist_smem = P.hist_smem; p.w_slot = P.w_slot;
    if (mode == KF_MODE_PEER) {
        const long long n_f32 = (long long)f.NC * f.D + f.NC + 2LL * T;
        const int rc = peer_links(p.peer, n_f32, p.n_cm, "nkbk_heads_train_step");
        if (rc) return rc;
    }
    {
        const char* tm = getenv("NKBK_FUSED_TIMING");
        if (tm != nullptr && tm[0] == '1') {
            if (dv.timing == nullptr)
                NKBK_CHECK_CUDA(cudaMalloc(&dv.timing, sizeof(unsigned long long) * KF_TIMING_SLOTS * PEER_MAX_SLICES));
            p.timing = dv.timing;
            dv.timing_grid = P.grid;
        }
    }
#define KF_CASE(NCPV, KQV)                                                                                     \
    if (P.NCP == NCPV && P.KQ == KQV)                                                                          \
        return emb_dtype == NKBK_F32 ? kf_launch<float, NCPV, KQV>(p, P, dev, st)                                    \
                                     : kf_launch<__nv_bfloat16, NCPV, KQV>(p, P, dev, st)
    KF_CASE(16, 1); KF_CASE(12, 1); KF_CASE(8, 1); KF_CASE(8, 2);
    KF_CASE(4, 1); KF_CASE(4, 2); KF_CASE(4, 3); KF_CASE(4, 4);
#undef KF_CASE
    return 0;
}

}  // namespace nkbk

// Debug helper (host-synchronous): the per-CTA phase stamps of the last fused launch made with NKBK_FUSED_TIMING=1.
// out_host: [n_ctas][16] = {globaltimer ns at entry, clock64 at: entry, tables done, weights in, pass 1 done, epilogue
// done, pass 2 done, partials written, grid barrier passed, reduce set up, kernel work done, globaltimer ns at exit,
// then (exchange over peer memory only) clock64 at: pushes done, flags published, first slice's peers arrived, -}.
// Returns the number of CTAs written (0 when no timed launch happened on the current device).
extern "C" int nkbk_debug_fused_timing(uint64_t* out_host, int max_ctas) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
    nkbk::KFDevice& dv = nkbk::g_kfdev[dev];
    if (dv.timing == nullptr || out_host == nullptr) return 0;
    const int n = dv.timing_grid < max_ctas ? dv.timing_grid : max_ctas;
    if (cudaMemcpy(out_host, dv.timing, sizeof(unsigned long long) * nkbk::KF_TIMING_SLOTS * n, cudaMemcpyDeviceToHost) != cudaSuccess)
        return 0;
    return n;
}
