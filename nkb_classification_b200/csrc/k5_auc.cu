// K5: exact one-vs-rest ROC-AUC statistics for every class column of every task, on the device
// (the step right after the path -- replaces the epoch lists + sklearn roc_auc_score of
// nkb_classification/metrics.py:33-42 for the part that needs all N probabilities).
//
// roc_auc_score (trapezoid over the ROC curve with tied scores merged) equals the Mann-Whitney statistic
//     AUC_c = ( #{(i, j): y_i = c, y_j != c, s_i > s_j}  +  1/2 #{... s_i == s_j} ) / (P_c * N_c)
// so the kernel returns, per column, the INTEGERS  num2 = 2 * #greater + #equal,  P,  N  and the host divides
// once in float64:  AUC = num2 / (2 P N).  Integer, order independent, exact -- no sort, no float accumulation.
//
//   k5_split   one pass over the [N, ld] fp32 probabilities: per column, positives' and negatives' scores are
//              compacted into two dense lists (warp-aggregated atomics; order is irrelevant to the counts)
//   k5_pairs   tiles of 1024 positives (4 per thread, in registers) x 4096 negatives (streamed through shared
//              memory, broadcast reads): 2*(p > n) + (p == n) summed in integers, one 64-bit atomic per CTA
// Work is P*N per column instead of a sort's N log N, but it is ~3 instructions per pair on 148 SMs: a
// 10-class epoch of 10^6 samples is ~10^12 pairs, tens of milliseconds, once per epoch.
#include "k2_common.cuh"

namespace nkbk {

constexpr int K5_THREADS = 256;
constexpr int K5_PPT = 4;                          // positives per thread
constexpr int K5_POS_TILE = K5_THREADS * K5_PPT;   // 1024
constexpr int K5_NEG_TILE = 1024;                  // negatives per shared-memory stage
constexpr int K5_NEG_CHUNK = 4096;                 // negatives per work item

__global__ void __launch_bounds__(K5_THREADS) k5_split(const float* __restrict__ probs, long long N, int ld, K2Seg seg,
                                                       const int64_t* __restrict__ labels, float* __restrict__ pos,
                                                       float* __restrict__ neg, unsigned int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int T = seg.T, NC = seg.off[T];
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long base = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < N; base += nwarps * 32) {
        const long long row = base + lane;
        const bool valid = row < N;
        int t = 0;
        long long y = valid ? labels[row * T] : -1;
        for (int c = 0; c < NC; ++c) {
            if (c >= seg.off[t + 1]) {
                ++t;
                y = valid ? labels[row * T + t] : -1;
            }
            const float s = valid ? __ldg(probs + row * ld + c) : 0.f;
            const bool is_pos = valid && y == (long long)(c - seg.off[t]);
            const unsigned mp = __ballot_sync(0xffffffffu, is_pos);
            const unsigned mn = __ballot_sync(0xffffffffu, valid && !is_pos);
            unsigned bp = 0, bn = 0;
            if (lane == 0) {
                if (mp) bp = atomicAdd(counts + 2 * c, __popc(mp));
                if (mn) bn = atomicAdd(counts + 2 * c + 1, __popc(mn));
            }
            bp = __shfl_sync(0xffffffffu, bp, 0);
            bn = __shfl_sync(0xffffffffu, bn, 0);
            const unsigned lt = (1u << lane) - 1u;
            if (is_pos) pos[(long long)c * N + bp + __popc(mp & lt)] = s;
            else if (valid) neg[(long long)c * N + bn + __popc(mn & lt)] = s;
        }
    }
}

__global__ void __launch_bounds__(K5_THREADS) k5_pairs(const float* __restrict__ pos, const float* __restrict__ neg,
                                                       const unsigned int* __restrict__ counts, long long N,
                                                       long long* __restrict__ out) {
    __shared__ float ns[K5_NEG_TILE];
    __shared__ unsigned long long red[K5_THREADS / 32];
    const int c = blockIdx.y;
    const long long P = counts[2 * c], Q = counts[2 * c + 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[3 * c + 1] = P; out[3 * c + 2] = Q; }
    const long long ptiles = (P + K5_POS_TILE - 1) / K5_POS_TILE, nchunks = (Q + K5_NEG_CHUNK - 1) / K5_NEG_CHUNK;
    const float* pc = pos + (long long)c * N;
    const float* nc = neg + (long long)c * N;
    unsigned long long total = 0;
    for (long long item = blockIdx.x; item < ptiles * nchunks; item += gridDim.x) {
        const long long pt = item / nchunks, ch = item - pt * nchunks;
        float p[K5_PPT];
#pragma unroll
        for (int k = 0; k < K5_PPT; ++k) {
            const long long i = pt * K5_POS_TILE + (long long)k * K5_THREADS + threadIdx.x;
            p[k] = i < P ? __ldg(pc + i) : -INFINITY;     // padding never compares greater or equal
        }
        unsigned int cnt = 0;                               // <= 4 * 4096 * 2 per item
        for (int st = 0; st < K5_NEG_CHUNK / K5_NEG_TILE; ++st) {
            const long long j0 = ch * K5_NEG_CHUNK + (long long)st * K5_NEG_TILE;
            if (j0 >= Q) break;                              // block-uniform
            __syncthreads();
            for (int j = threadIdx.x; j < K5_NEG_TILE; j += K5_THREADS)
                ns[j] = (j0 + j < Q) ? __ldg(nc + j0 + j) : INFINITY;
            __syncthreads();
#pragma unroll 8
            for (int j = 0; j < K5_NEG_TILE; ++j) {
                const float n = ns[j];
#pragma unroll
                for (int k = 0; k < K5_PPT; ++k) cnt += (p[k] > n ? 2u : 0u) + (p[k] == n ? 1u : 0u);
            }
        }
        total += cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = total;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < K5_THREADS / 32; ++w) s += red[w];
        if (s) atomicAdd(reinterpret_cast<unsigned long long*>(out + 3 * c), s);
    }
}

}  // namespace nkbk

using namespace nkbk;

extern "C" int64_t nkbk_auc_workspace_bytes(int64_t N, int NC) {
    if (N < 1) N = 1;
    if (NC < 1) NC = 1;
    return 2 * (int64_t)NC * N * (int64_t)sizeof(float) + 2 * (int64_t)NC * (int64_t)sizeof(unsigned int) + 256;
}

extern "C" int nkbk_roc_auc_counts(const float* probs, int64_t N, int ld, const int32_t* seg_offsets, int T,
                                   const int64_t* labels, int64_t* out_counts, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    NKBK_CHECK_ARG(seg_offsets && T >= 1 && T <= K2_MAX_TASKS, "nkbk_roc_auc_counts: bad seg_offsets / T=%d", T);
    K2Seg seg;
    seg.T = T;
    for (int t = 0; t <= T; ++t) seg.off[t] = seg_offsets[t];
    NKBK_CHECK_ARG(seg.off[0] == 0, "nkbk_roc_auc_counts: seg_offsets[0] != 0");
    for (int t = 0; t < T; ++t) NKBK_CHECK_ARG(seg.off[t + 1] > seg.off[t], "nkbk_roc_auc_counts: task %d has no classes", t);
    const int NC = seg.off[T];
    if (NC > K2_MAX_NC) { set_error("nkbk_roc_auc_counts: %d classes > %d", NC, K2_MAX_NC); return NKBK_E_SHAPE; }
    NKBK_CHECK_ARG(N >= 0 && N < (int64_t(1) << 31) && ld >= NC, "nkbk_roc_auc_counts: N=%lld ld=%d NC=%d", (long long)N, ld, NC);
    NKBK_CHECK_ARG(out_counts != nullptr, "nkbk_roc_auc_counts: NULL out_counts");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    NKBK_CHECK_CUDA(cudaMemsetAsync(out_counts, 0, (size_t)NC * 3 * sizeof(int64_t), st));
    if (N == 0) return NKBK_OK;
    NKBK_CHECK_ARG(probs && labels && workspace, "nkbk_roc_auc_counts: NULL probs / labels / workspace");
    if ((int64_t)workspace_bytes < nkbk_auc_workspace_bytes(N, NC)) {
        set_error("nkbk_roc_auc_counts: workspace %zu < %lld bytes", workspace_bytes, (long long)nkbk_auc_workspace_bytes(N, NC));
        return NKBK_E_ARG;
    }
    float* pos = static_cast<float*>(workspace);
    float* neg = pos + (int64_t)NC * N;
    unsigned int* counts = reinterpret_cast<unsigned int*>(neg + (int64_t)NC * N);
    NKBK_CHECK_CUDA(cudaMemsetAsync(counts, 0, 2 * (size_t)NC * sizeof(unsigned int), st));
    int blocks = (int)((N + K5_THREADS - 1) / K5_THREADS);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k5_split<<<blocks, K5_THREADS, 0, st>>>(probs, (long long)N, ld, seg, labels, pos, neg, counts);
    NKBK_CHECK_LAUNCH("k5_split");
    // enough CTAs to fill the machine for every column; each strides over that column's (pos tile, neg chunk) items
    const long long worst = ((N + K5_POS_TILE - 1) / K5_POS_TILE) * ((N + K5_NEG_CHUNK - 1) / K5_NEG_CHUNK) / 4 + 1;
    int gx = (int)(worst < 148 * 4 ? worst : 148 * 4);
    dim3 grid((unsigned)gx, (unsigned)NC);
    k5_pairs<<<grid, K5_THREADS, 0, st>>>(pos, neg, counts, (long long)N, reinterpret_cast<long long*>(out_counts));
    NKBK_CHECK_LAUNCH("k5_pairs");
    return NKBK_OK;
}
