// K3: per-task argmax + confusion-matrix accumulation (integer, exact).
// One thread per (row, task); counts are privatised in a shared-memory
// histogram per CTA and flushed with one 64-bit atomic per non-zero bin, so
// global atomics scale with CTAs * bins, not with rows.
#include "nkbk_common.cuh"

namespace nkbk {

constexpr int K3_MAX_TASKS = 64;
constexpr int K3_SMEM_BINS = 8192;  // 32 KB of uint32 counters
constexpr int K3_THREADS = 256;

struct K3Seg {
    int T;
    int off[K3_MAX_TASKS + 1];     // class offsets into a logits row
    int cm_off[K3_MAX_TASKS + 1];  // offsets of each task's C_t x C_t matrix
};

template <typename T>
__device__ __forceinline__ float ld_logit(const T* p);
template <>
__device__ __forceinline__ float ld_logit<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld_logit<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
}

template <typename LT, bool SMEM_HIST>
__global__ void __launch_bounds__(K3_THREADS) k3_argmax_confusion(const LT* __restrict__ logits, int B, int ld,
                                                                 const K3Seg seg, const int64_t* __restrict__ labels,
                                                                 int32_t* __restrict__ out_pred,
                                                                 unsigned long long* __restrict__ cm) {
    __shared__ uint32_t hist[SMEM_HIST ? K3_SMEM_BINS : 1];
    const int nbins = seg.cm_off[seg.T];
    const bool do_cm = cm != nullptr && labels != nullptr;
    if (SMEM_HIST && do_cm) {
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0u;
        __syncthreads();
    }
    const int64_t total = (int64_t)B * seg.T;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(idx / seg.T);
        const int t = (int)(idx - (int64_t)row * seg.T);
        const int c0 = seg.off[t], C = seg.off[t + 1] - c0;
        const LT* z = logits + (int64_t)row * ld + c0;
        // torch.argmax: first maximal element, NaN counts as maximal
        float best = ld_logit<LT>(z);
        int bi = 0;
        for (int j = 1; j < C; ++j) {
            const float v = ld_logit<LT>(z + j);
            if (!(best != best) && (v > best || v != v)) { best = v; bi = j; }
        }
        if (out_pred) out_pred[idx] = bi;
        if (do_cm) {
            const int64_t y = labels[idx];
            if (y >= 0 && y < C) {
                const int bin = seg.cm_off[t] + (int)y * C + bi;
                if (SMEM_HIST) atomicAdd(&hist[bin], 1u);
                else atomicAdd(&cm[bin], 1ull);
            }
        }
    }
    if (SMEM_HIST && do_cm) {
        __syncthreads();
        for (int i = threadIdx.x; i < nbins; i += blockDim.x) {
            const uint32_t v = hist[i];
            if (v) atomicAdd(&cm[i], (unsigned long long)v);
        }
    }
}

}  // namespace nkbk

using namespace nkbk;

extern "C" int nkbk_argmax_confusion(const void* logits, int dtype, int B, int ld, const int32_t* seg_offsets, int T,
                                     const int64_t* labels, int32_t* out_pred, int64_t* cm, void* stream) {
    NKBK_CHECK_ARG(B >= 0, "nkbk_argmax_confusion: B=%d", B);
    NKBK_CHECK_ARG(seg_offsets && T >= 1, "nkbk_argmax_confusion: NULL seg_offsets or T=%d", T);
    if (T > K3_MAX_TASKS) {
        set_error("nkbk_argmax_confusion: T=%d > %d tasks", T, K3_MAX_TASKS);
        return NKBK_E_SHAPE;
    }
    NKBK_CHECK_ARG(dtype == NKBK_F32 || dtype == NKBK_BF16, "nkbk_argmax_confusion: dtype=%d", dtype);
    K3Seg seg;
    seg.T = T;
    seg.cm_off[0] = 0;
    for (int t = 0; t <= T; ++t) seg.off[t] = seg_offsets[t];
    NKBK_CHECK_ARG(seg.off[0] == 0, "nkbk_argmax_confusion: seg_offsets[0] != 0");
    for (int t = 0; t < T; ++t) {
        const int C = seg.off[t + 1] - seg.off[t];
        NKBK_CHECK_ARG(C >= 1, "nkbk_argmax_confusion: task %d has %d classes", t, C);
        const int64_t next = (int64_t)seg.cm_off[t] + (int64_t)C * C;
        if (next > (1 << 30)) { set_error("nkbk_argmax_confusion: confusion matrices too large"); return NKBK_E_SHAPE; }
        seg.cm_off[t + 1] = (int)next;
    }
    NKBK_CHECK_ARG(ld >= seg.off[T], "nkbk_argmax_confusion: ld=%d < %d classes", ld, seg.off[T]);
    if (B == 0) return NKBK_OK;
    NKBK_CHECK_ARG(logits != nullptr, "nkbk_argmax_confusion: NULL logits");
    NKBK_CHECK_ARG(out_pred || (cm && labels), "nkbk_argmax_confusion: nothing to produce");

    const int64_t total = (int64_t)B * T;
    int blocks = (int)((total + K3_THREADS - 1) / K3_THREADS);
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* cmu = reinterpret_cast<unsigned long long*>(cm);
    const bool smem = seg.cm_off[T] <= K3_SMEM_BINS;
    if (dtype == NKBK_F32) {
        const float* z = static_cast<const float*>(logits);
        if (smem) k3_argmax_confusion<float, true><<<blocks, K3_THREADS, 0, st>>>(z, B, ld, seg, labels, out_pred, cmu);
        else k3_argmax_confusion<float, false><<<blocks, K3_THREADS, 0, st>>>(z, B, ld, seg, labels, out_pred, cmu);
    } else {
        const __nv_bfloat16* z = static_cast<const __nv_bfloat16*>(logits);
        if (smem)
            k3_argmax_confusion<__nv_bfloat16, true><<<blocks, K3_THREADS, 0, st>>>(z, B, ld, seg, labels, out_pred, cmu);
        else
            k3_argmax_confusion<__nv_bfloat16, false><<<blocks, K3_THREADS, 0, st>>>(z, B, ld, seg, labels, out_pred, cmu);
    }
    NKBK_CHECK_LAUNCH("k3_argmax_confusion");
    return NKBK_OK;
}
