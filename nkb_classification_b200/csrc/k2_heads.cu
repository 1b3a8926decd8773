// K2: all task heads as one segmented GEMM over the pooled features, fused
// with per-task softmax, CE / focal loss and the logit gradient; then the
// weight / bias gradient as a second segmented GEMM (dlogits^T x emb).
// See include/nkbk.h for the contract, DESIGN.md "K2" for the reasoning.
//
// This file is the exact-fp32 path (FFMA accumulate; TF32/bf16 tensor cores
// would break the 1e-5 fp32 parity bar).  bf16 embeddings are widened on load.
//
//   k2_heads_forward   warp = 4 rows x 16 classes per pass, lanes split K with
//                      128-bit loads; a 31-shuffle transposing reduction leaves
//                      logit (row, class) = lane; per-(row, task) lanes then do
//                      softmax / loss / dlogits in registers.
//   k2_heads_dw        thread = one (or two, bf16) embedding columns, all NC
//                      accumulators in registers, dlogits chunk broadcast from
//                      shared memory; per-row-chunk partials (deterministic).
//   k2_heads_reduce    fixed-order sum of the partials into the reduce buffer.
//   k2_heads_finalize  divide by the (all-reduced) denominators, emit losses.
//   k2_heads_demb      d(loss)/d(emb) for an unfrozen backbone.
#include "nkbk_common.cuh"

namespace nkbk {

constexpr int K2_MAX_TASKS = 64;
constexpr int K2_MAX_NC = 1024;
constexpr int K2_FWD_WARPS = 4;
constexpr int K2_FWD_ROWS = 4;    // rows per warp
constexpr int K2_FWD_NCB = 16;    // classes per pass
constexpr int K2_DW_THREADS = 128;
constexpr int K2_DW_ROWS = 128;   // rows per dW chunk

struct K2Seg {
    int T;
    int off[K2_MAX_TASKS + 1];
};

struct K2Layout {  // workspace carve-up, in floats
    int fwd_blocks, fwd_warps_total;
    int dw_chunks;
    int64_t loss_part;  // [fwd_warps_total][2T]
    int64_t dw_part;    // [dw_chunks][NC][D]
    int64_t db_part;    // [dw_chunks][NC]
    int64_t total;
};

static K2Layout k2_layout(int B, int D, int NC, int T) {
    K2Layout L;
    const int rows_per_block = K2_FWD_WARPS * K2_FWD_ROWS;
    L.fwd_blocks = (B + rows_per_block - 1) / rows_per_block;
    L.fwd_warps_total = L.fwd_blocks * K2_FWD_WARPS;
    L.dw_chunks = (B + K2_DW_ROWS - 1) / K2_DW_ROWS;
    L.loss_part = 0;
    L.dw_part = L.loss_part + (int64_t)L.fwd_warps_total * 2 * T;
    L.dw_part = (L.dw_part + 3) & ~int64_t(3);
    L.db_part = L.dw_part + (int64_t)L.dw_chunks * NC * D;
    L.total = L.db_part + (int64_t)L.dw_chunks * NC;
    return L;
}

// ---- 4-element vector loads of an embedding / weight row -----------------
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
}

// Transposing warp reduction: on entry every lane holds 32 partial sums v[0..31];
// on exit lane L holds in v[0] the warp-wide total of partial sum number L.
__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
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

struct K2FwdParams {
    const void* emb;
    const float* W;
    const float* bias;
    const int64_t* labels;
    const float* class_weight;
    float* out_logits;
    float* out_probs;
    float* dlogits;
    float* loss_part;  // [warps_total][2T]
    int B, D, NC;
    int loss_kind;
    float gamma;
    int64_t ignore_index;
    K2Seg seg;
};

template <typename ET>
__global__ void __launch_bounds__(K2_FWD_WARPS * 32) k2_heads_forward(const K2FwdParams p) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.seg.T, NC = p.NC, D = p.D;
    float* zs = smem + (size_t)warp * (K2_FWD_ROWS * NC + K2_FWD_ROWS * 2 * T);  // [ROWS][NC] logits
    float* ls = zs + K2_FWD_ROWS * NC;                                          // [ROWS][2T] loss / denom terms
    const int row0 = (blockIdx.x * K2_FWD_WARPS + warp) * K2_FWD_ROWS;
    const ET* emb = static_cast<const ET*>(p.emb);

    // ---- logits: passes of 16 classes, K split across lanes ----
    for (int cb = 0; cb < NC; cb += K2_FWD_NCB) {
        float acc[K2_FWD_ROWS * K2_FWD_NCB];
#pragma unroll
        for (int i = 0; i < K2_FWD_ROWS * K2_FWD_NCB; ++i) acc[i] = 0.f;
        for (int k = lane * 4; k < D; k += 128) {
            float4 e[K2_FWD_ROWS];
#pragma unroll
            for (int r = 0; r < K2_FWD_ROWS; ++r) {
                const int row = min(row0 + r, p.B - 1);  // tail rows recompute the last row, never stored
                e[r] = ld4(emb + (int64_t)row * D + k);
            }
#pragma unroll
            for (int c = 0; c < K2_FWD_NCB; ++c) {
                const int cls = min(cb + c, NC - 1);
                const float4 w = ld4(p.W + (int64_t)cls * D + k);
#pragma unroll
                for (int r = 0; r < K2_FWD_ROWS; ++r) {
                    float a = acc[r * K2_FWD_NCB + c];
                    a = fmaf(e[r].x, w.x, a);
                    a = fmaf(e[r].y, w.y, a);
                    a = fmaf(e[r].z, w.z, a);
                    a = fmaf(e[r].w, w.w, a);
                    acc[r * K2_FWD_NCB + c] = a;
                }
            }
        }
        // 64 partial sums -> lane L owns entries L and 32+L  (entry = r*16 + c)
        float lo[32], hi[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { lo[i] = acc[i]; hi[i] = acc[32 + i]; }
        const float s0 = warp_reduce_scatter32(lo, lane);
        const float s1 = warp_reduce_scatter32(hi, lane);
        const int c = cb + (lane & 15);
        if (c < NC) {
            const float b = __ldg(p.bias + c);
            zs[(lane >> 4) * NC + c] = s0 + b;
            zs[(2 + (lane >> 4)) * NC + c] = s1 + b;
        }
    }
    __syncwarp();

    // ---- per (row, task): softmax, loss term, dlogits ----
    for (int i = lane; i < K2_FWD_ROWS * 2 * T; i += 32) ls[i] = 0.f;
    __syncwarp();
    for (int idx = lane; idx < K2_FWD_ROWS * T; idx += 32) {
        const int r = idx / T, t = idx - r * T;
        const int row = row0 + r;
        if (row >= p.B) continue;
        const int c0 = p.seg.off[t], C = p.seg.off[t + 1] - c0;
        const float* z = zs + r * NC + c0;
        float mx = z[0];
        for (int j = 1; j < C; ++j) mx = fmaxf(mx, z[j]);
        float se = 0.f;
        for (int j = 0; j < C; ++j) se += expf(z[j] - mx);
        const float lse = mx + logf(se);
        const int64_t y = p.labels ? p.labels[(int64_t)row * T + t] : p.ignore_index;
        const bool keep = p.labels != nullptr && y != p.ignore_index && y >= 0 && y < C;
        float q = 0.f;  // dlogit_j = q * (delta_jy - p_j)
        if (keep) {
            const float logpt = z[y] - lse;
            const float a = p.class_weight ? __ldg(p.class_weight + c0 + (int)y) : 1.f;
            float loss_i, den_i;
            if (p.loss_kind == NKBK_LOSS_FOCAL) {
                const float pt = expf(logpt);
                const float om = 1.f - pt;
                const float g = p.gamma;
                float ft, dterm;  // ft = om^g ; dterm = g * pt * om^(g-1) * logpt
                if (g == 0.f) { ft = 1.f; dterm = 0.f; }
                else {
                    const float pw1 = (g == 1.f) ? 1.f : ((g == 2.f) ? om : powf(om, g - 1.f));
                    ft = pw1 * om;
                    dterm = g * pt * pw1 * logpt;
                }
                loss_i = -a * ft * logpt;
                q = a * (dterm - ft);
                den_i = 1.f;
            } else {
                loss_i = -a * logpt;
                q = -a;
                den_i = a;
            }
            ls[r * 2 * T + t] = loss_i;
            ls[r * 2 * T + T + t] = den_i;
        }
        float* zo = p.out_logits ? p.out_logits + (int64_t)row * NC + c0 : nullptr;
        float* po = p.out_probs ? p.out_probs + (int64_t)row * NC + c0 : nullptr;
        float* go = p.dlogits ? p.dlogits + (int64_t)row * NC + c0 : nullptr;
        for (int j = 0; j < C; ++j) {
            const float zj = z[j];
            const float pj = expf(zj - lse);
            if (zo) zo[j] = zj;
            if (po) po[j] = pj;
            if (go) go[j] = keep ? q * ((j == (int)y ? 1.f : 0.f) - pj) : 0.f;
        }
    }
    __syncwarp();
    // fixed-order per-warp partial: sum over this warp's rows
    float* part = p.loss_part + (int64_t)(blockIdx.x * K2_FWD_WARPS + warp) * 2 * T;
    for (int i = lane; i < 2 * T; i += 32) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < K2_FWD_ROWS; ++r) s += ls[r * 2 * T + i];
        part[i] = s;
    }
}

// ---- dW / db partials ------------------------------------------------------
// VEC = embedding columns per thread (1 for fp32, 2 for bf16 pairs)
template <typename ET, int NCP>
__global__ void __launch_bounds__(K2_DW_THREADS) k2_heads_dw(const ET* __restrict__ emb,
                                                            const float* __restrict__ dlogits, int B, int D, int NC,
                                                            int cls0, float* __restrict__ dw_part,
                                                            float* __restrict__ db_part) {
    __shared__ __align__(16) float dl[K2_DW_ROWS][NCP];
    const int chunk = blockIdx.y;
    const int rbeg = chunk * K2_DW_ROWS, rows = min(K2_DW_ROWS, B - rbeg);
    const int ncls = min(NCP, NC - cls0);
    for (int i = threadIdx.x; i < K2_DW_ROWS * NCP; i += blockDim.x) {
        const int r = i / NCP, c = i - r * NCP;
        dl[r][c] = (r < rows && c < ncls) ? __ldg(dlogits + (int64_t)(rbeg + r) * NC + cls0 + c) : 0.f;
    }
    __syncthreads();
    const int k = blockIdx.x * K2_DW_THREADS + threadIdx.x;
    if (k < D) {
        float acc[NCP];
#pragma unroll
        for (int c = 0; c < NCP; ++c) acc[c] = 0.f;
        const ET* ep = emb + (int64_t)rbeg * D + k;
#pragma unroll 4
        for (int r = 0; r < rows; ++r) {
            const float e = load_as_float(ep + (int64_t)r * D);
#pragma unroll
            for (int c4 = 0; c4 < NCP; c4 += 4) {
                const float4 g = *reinterpret_cast<const float4*>(&dl[r][c4]);
                acc[c4 + 0] = fmaf(g.x, e, acc[c4 + 0]);
                acc[c4 + 1] = fmaf(g.y, e, acc[c4 + 1]);
                acc[c4 + 2] = fmaf(g.z, e, acc[c4 + 2]);
                acc[c4 + 3] = fmaf(g.w, e, acc[c4 + 3]);
            }
        }
        float* o = dw_part + ((int64_t)chunk * NC + cls0) * D + k;
#pragma unroll
        for (int c = 0; c < NCP; ++c)
            if (c < ncls) o[(int64_t)c * D] = acc[c];
    }
    if (blockIdx.x == 0 && threadIdx.x < ncls) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += dl[r][threadIdx.x];
        db_part[(int64_t)chunk * NC + cls0 + threadIdx.x] = s;
    }
}

// reduce_buf = [dW NC*D | db NC | loss_sum T | denom T]
__global__ void __launch_bounds__(256) k2_heads_reduce(const float* __restrict__ dw_part,
                                                       const float* __restrict__ db_part,
                                                       const float* __restrict__ loss_part, int chunks,
                                                       int warps_total, int NC, int D, int T, int have_grads,
                                                       float* __restrict__ reduce_buf) {
    const int64_t nW = (int64_t)NC * D;
    const int64_t n = nW + NC + 2 * T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        if (i < nW) {
            if (have_grads)
                for (int ch = 0; ch < chunks; ++ch) s += dw_part[(int64_t)ch * nW + i];
        } else if (i < nW + NC) {
            if (have_grads)
                for (int ch = 0; ch < chunks; ++ch) s += db_part[(int64_t)ch * NC + (i - nW)];
        } else {
            const int j = (int)(i - nW - NC);
            for (int w = 0; w < warps_total; ++w) s += loss_part[(int64_t)w * 2 * T + j];
        }
        reduce_buf[i] = s;
    }
}

__global__ void __launch_bounds__(256) k2_heads_finalize(float* __restrict__ reduce_buf, int NC, int D, const K2Seg seg,
                                                         float* __restrict__ out_loss, int64_t* __restrict__ cm_total,
                                                         int64_t* __restrict__ cm_step, int64_t n_cm) {
    const int T = seg.T;
    const int64_t nW = (int64_t)NC * D;
    const float* loss_sum = reduce_buf + nW + NC;
    const float* denom = loss_sum + T;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nW + NC; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (i < nW) ? (int)(i / D) : (int)(i - nW);
        int t = 0;
        while (t + 1 < T && c >= seg.off[t + 1]) ++t;
        const float dn = denom[t];
        reduce_buf[i] = dn > 0.f ? __fdiv_rn(reduce_buf[i], dn) : 0.f;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cm; i += (int64_t)gridDim.x * blockDim.x) {
        cm_total[i] += cm_step[i];
        cm_step[i] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && out_loss != nullptr) {
        float total = 0.f;
        for (int t = 0; t < T; ++t) {
            const float l = denom[t] > 0.f ? __fdiv_rn(loss_sum[t], denom[t]) : 0.f;
            out_loss[t] = l;
            total += l;
        }
        out_loss[T] = total;
    }
}

template <typename OT>
__device__ __forceinline__ void st4(OT* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}

template <typename OT>
__global__ void __launch_bounds__(256) k2_heads_demb(const float* __restrict__ dlogits, const float* __restrict__ denom,
                                                     const float* __restrict__ W, const K2Seg seg, int NC, int D,
                                                     OT* __restrict__ out) {
    extern __shared__ float g[];  // [NC] normalised dlogits of this row
    const int row = blockIdx.x;
    for (int c = threadIdx.x; c < NC; c += blockDim.x) {
        int t = 0;
        while (t + 1 < seg.T && c >= seg.off[t + 1]) ++t;
        const float dn = denom[t];
        g[c] = dn > 0.f ? __fdiv_rn(__ldg(dlogits + (int64_t)row * NC + c), dn) : 0.f;
    }
    __syncthreads();
    for (int k = threadIdx.x * 4; k < D; k += blockDim.x * 4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < NC; ++c) {
            const float4 w = ld4(W + (int64_t)c * D + k);
            const float gc = g[c];
            a.x = fmaf(gc, w.x, a.x); a.y = fmaf(gc, w.y, a.y);
            a.z = fmaf(gc, w.z, a.z); a.w = fmaf(gc, w.w, a.w);
        }
        st4<OT>(out + (int64_t)row * D + k, a);
    }
}

static int fill_seg(K2Seg& seg, const int32_t* seg_offsets, int T, const char* who) {
    if (!seg_offsets || T < 1) { set_error("%s: NULL seg_offsets or T=%d", who, T); return NKBK_E_ARG; }
    if (T > K2_MAX_TASKS) { set_error("%s: T=%d > %d tasks", who, T, K2_MAX_TASKS); return NKBK_E_SHAPE; }
    seg.T = T;
    for (int t = 0; t <= T; ++t) seg.off[t] = seg_offsets[t];
    if (seg.off[0] != 0) { set_error("%s: seg_offsets[0] != 0", who); return NKBK_E_ARG; }
    for (int t = 0; t < T; ++t)
        if (seg.off[t + 1] <= seg.off[t]) { set_error("%s: task %d has no classes", who, t); return NKBK_E_ARG; }
    if (seg.off[T] > K2_MAX_NC) { set_error("%s: %d classes > %d", who, seg.off[T], K2_MAX_NC); return NKBK_E_SHAPE; }
    return NKBK_OK;
}

}  // namespace nkbk

using namespace nkbk;

extern "C" int64_t nkbk_heads_reduce_buf_len(int D, int NC, int T) { return (int64_t)NC * D + NC + 2 * (int64_t)T; }

extern "C" int64_t nkbk_heads_workspace_bytes(int B, int D, int NC, int T) {
    if (B < 1) B = 1;
    return k2_layout(B, D, NC, T).total * (int64_t)sizeof(float);
}

extern "C" int nkbk_heads_fwd_loss_bwd(const void* emb, int emb_dtype, int B, int D, const float* W_cat,
                                       const float* b_cat, const int32_t* seg_offsets, int T, const int64_t* labels,
                                       int loss_kind, float gamma, const float* class_weight, int64_t ignore_index,
                                       float* out_logits, float* out_probs, float* dlogits, float* reduce_buf,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    K2Seg seg;
    int rc = fill_seg(seg, seg_offsets, T, "nkbk_heads_fwd_loss_bwd");
    if (rc) return rc;
    const int NC = seg.off[T];
    NKBK_CHECK_ARG(B >= 0 && D >= 1, "nkbk_heads_fwd_loss_bwd: B=%d D=%d", B, D);
    NKBK_CHECK_ARG(emb_dtype == NKBK_F32 || emb_dtype == NKBK_BF16, "nkbk_heads_fwd_loss_bwd: emb_dtype=%d", emb_dtype);
    NKBK_CHECK_ARG(loss_kind == NKBK_LOSS_CE || loss_kind == NKBK_LOSS_FOCAL, "nkbk_heads_fwd_loss_bwd: loss_kind=%d",
                   loss_kind);
    NKBK_CHECK_ARG(reduce_buf && workspace, "nkbk_heads_fwd_loss_bwd: NULL reduce_buf/workspace");
    if (D % 4 != 0) {
        set_error("nkbk_heads_fwd_loss_bwd: D=%d must be a multiple of 4 (128-bit row loads)", D);
        return NKBK_E_SHAPE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nbuf = nkbk_heads_reduce_buf_len(D, NC, T);
    if (B == 0) {
        NKBK_CHECK_CUDA(cudaMemsetAsync(reduce_buf, 0, nbuf * sizeof(float), st));
        return NKBK_OK;
    }
    NKBK_CHECK_ARG(emb && W_cat && b_cat, "nkbk_heads_fwd_loss_bwd: NULL emb/W/b");
    const K2Layout L = k2_layout(B, D, NC, T);
    if ((int64_t)workspace_bytes < L.total * (int64_t)sizeof(float)) {
        set_error("nkbk_heads_fwd_loss_bwd: workspace %zu < %lld bytes", workspace_bytes,
                  (long long)(L.total * sizeof(float)));
        return NKBK_E_ARG;
    }
    float* ws = static_cast<float*>(workspace);

    K2FwdParams p;
    p.emb = emb; p.W = W_cat; p.bias = b_cat; p.labels = labels; p.class_weight = class_weight;
    p.out_logits = out_logits; p.out_probs = out_probs; p.dlogits = dlogits;
    p.loss_part = ws + L.loss_part;
    p.B = B; p.D = D; p.NC = NC; p.loss_kind = loss_kind; p.gamma = gamma; p.ignore_index = ignore_index;
    p.seg = seg;
    const size_t smem = (size_t)K2_FWD_WARPS * (K2_FWD_ROWS * NC + K2_FWD_ROWS * 2 * T) * sizeof(float);
    if (emb_dtype == NKBK_F32) {
        if (smem > 48 * 1024)
            NKBK_CHECK_CUDA(cudaFuncSetAttribute(k2_heads_forward<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)smem));
        k2_heads_forward<float><<<L.fwd_blocks, K2_FWD_WARPS * 32, smem, st>>>(p);
    } else {
        if (smem > 48 * 1024)
            NKBK_CHECK_CUDA(cudaFuncSetAttribute(k2_heads_forward<__nv_bfloat16>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k2_heads_forward<__nv_bfloat16><<<L.fwd_blocks, K2_FWD_WARPS * 32, smem, st>>>(p);
    }
    NKBK_CHECK_LAUNCH("k2_heads_forward");

    const int have_grads = dlogits != nullptr;
    if (have_grads) {
        dim3 grid((D + K2_DW_THREADS - 1) / K2_DW_THREADS, L.dw_chunks);
        for (int cls0 = 0; cls0 < NC; cls0 += 64) {
            const int rem = NC - cls0;
            float* dwp = ws + L.dw_part;
            float* dbp = ws + L.db_part;
#define NKBK_DW(ET, NCP)                                                                                       \
    k2_heads_dw<ET, NCP><<<grid, K2_DW_THREADS, 0, st>>>(static_cast<const ET*>(emb), dlogits, B, D, NC, cls0, \
                                                         dwp, dbp)
            if (emb_dtype == NKBK_F32) {
                if (rem <= 16) NKBK_DW(float, 16);
                else if (rem <= 32) NKBK_DW(float, 32);
                else NKBK_DW(float, 64);
            } else {
                if (rem <= 16) NKBK_DW(__nv_bfloat16, 16);
                else if (rem <= 32) NKBK_DW(__nv_bfloat16, 32);
                else NKBK_DW(__nv_bfloat16, 64);
            }
#undef NKBK_DW
            NKBK_CHECK_LAUNCH("k2_heads_dw");
        }
    }
    {
        int blocks = (int)((nbuf + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        k2_heads_reduce<<<blocks, 256, 0, st>>>(ws + L.dw_part, ws + L.db_part, ws + L.loss_part, L.dw_chunks,
                                                L.fwd_warps_total, NC, D, T, have_grads, reduce_buf);
        NKBK_CHECK_LAUNCH("k2_heads_reduce");
    }
    return NKBK_OK;
}

extern "C" int nkbk_heads_finalize(float* reduce_buf, int D, const int32_t* seg_offsets, int T, float* out_loss,
                                   int64_t* cm_total, int64_t* cm_step, int64_t n_cm, void* stream) {
    K2Seg seg;
    int rc = fill_seg(seg, seg_offsets, T, "nkbk_heads_finalize");
    if (rc) return rc;
    NKBK_CHECK_ARG(reduce_buf && D >= 1, "nkbk_heads_finalize: NULL reduce_buf or D=%d", D);
    const int NC = seg.off[T];
    const int64_t n = (int64_t)NC * D + NC;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    NKBK_CHECK_ARG(n_cm >= 0 && (n_cm == 0 || (cm_total && cm_step)), "nkbk_heads_finalize: bad confusion buffers");
    k2_heads_finalize<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reduce_buf, NC, D, seg, out_loss, cm_total,
                                                                             cm_step, n_cm);
    NKBK_CHECK_LAUNCH("k2_heads_finalize");
    return NKBK_OK;
}

extern "C" int nkbk_heads_demb(const float* dlogits, const float* reduce_buf, const float* W_cat,
                               const int32_t* seg_offsets, int T, int B, int D, void* out_demb, int out_dtype,
                               void* stream) {
    K2Seg seg;
    int rc = fill_seg(seg, seg_offsets, T, "nkbk_heads_demb");
    if (rc) return rc;
    NKBK_CHECK_ARG(B >= 0 && D >= 1 && D % 4 == 0, "nkbk_heads_demb: B=%d D=%d (D %% 4 must be 0)", B, D);
    if (B == 0) return NKBK_OK;
    NKBK_CHECK_ARG(dlogits && reduce_buf && W_cat && out_demb, "nkbk_heads_demb: NULL pointer");
    NKBK_CHECK_ARG(out_dtype == NKBK_F32 || out_dtype == NKBK_BF16, "nkbk_heads_demb: out_dtype=%d", out_dtype);
    const int NC = seg.off[T];
    const float* denom = reduce_buf + (int64_t)NC * D + NC + T;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (out_dtype == NKBK_F32)
        k2_heads_demb<float><<<B, 256, NC * sizeof(float), st>>>(dlogits, denom, W_cat, seg, NC, D,
                                                                 static_cast<float*>(out_demb));
    else
        k2_heads_demb<__nv_bfloat16><<<B, 256, NC * sizeof(float), st>>>(dlogits, denom, W_cat, seg, NC, D,
                                                                         static_cast<__nv_bfloat16*>(out_demb));
    NKBK_CHECK_LAUNCH("k2_heads_demb");
    return NKBK_OK;
}
