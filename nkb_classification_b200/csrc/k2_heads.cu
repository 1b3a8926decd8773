// K2: all task heads as one segmented GEMM over the pooled features, fused
// with per-task softmax, CE / focal loss and the logit gradient; then the
// weight / bias gradient as a second segmented GEMM (dlogits^T x emb).
// See include/nkbk.h for the contract, DESIGN.md "K2" for the reasoning.
//
// This file is the exact-fp32 path (FFMA accumulate; TF32/bf16 tensor cores
// would break the 1e-5 fp32 parity bar).  bf16 embeddings are widened on load.
// Everything is deterministic: partial sums are combined in a fixed order.
//
//   k2_heads_forward_v3  warp = 4 rows over the whole K (4 chunks of 128 columns in flight,
//                      streamed past L1), weights of a 16-class pass staged once per
//                      CTA in shared memory; a 31-shuffle transposing reduction leaves
//                      the logits in the warp's private shared-memory slice; per-(row,
//                      task) lanes then do softmax / loss / dlogits and the fused K3
//                      (argmax + confusion counts).  No cross-warp reduction.
//   k2_heads_dw        CTA = 256 rows x 128 columns, 8 warps x 32 rows, thread =
//                      4 columns x NCP classes in registers, dlogits broadcast
//                      from shared memory; fixed-order cross-warp tree; the last
//                      CTA to finish a column block sums the per-chunk partials
//                      (again in fixed order) straight into the reduce buffer.
//   k2_heads_finalize  divide by the (all-reduced) denominators, emit losses.
//   k2_heads_demb      d(loss)/d(emb) for an unfrozen backbone.
#include "k2_common.cuh"

#include <algorithm>

namespace nkbk {

struct K2Layout {  // workspace carve-up, in floats
    int fwd_blocks, dw_chunks, dw_xblocks, dw_passes;
    int64_t loss_part;  // [fwd_blocks][2T]
    int64_t dw_part;    // [dw_chunks][NC][D]
    int64_t db_part;    // [dw_chunks][NC]
    int64_t counters;   // uint32 [dw_passes * dw_xblocks]
    int64_t tc_w;       // bf16 copy of the head weights for the tcgen05 path
    int64_t fused;      // per-CTA partials of the fused step (k2_fused.cu)
    int64_t total;
};

static K2Layout k2_layout(int B, int D, int NC, int T) {
    K2Layout L;
    L.fwd_blocks = (B + K2_FWD_ROWS - 1) / K2_FWD_ROWS;
    L.dw_chunks = (B + K2_DW_ROWS - 1) / K2_DW_ROWS;
    L.dw_xblocks = (D + K2_DW_COLS - 1) / K2_DW_COLS;
    L.dw_passes = (NC + K2_DW_NCB - 1) / K2_DW_NCB;
    L.loss_part = 0;
    L.dw_part = (L.loss_part + (int64_t)L.fwd_blocks * 2 * T + 3) & ~int64_t(3);
    L.db_part = L.dw_part + (int64_t)L.dw_chunks * NC * D;
    L.counters = (L.db_part + (int64_t)L.dw_chunks * NC + 3) & ~int64_t(3);
    L.tc_w = (L.counters + (int64_t)L.dw_passes * L.dw_xblocks + 3) & ~int64_t(3);
    L.fused = (L.tc_w + k2_tc_workspace_floats(B, D, NC) + 3) & ~int64_t(3);
    L.total = L.fused + k2_fused_workspace_floats(B, D, NC, T);
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

// ---- packed fp32 (sm_100 FFMA2): two independent IEEE fused multiply-adds per instruction ------------------------
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// acc.lo = fma(a.lo, s, acc.lo); acc.hi = fma(a.hi, s, acc.hi)   (the scalar is a broadcast operand in SASS)
__device__ __forceinline__ void ffma2_bcast(unsigned long long& acc, unsigned long long a, float s) {
    unsigned long long s2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(s2) : "f"(s));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(s2));
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

// Softmax, loss term and dlogits for the K2_FWD_ROWS rows whose logits sit in shared memory `zs` [ROWS][NC];
// executed by one warp in four short stages so that no lane loops over expf():
//   (row, task) lanes: max -> (row, class) lanes: e = exp(z - max) -> (row, task) lanes: sum, lse, loss term, q
//   -> (row, class) lanes: p = exp(z - lse), dlogit = q * (delta - p).
// `ls` is scratch of ROWS * (NC + 6 * T) floats.  Writes this CTA's fixed-order loss / denominator partial.
__device__ __forceinline__ void heads_row_epilogue(const K2FwdParams& p, const float* zs, float* ls, int row0, int lane,
                                                   float* lp) {
    const int T = p.seg.T, NC = p.NC;
    float* es = ls;                                  // [ROWS][NC]   exp(z - max)
    float* mxs = es + K2_FWD_ROWS * NC;              // [ROWS][T]    row max
    float* lses = mxs + K2_FWD_ROWS * T;             // [ROWS][T]    log-sum-exp
    float* qs = lses + K2_FWD_ROWS * T;              // [ROWS][T]    dlogit scale (0 when the row is ignored)
    float* ys = qs + K2_FWD_ROWS * T;                // [ROWS][T]    label as float (-1 when ignored)
    float* lt = ys + K2_FWD_ROWS * T;                // [ROWS][2T]   loss term, denominator term
    for (int idx = lane; idx < K2_FWD_ROWS * T; idx += 32) {
        const int r = idx / T, t = idx - r * T;
        const int c0 = p.seg.off[t], C = p.seg.off[t + 1] - c0;
        const float* z = zs + r * NC + c0;
        float mx = z[0];
        for (int j = 1; j < C; ++j) mx = fmaxf(mx, z[j]);
        mxs[idx] = mx;
    }
    __syncwarp();
    for (int idx = lane; idx < K2_FWD_ROWS * NC; idx += 32) {
        const int r = idx / NC, c = idx - r * NC;
        int t = 0;
        while (t + 1 < T && c >= p.seg.off[t + 1]) ++t;
        es[idx] = expf(zs[idx] - mxs[r * T + t]);
    }
    __syncwarp();
    for (int idx = lane; idx < K2_FWD_ROWS * T; idx += 32) {
        const int r = idx / T, t = idx - r * T;
        const int row = row0 + r;
        const int c0 = p.seg.off[t], C = p.seg.off[t + 1] - c0;
        float se = 0.f;
        for (int j = 0; j < C; ++j) se += es[r * NC + c0 + j];
        const float lse = mxs[idx] + logf(se);
        lses[idx] = lse;
        float q = 0.f, loss_i = 0.f, den_i = 0.f, yf = -1.f;
        if (row < p.B && p.labels != nullptr) {
            const int64_t y = p.labels[(int64_t)row * T + t];
            if (y != p.ignore_index && y >= 0 && y < C) {
                const float logpt = zs[r * NC + c0 + (int)y] - lse;
                const float a = p.class_weight ? __ldg(p.class_weight + c0 + (int)y) : 1.f;
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
                yf = (float)(int)y;
            }
        }
        qs[idx] = q;
        ys[idx] = yf;
        if (p.row_loss != nullptr && row < p.B) p.row_loss[(int64_t)row * T + t] = loss_i;
        lt[r * 2 * T + t] = loss_i;
        lt[r * 2 * T + T + t] = den_i;
    }
    __syncwarp();
    for (int idx = lane; idx < K2_FWD_ROWS * NC; idx += 32) {
        const int r = idx / NC, c = idx - r * NC;
        const int row = row0 + r;
        if (row >= p.B) continue;
        int t = 0;
        while (t + 1 < T && c >= p.seg.off[t + 1]) ++t;
        const float zj = zs[idx];
        const float pj = expf(zj - lses[r * T + t]);
        const int64_t o = (int64_t)row * NC + c;
        if (p.out_logits) p.out_logits[o] = zj;
        if (p.out_probs) p.out_probs[o] = pj;
        if (p.dlogits) {
            const float yf = ys[r * T + t];
            p.dlogits[o] = yf >= 0.f ? qs[r * T + t] * (((float)(c - p.seg.off[t]) == yf ? 1.f : 0.f) - pj) : 0.f;
        }
    }
    for (int i = lane; i < 2 * T; i += 32) {  // fixed order over this CTA's rows
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < K2_FWD_ROWS; ++r) s += lt[r * 2 * T + i];
        lp[i] = s;
    }
}

// ---- forward v3: one warp = K2_FWD_ROWS rows over the whole K, no cross-warp reduction, K3 fused ----------------
// Each warp streams its 4 embedding rows once (K2_V3_PD chunks of 128 columns in flight per lane, so ~8 KB per warp
// is always outstanding), multiplies them against the head weights (read through L1, shared by every warp of the SM),
// folds the 64 per-lane partial sums with the transposing shuffle reduction, and runs the softmax / loss / dlogits
// epilogue on its own rows from its private slice of shared memory -- no __syncthreads anywhere.  The per-task
// argmax and the confusion counts (K3) are taken from the logits while they are still on chip.
// Embedding rows are read exactly once: stream them past L1 so the head weights stay resident there.
__device__ __forceinline__ float4 ld4_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld4_stream(const __nv_bfloat16* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    float4 v;
    v.x = __uint_as_float(r.x << 16);
    v.y = __uint_as_float(r.x & 0xffff0000u);
    v.z = __uint_as_float(r.y << 16);
    v.w = __uint_as_float(r.y & 0xffff0000u);
    return v;
}

// WS = true: the weights of the current 16-class pass ([<=16][D] fp32) are staged once per CTA in shared memory and
// read from there (no L1 misses in the inner loop); WS = false (tile larger than shared memory): through L1.
template <typename ET, bool WS>
__global__ void __launch_bounds__(K2_V3_WARPS * 32) k2_heads_forward_v3(const K2FwdParams p) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.seg.T, NC = p.NC, D = p.D;
    float* zs = smem + warp * (K2_FWD_ROWS * (2 * NC + 6 * T));  // [ROWS][NC] logits of this warp's rows
    float* ls = zs + K2_FWD_ROWS * NC;                            // epilogue scratch, ROWS * (NC + 6T)
    // staged weights start on a 16-byte boundary after the per-warp slices
    float* wsm = smem + ((K2_V3_WARPS * K2_FWD_ROWS * (2 * NC + 6 * T) + 3) & ~3);
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < p.n_counters; i += blockDim.x) p.counters[i] = 0u;
    const int gw = blockIdx.x * K2_V3_WARPS + warp;
    const int row0 = gw * K2_FWD_ROWS;
    const bool active = row0 < p.B;
    if (!WS && !active) return;
    const ET* er[K2_FWD_ROWS];
#pragma unroll
    for (int r = 0; r < K2_FWD_ROWS; ++r)  // tail rows recompute the last row, never stored
        er[r] = static_cast<const ET*>(p.emb) + (int64_t)min(row0 + r, p.B - 1) * D;
    const int nchunks = (D + 127) / 128;

    for (int cb = 0; cb < NC; cb += K2_FWD_NCB) {
        // accumulators as fp32 pairs over adjacent rows: acc2[rp][c] = (row 2 rp, row 2 rp + 1) of class c; one FFMA2
        // per (row pair, class, k) with the weight as the broadcast operand -- per accumulator the same chain of IEEE
        // fmas (k ascending) as the scalar form, in half the instructions
        unsigned long long acc2[K2_FWD_ROWS / 2][K2_FWD_NCB];
#pragma unroll
        for (int rp = 0; rp < K2_FWD_ROWS / 2; ++rp)
#pragma unroll
            for (int i = 0; i < K2_FWD_NCB; ++i) acc2[rp][i] = 0ull;
        float4 e[K2_V3_PD][K2_FWD_ROWS];
        auto load_chunk = [&](float4 (&dst)[K2_FWD_ROWS], int chunk) {
            const int k = chunk * 128 + lane * 4;
#pragma unroll
            for (int r = 0; r < K2_FWD_ROWS; ++r)
                dst[r] = (active && chunk < nchunks && k < D) ? ld4_stream(er[r] + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
#pragma unroll
        for (int u = 0; u < K2_V3_PD; ++u) load_chunk(e[u], u);   // embedding loads are in flight while W is staged
        if (WS) {
            if (cb > 0) __syncthreads();                          // everyone is done with the previous pass's tile
            const int ncls = min(K2_FWD_NCB, NC - cb);
            // cp.async (LDGSTS): every 16-byte piece of the tile is in flight at once, no registers involved
            const float4* src = reinterpret_cast<const float4*>(p.W + (int64_t)cb * D);
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(wsm);
            const int n4 = ncls * (D >> 2);
            for (int i = threadIdx.x; i < n4; i += K2_V3_WARPS * 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)i), "l"(src + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
        if (active)
        for (int c0 = 0; c0 < nchunks; c0 += K2_V3_PD) {
#pragma unroll
            for (int u = 0; u < K2_V3_PD; ++u) {
                const int chunk = c0 + u;
                if (chunk < nchunks) {  // warp-uniform
                    const int k = min(chunk * 128 + lane * 4, D - 4);  // lanes past D hold zeros in e: any valid W address
                    unsigned long long ex[K2_FWD_ROWS / 2], ey[K2_FWD_ROWS / 2], ez[K2_FWD_ROWS / 2], ew[K2_FWD_ROWS / 2];
#pragma unroll
                    for (int rp = 0; rp < K2_FWD_ROWS / 2; ++rp) {
                        ex[rp] = pack_f32x2(e[u][2 * rp].x, e[u][2 * rp + 1].x);
                        ey[rp] = pack_f32x2(e[u][2 * rp].y, e[u][2 * rp + 1].y);
                        ez[rp] = pack_f32x2(e[u][2 * rp].z, e[u][2 * rp + 1].z);
                        ew[rp] = pack_f32x2(e[u][2 * rp].w, e[u][2 * rp + 1].w);
                    }
#pragma unroll
                    for (int ch = 0; ch < K2_FWD_NCB; ch += 8) {
                        if (cb + ch >= NC) break;  // warp-uniform
                        float4 w[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            if (WS) w[c] = *reinterpret_cast<const float4*>(wsm + min(ch + c, NC - cb - 1) * D + k);
                            else w[c] = ld4(p.W + (int64_t)min(cb + ch + c, NC - 1) * D + k);
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            if (cb + ch + c < NC) {  // warp-uniform: padded classes skip their FFMAs
#pragma unroll
                                for (int rp = 0; rp < K2_FWD_ROWS / 2; ++rp) {
                                    ffma2_bcast(acc2[rp][ch + c], ex[rp], w[c].x);
                                    ffma2_bcast(acc2[rp][ch + c], ey[rp], w[c].y);
                                    ffma2_bcast(acc2[rp][ch + c], ez[rp], w[c].z);
                                    ffma2_bcast(acc2[rp][ch + c], ew[rp], w[c].w);
                                }
                            }
                        }
                    }
                    load_chunk(e[u], chunk + K2_V3_PD);  // refill the stage just consumed
                }
            }
        }
        float acc[K2_FWD_ROWS * K2_FWD_NCB];
#pragma unroll
        for (int rp = 0; rp < K2_FWD_ROWS / 2; ++rp)
#pragma unroll
            for (int i = 0; i < K2_FWD_NCB; ++i)
                unpack_f32x2(acc2[rp][i], acc[(2 * rp) * K2_FWD_NCB + i], acc[(2 * rp + 1) * K2_FWD_NCB + i]);
        // 32 * (ROWS / 2) partial sums -> lane L owns entry L (and 32 + L with 4 rows)  (entry = r * 16 + c)
        const int c = cb + (lane & 15), r = lane >> 4;
        const float b = c < NC ? __ldg(p.bias + c) : 0.f;
        {
            float lo[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) lo[i] = acc[i];
            const float s_lo = warp_reduce_scatter32(lo, lane);
            if (c < NC) zs[r * NC + c] = s_lo + b;
        }
        if constexpr (K2_FWD_ROWS == 4) {
            float hi[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) hi[i] = acc[32 + i];
            const float s_hi = warp_reduce_scatter32(hi, lane);
            if (c < NC) zs[(r + 2) * NC + c] = s_hi + b;
        }
    }
    if (!active) return;   // (WS: inactive warps only took part in the staging barriers)
    __syncwarp();
    heads_row_epilogue(p, zs, ls, row0, lane, p.loss_part + (int64_t)gw * 2 * T);

    // ---- K3 fused: per-task argmax (first maximum, NaN maximal -- torch.argmax) + confusion counts ----
    if (p.out_pred != nullptr || p.cm_step != nullptr) {
        for (int idx = lane; idx < K2_FWD_ROWS * T; idx += 32) {
            const int rr = idx / T, t = idx - rr * T;
            const int row = row0 + rr;
            if (row >= p.B) continue;
            const int c0 = p.seg.off[t], C = p.seg.off[t + 1] - c0;
            const float* z = zs + rr * NC + c0;
            float best = z[0];
            int bi = 0;
            for (int j = 1; j < C; ++j) {
                const float v = z[j];
                if (!(best != best) && (v > best || v != v)) { best = v; bi = j; }
            }
            if (p.out_pred) p.out_pred[(int64_t)row * T + t] = bi;
            if (p.cm_step != nullptr && p.labels != nullptr) {
                const int64_t y = p.labels[(int64_t)row * T + t];
                long long bin = -1;
                if (y >= 0 && y < C) {
                    int64_t off = 0;
                    for (int s = 0; s < t; ++s) {
                        const int64_t Cs = p.seg.off[s + 1] - p.seg.off[s];
                        off += Cs * Cs;
                    }
                    bin = off + y * C + bi;
                }
                k3_count_aggregated(p.cm_step, bin);
            }
        }
    }
}

// Sum `n` floats spaced `stride` apart in a fixed order with one warp (lane-strided partial sums, xor tree).
__device__ __forceinline__ float warp_fixed_sum(const float* base, int n, int64_t stride, int lane) {
    float s = 0.f;
    for (int i0 = lane; i0 < n; i0 += 32 * 8) {   // 8 loads in flight per lane, added in index order
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

// ---- dW / db ---------------------------------------------------------------
template <typename ET, int NCP>
__global__ void __launch_bounds__(K2_DW_WARPS * 32, 16 / K2_DW_WARPS) k2_heads_dw(
    const ET* __restrict__ emb, const float* __restrict__ dlogits, int B, int D, int NC, int T, int cls0, int pass,
    float* __restrict__ dw_part, float* __restrict__ db_part, const float* __restrict__ loss_part, int fwd_blocks,
    unsigned int* __restrict__ counters, float* __restrict__ reduce_buf) {
    // dl [ROWS][NCP] while accumulating, then the cross-warp reduction buffer [WARPS/2][NCP][128] (aliased)
    constexpr int RED = (K2_DW_WARPS / 2) * NCP * K2_DW_COLS;
    __shared__ __align__(16) float sm[RED > K2_DW_ROWS * NCP ? RED : K2_DW_ROWS * NCP];
    __shared__ int is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunk = blockIdx.y, chunks = gridDim.y;
    const int rbeg = chunk * K2_DW_ROWS, rows = min(K2_DW_ROWS, B - rbeg);
    const int ncls = min(NCP, NC - cls0);
    float (*dl)[NCP] = reinterpret_cast<float (*)[NCP]>(sm);
    // The warp's 32 rows go through two register stages of 8 rows: 16 row loads are in flight before the dlogits are
    // even staged, and every stage is refilled as soon as it has been consumed.
    constexpr int RB = 8;
    const int k0 = blockIdx.x * K2_DW_COLS + lane * 4;
    const int r_lo = warp * 32, r_hi = min(r_lo + 32, rows);
    float4 eb[2][RB];
    auto load_batch = [&](float4 (&dst)[RB], int r_start) {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            const int r = r_start + i;
            dst[i] = (k0 < D && r < r_hi) ? ld4(emb + (int64_t)(rbeg + r) * D + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    load_batch(eb[0], r_lo);
    load_batch(eb[1], r_lo + RB);
#pragma unroll
    for (int i = threadIdx.x; i < K2_DW_ROWS * NCP; i += K2_DW_WARPS * 32) {
        const int r = i / NCP, c = i - r * NCP;
        dl[r][c] = (r < rows && c < ncls) ? __ldg(dlogits + (int64_t)(rbeg + r) * NC + cls0 + c) : 0.f;
    }
    __syncthreads();

    // accumulators as fp32 pairs: acc2[c][0] = columns (k0, k0+1), acc2[c][1] = (k0+2, k0+3); one FFMA2 (packed fp32
    // FMA, sm_100) updates a pair with the dlogit broadcast -- the same IEEE fma per column, half the instructions
    unsigned long long acc2[NCP][2];
#pragma unroll
    for (int c = 0; c < NCP; ++c) acc2[c][0] = acc2[c][1] = 0ull;
    auto consume = [&](const float4 (&src)[RB], int r_start) {   // rows past r_hi hold e = 0 (and dl = 0)
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            const unsigned long long e01 = pack_f32x2(src[i].x, src[i].y), e23 = pack_f32x2(src[i].z, src[i].w);
#pragma unroll
            for (int c4 = 0; c4 < NCP; c4 += 4) {
                const float4 g = *reinterpret_cast<const float4*>(&dl[r_start + i][c4]);
                const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    ffma2_bcast(acc2[c4 + q][0], e01, gg[q]);
                    ffma2_bcast(acc2[c4 + q][1], e23, gg[q]);
                }
            }
        }
    };
#pragma unroll
    for (int b = 0; b < 32 / RB; b += 2) {
        consume(eb[0], r_lo + b * RB);
        if (b + 2 < 32 / RB) load_batch(eb[0], r_lo + (b + 2) * RB);
        consume(eb[1], r_lo + (b + 1) * RB);
        if (b + 3 < 32 / RB) load_batch(eb[1], r_lo + (b + 3) * RB);
    }
    float acc[NCP][4];
#pragma unroll
    for (int c = 0; c < NCP; ++c) {
        unpack_f32x2(acc2[c][0], acc[c][0], acc[c][1]);
        unpack_f32x2(acc2[c][1], acc[c][2], acc[c][3]);
    }
    // db partial of this chunk (column block 0 only), before dl is overwritten
    if (blockIdx.x == 0 && threadIdx.x < ncls) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += dl[r][threadIdx.x];
        db_part[(int64_t)chunk * NC + cls0 + threadIdx.x] = s;
    }
    __syncthreads();

    // fixed-order cross-warp tree: (0+4) (1+5) (2+6) (3+7) -> (0+2) (1+3) -> (0+1)
    float4 (*red)[NCP][32] = reinterpret_cast<float4 (*)[NCP][32]>(sm);  // [slot][class][lane]
#pragma unroll
    for (int half = K2_DW_WARPS / 2; half >= 1; half >>= 1) {
        if (warp >= half && warp < 2 * half) {
#pragma unroll
            for (int c = 0; c < NCP; ++c)
                red[warp - half][c][lane] = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
        }
        __syncthreads();
        if (warp < half) {
#pragma unroll
            for (int c = 0; c < NCP; ++c) {
                const float4 v = red[warp][c][lane];
                acc[c][0] += v.x; acc[c][1] += v.y; acc[c][2] += v.z; acc[c][3] += v.w;
            }
        }
        __syncthreads();
    }
    const int64_t nW = (int64_t)NC * D;
    if (warp == 0 && k0 < D) {
        float* o = dw_part + ((int64_t)chunk * NC + cls0) * D + k0;
#pragma unroll
        for (int c = 0; c < NCP; ++c)
            if (c < ncls)
                *reinterpret_cast<float4*>(o + (int64_t)c * D) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
    }

    // ---- the last CTA of this column block sums the per-chunk partials into the reduce buffer ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&counters[pass * gridDim.x + blockIdx.x], 1u);
        is_last = (prev == (unsigned int)chunks - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // D % 4 == 0 and column blocks start on multiples of 128: the sums go four columns at a time; the loads of 8
    // chunks are issued together (one L2 round trip per 8 partials), the additions keep their fixed chunk order
    const int cols4 = min(K2_DW_COLS, D - blockIdx.x * K2_DW_COLS) >> 2;
    for (int i = threadIdx.x; i < ncls * cols4; i += K2_DW_WARPS * 32) {
        const int c = i / cols4, q = i - c * cols4;
        const int64_t e = (int64_t)(cls0 + c) * D + blockIdx.x * K2_DW_COLS + 4 * q;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ch0 = 0; ch0 < chunks; ch0 += 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (ch0 + u < chunks) v[u] = __ldcg(reinterpret_cast<const float4*>(dw_part + (int64_t)(ch0 + u) * nW + e));
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (ch0 + u < chunks) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
        }
        *reinterpret_cast<float4*>(reduce_buf + e) = s;
    }
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < ncls; c += K2_DW_WARPS * 32) {
            float s = 0.f;
            for (int ch0 = 0; ch0 < chunks; ch0 += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (ch0 + u < chunks) v[u] = __ldcg(db_part + (int64_t)(ch0 + u) * NC + cls0 + c);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (ch0 + u < chunks) s += v[u];
            }
            reduce_buf[nW + cls0 + c] = s;
        }
        if (pass == 0)  // loss sums / denominators: one warp per entry, fixed order
            for (int j = warp; j < 2 * T; j += K2_DW_WARPS) {
                const float s = warp_fixed_sum(loss_part + j, fwd_blocks, 2 * T, lane);
                if (lane == 0) reduce_buf[nW + NC + j] = s;
            }
    }
    if (threadIdx.x == 0) counters[pass * gridDim.x + blockIdx.x] = 0u;
}

// forward-only calls: loss sums / denominators into the reduce buffer (gradients are zeroed by a memset)
__global__ void __launch_bounds__(256) k2_heads_reduce_loss(const float* __restrict__ loss_part, int fwd_blocks, int T,
                                                            float* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < 2 * T; j += 8) {
        const float s = warp_fixed_sum(loss_part + j, fwd_blocks, 2 * T, lane);
        if (lane == 0) out[j] = s;
    }
}

// Loss on given logits (criterion(pred, true) of the reference API): one warp per 4 rows.
template <typename LT>
__global__ void __launch_bounds__(32) k2_loss_on_logits(const K2FwdParams p, const LT* __restrict__ logits, int ld) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x, NC = p.NC, T = p.seg.T;
    float* zs = smem;
    float* ls = zs + K2_FWD_ROWS * NC;
    const int row0 = blockIdx.x * K2_FWD_ROWS;
    for (int i = lane; i < K2_FWD_ROWS * NC; i += 32) {
        const int r = i / NC, c = i - r * NC;
        const int row = min(row0 + r, p.B - 1);
        zs[i] = load_as_float(logits + (int64_t)row * ld + c);
    }
    __syncwarp();
    heads_row_epilogue(p, zs, ls, row0, lane, p.loss_part + (int64_t)blockIdx.x * 2 * T);
}

// sums[2T] (loss sums, denominators) -> out_loss[T+1]; dlogits /= denom[task]
__global__ void __launch_bounds__(256) k2_loss_normalise(float* __restrict__ dlogits, int B, int NC, const K2Seg seg,
                                                         const float* __restrict__ sums, float* __restrict__ out_loss) {
    const int T = seg.T;
    if (dlogits != nullptr)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)B * NC;
             i += (int64_t)gridDim.x * blockDim.x) {
            const int c = (int)(i % NC);
            int t = 0;
            while (t + 1 < T && c >= seg.off[t + 1]) ++t;
            const float dn = sums[T + t];
            dlogits[i] = dn > 0.f ? __fdiv_rn(dlogits[i], dn) : 0.f;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float total = 0.f;
        for (int t = 0; t < T; ++t) {
            const float l = sums[T + t] > 0.f ? __fdiv_rn(sums[t], sums[T + t]) : 0.f;
            out_loss[t] = l;
            total += l;
        }
        out_loss[T] = total;
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
                                                     const float* __restrict__ task_scale,
                                                     const float* __restrict__ W, const K2Seg seg, int NC, int D,
                                                     OT* __restrict__ out) {
    extern __shared__ float g[];  // [NC] normalised dlogits of this row
    const int row = blockIdx.x;
    for (int c = threadIdx.x; c < NC; c += blockDim.x) {
        int t = 0;
        while (t + 1 < seg.T && c >= seg.off[t + 1]) ++t;
        const float dn = denom[t];
        float v = dn > 0.f ? __fdiv_rn(__ldg(dlogits + (int64_t)row * NC + c), dn) : 0.f;
        if (task_scale != nullptr) v *= task_scale[t];
        g[c] = v;
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
    return nkbk_heads_step(emb, emb_dtype, B, D, W_cat, b_cat, seg_offsets, T, labels, loss_kind, gamma, class_weight,
                           ignore_index, out_logits, out_probs, dlogits, reduce_buf, nullptr, nullptr, workspace,
                           workspace_bytes, stream);
}

static thread_local int g_heads_path = 0;
void nkbk::set_heads_path(int bits) { g_heads_path = bits; }
extern "C" int nkbk_heads_last_path(void) { return g_heads_path; }

// `mode`: KF_MODE_SUMS = the nkbk_heads_step contract (unnormalised sums in reduce_buf); KF_MODE_FINALIZE / KF_MODE_PEER =
// nkbk_heads_train_step (finalize, or K4' exchange + finalize, applied as well).
static int heads_step_impl(const void* emb, int emb_dtype, int B, int D, const float* W_cat, const float* b_cat,
                           const int32_t* seg_offsets, int T, const int64_t* labels, int loss_kind, float gamma,
                           const float* class_weight, int64_t ignore_index, float* out_logits, float* out_probs,
                           float* dlogits, float* reduce_buf, int32_t* out_pred, int64_t* cm_step, void* workspace,
                           size_t workspace_bytes, void* stream, int mode, float* out_loss, int64_t* cm_total,
                           int64_t n_cm, int* fused_done) {
    *fused_done = 0;
    const int64_t w_version = take_weights_version();   // consumed here, so that it can never leak into a later call
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
    p.out_pred = out_pred;
    p.row_loss = nullptr;
    p.cm_step = (cm_step != nullptr && labels != nullptr) ? reinterpret_cast<unsigned long long*>(cm_step) : nullptr;
    p.loss_part = ws + L.loss_part;
    p.counters = reinterpret_cast<unsigned int*>(ws + L.counters);
    p.n_counters = L.dw_passes * L.dw_xblocks;
    p.B = B; p.D = D; p.NC = NC; p.loss_kind = loss_kind; p.gamma = gamma; p.ignore_index = ignore_index;
    p.seg = seg;
    // the whole step as one persistent kernel when the shape qualifies (training calls: gradients wanted)
    {
        const int fz = launch_k2_fused(p, emb_dtype, reduce_buf, ws + L.fused, out_loss, cm_total, n_cm, mode, st);
        if (fz < 0) return fz;
        if (fz > 0) {
            *fused_done = 1;
            set_heads_path(NKBK_PATH_FUSED);
            return NKBK_OK;
        }
    }
    // forward v3: warp = 4 rows over the whole K, private shared-memory slice per warp, K3 fused
    const size_t smem = (size_t)K2_V3_WARPS * K2_FWD_ROWS * (2 * NC + 6 * T) * sizeof(float);
    const int v3_blocks = (L.fwd_blocks + K2_V3_WARPS - 1) / K2_V3_WARPS;
    int loss_parts = L.fwd_blocks;  // rows of the per-warp loss / denominator partial table
    int tc = 0;
    if (emb_dtype == NKBK_BF16) {   // bf16 embeddings: tcgen05 / TMEM / TMA forward when the shape allows
        set_weights_version(w_version);
        tc = launch_k2_tc_forward(p, ws + L.tc_w, st);
        if (tc < 0) return tc;
        if (tc > 0) loss_parts = tc;
    }
    if (tc > 0) {
        // done on the tensor cores, argmax / confusion counts included (same fused epilogue)
    } else {
        // stage the <= 16 x D weight tile of a pass in shared memory when it fits next to the per-warp slices
        const size_t wtile = (size_t)std::min(NC, K2_FWD_NCB) * D * sizeof(float);
        const size_t smem_ws = ((smem + 15) & ~size_t(15)) + wtile;
        const bool ws_fits = smem_ws <= 200 * 1024;
        const size_t sm = ws_fits ? smem_ws : smem;
#define NKBK_FWD(ET, WSV)                                                                                              \
    do {                                                                                                               \
        if (sm > 48 * 1024)                                                                                            \
            NKBK_CHECK_CUDA(cudaFuncSetAttribute(k2_heads_forward_v3<ET, WSV>,                                         \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));               \
        k2_heads_forward_v3<ET, WSV><<<v3_blocks, K2_V3_WARPS * 32, sm, st>>>(p);                                      \
    } while (0)
        if (emb_dtype == NKBK_F32) { if (ws_fits) NKBK_FWD(float, true); else NKBK_FWD(float, false); }
        else { if (ws_fits) NKBK_FWD(__nv_bfloat16, true); else NKBK_FWD(__nv_bfloat16, false); }
#undef NKBK_FWD
    }
    if (tc == 0) NKBK_CHECK_LAUNCH("k2_heads_forward_v3");
    set_heads_path(tc > 0 ? NKBK_PATH_TC_FWD : NKBK_PATH_FFMA_FWD);

    if (dlogits != nullptr) {
        dim3 grid(L.dw_xblocks, L.dw_chunks);
        float* dwp = ws + L.dw_part;
        float* dbp = ws + L.db_part;
        for (int pass = 0; pass < L.dw_passes; ++pass) {
            const int cls0 = pass * K2_DW_NCB;
            const int rem = NC - cls0;
#define NKBK_DW(ET, NCP)                                                                                          \
    k2_heads_dw<ET, NCP><<<grid, K2_DW_WARPS * 32, 0, st>>>(static_cast<const ET*>(emb), dlogits, B, D, NC, T, cls0, \
                                                            pass, dwp, dbp, ws + L.loss_part, loss_parts,          \
                                                            p.counters, reduce_buf)
            if (emb_dtype == NKBK_F32) {
                if (rem <= 4) NKBK_DW(float, 4);
                else if (rem <= 8) NKBK_DW(float, 8);
                else if (rem <= 12) NKBK_DW(float, 12);
                else NKBK_DW(float, 16);
            } else {
                if (rem <= 4) NKBK_DW(__nv_bfloat16, 4);
                else if (rem <= 8) NKBK_DW(__nv_bfloat16, 8);
                else if (rem <= 12) NKBK_DW(__nv_bfloat16, 12);
                else NKBK_DW(__nv_bfloat16, 16);
            }
#undef NKBK_DW
            NKBK_CHECK_LAUNCH("k2_heads_dw");
        }
    } else {
        NKBK_CHECK_CUDA(cudaMemsetAsync(reduce_buf, 0, ((int64_t)NC * D + NC) * sizeof(float), st));
        k2_heads_reduce_loss<<<1, 256, 0, st>>>(ws + L.loss_part, loss_parts, T, reduce_buf + (int64_t)NC * D + NC);
        NKBK_CHECK_LAUNCH("k2_heads_reduce_loss");
    }
    return NKBK_OK;
}

extern "C" int nkbk_heads_step(const void* emb, int emb_dtype, int B, int D, const float* W_cat, const float* b_cat,
                               const int32_t* seg_offsets, int T, const int64_t* labels, int loss_kind, float gamma,
                               const float* class_weight, int64_t ignore_index, float* out_logits, float* out_probs,
                               float* dlogits, float* reduce_buf, int32_t* out_pred, int64_t* cm_step, void* workspace,
                               size_t workspace_bytes, void* stream) {
    int fused = 0;
    return heads_step_impl(emb, emb_dtype, B, D, W_cat, b_cat, seg_offsets, T, labels, loss_kind, gamma, class_weight,
                           ignore_index, out_logits, out_probs, dlogits, reduce_buf, out_pred, cm_step, workspace,
                           workspace_bytes, stream, KF_MODE_SUMS, nullptr, nullptr, 0, &fused);
}

extern "C" int nkbk_heads_train_step(const void* emb, int emb_dtype, int B, int D, const float* W_cat,
                                     const float* b_cat, const int32_t* seg_offsets, int T, const int64_t* labels,
                                     int loss_kind, float gamma, const float* class_weight, int64_t ignore_index,
                                     float* out_logits, float* out_probs, float* dlogits, float* reduce_buf,
                                     int32_t* out_pred, int64_t* cm_step, int64_t* cm_total, int64_t n_cm,
                                     float* out_loss, int exchange, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    NKBK_CHECK_ARG(exchange == NKBK_EXCHANGE_LOCAL || exchange == NKBK_EXCHANGE_PEER,
                   "nkbk_heads_train_step: exchange=%d", exchange);
    NKBK_CHECK_ARG(n_cm >= 0 && (n_cm == 0 || (cm_total && cm_step)), "nkbk_heads_train_step: bad confusion buffers");
    const bool peer = exchange == NKBK_EXCHANGE_PEER && nkbk_peer_world() > 1;
    if (exchange == NKBK_EXCHANGE_PEER && nkbk_peer_world() < 1) {
        set_error("nkbk_heads_train_step: NKBK_EXCHANGE_PEER needs nkbk_peer_init + nkbk_peer_connect");
        return NKBK_E_NCCL;
    }
    int fused = 0;
    int rc = heads_step_impl(emb, emb_dtype, B, D, W_cat, b_cat, seg_offsets, T, labels, loss_kind, gamma, class_weight,
                             ignore_index, out_logits, out_probs, dlogits, reduce_buf, out_pred, cm_step, workspace,
                             workspace_bytes, stream, peer ? KF_MODE_PEER : KF_MODE_FINALIZE, out_loss, cm_total,
                             cm_step ? n_cm : 0, &fused);
    if (rc || fused) return rc;
    // shapes the fused kernel does not take: the same step as separate launches
    if (peer) return nkbk_peer_allreduce_finalize(reduce_buf, D, seg_offsets, T, out_loss, cm_total, cm_step, n_cm, stream);
    return nkbk_heads_finalize(reduce_buf, D, seg_offsets, T, out_loss, cm_total, cm_step, n_cm, stream);
}

extern "C" int64_t nkbk_loss_workspace_bytes(int B, int T) {
    if (B < 1) B = 1;
    return ((int64_t)((B + K2_FWD_ROWS - 1) / K2_FWD_ROWS) * 2 * T + 2 * T) * (int64_t)sizeof(float);
}

extern "C" int nkbk_loss_fwd_bwd(const void* logits, int dtype, int B, int ld, const int32_t* seg_offsets, int T,
                                 const int64_t* labels, int loss_kind, float gamma, const float* class_weight,
                                 int64_t ignore_index, float* out_probs, float* dlogits, float* out_loss,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    K2Seg seg;
    int rc = fill_seg(seg, seg_offsets, T, "nkbk_loss_fwd_bwd");
    if (rc) return rc;
    const int NC = seg.off[T];
    NKBK_CHECK_ARG(B >= 0 && ld >= NC, "nkbk_loss_fwd_bwd: B=%d ld=%d NC=%d", B, ld, NC);
    NKBK_CHECK_ARG(dtype == NKBK_F32 || dtype == NKBK_BF16, "nkbk_loss_fwd_bwd: dtype=%d", dtype);
    NKBK_CHECK_ARG(loss_kind == NKBK_LOSS_CE || loss_kind == NKBK_LOSS_FOCAL, "nkbk_loss_fwd_bwd: loss_kind=%d", loss_kind);
    NKBK_CHECK_ARG(out_loss && workspace && labels, "nkbk_loss_fwd_bwd: NULL out_loss/workspace/labels");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0) {
        NKBK_CHECK_CUDA(cudaMemsetAsync(out_loss, 0, (T + 1) * sizeof(float), st));
        return NKBK_OK;
    }
    NKBK_CHECK_ARG(logits != nullptr, "nkbk_loss_fwd_bwd: NULL logits");
    if ((int64_t)workspace_bytes < nkbk_loss_workspace_bytes(B, T)) {
        set_error("nkbk_loss_fwd_bwd: workspace %zu < %lld bytes", workspace_bytes, (long long)nkbk_loss_workspace_bytes(B, T));
        return NKBK_E_ARG;
    }
    const int blocks = (B + K2_FWD_ROWS - 1) / K2_FWD_ROWS;
    float* ws = static_cast<float*>(workspace);
    float* sums = ws + (int64_t)blocks * 2 * T;
    K2FwdParams p;
    p.emb = nullptr; p.W = nullptr; p.bias = nullptr; p.labels = labels; p.class_weight = class_weight;
    p.out_logits = nullptr; p.out_probs = out_probs; p.dlogits = dlogits; p.loss_part = ws;
    p.out_pred = nullptr; p.cm_step = nullptr; p.row_loss = nullptr;
    p.counters = nullptr; p.n_counters = 0;
    p.B = B; p.D = 0; p.NC = NC; p.loss_kind = loss_kind; p.gamma = gamma; p.ignore_index = ignore_index; p.seg = seg;
    const size_t smem = (size_t)(K2_FWD_ROWS * (2 * NC + 6 * T)) * sizeof(float);
    if (dtype == NKBK_F32) k2_loss_on_logits<float><<<blocks, 32, smem, st>>>(p, static_cast<const float*>(logits), ld);
    else k2_loss_on_logits<__nv_bfloat16><<<blocks, 32, smem, st>>>(p, static_cast<const __nv_bfloat16*>(logits), ld);
    NKBK_CHECK_LAUNCH("k2_loss_on_logits");
    k2_heads_reduce_loss<<<1, 256, 0, st>>>(ws, blocks, T, sums);
    NKBK_CHECK_LAUNCH("k2_heads_reduce_loss");
    int nb = (int)(((int64_t)B * NC + 255) / 256);
    if (nb > 148 * 8) nb = 148 * 8;
    k2_loss_normalise<<<nb, 256, 0, st>>>(dlogits, B, NC, seg, sums, out_loss);
    NKBK_CHECK_LAUNCH("k2_loss_normalise");
    return NKBK_OK;
}

extern "C" int nkbk_loss_rows(const void* logits, int dtype, int B, int ld, const int32_t* seg_offsets, int T,
                              const int64_t* labels, int loss_kind, float gamma, const float* class_weight,
                              int64_t ignore_index, float* out_row_loss, float* dlogits, void* workspace,
                              size_t workspace_bytes, void* stream) {
    K2Seg seg;
    int rc = fill_seg(seg, seg_offsets, T, "nkbk_loss_rows");
    if (rc) return rc;
    const int NC = seg.off[T];
    NKBK_CHECK_ARG(B >= 0 && ld >= NC, "nkbk_loss_rows: B=%d ld=%d NC=%d", B, ld, NC);
    NKBK_CHECK_ARG(dtype == NKBK_F32 || dtype == NKBK_BF16, "nkbk_loss_rows: dtype=%d", dtype);
    NKBK_CHECK_ARG(loss_kind == NKBK_LOSS_CE || loss_kind == NKBK_LOSS_FOCAL, "nkbk_loss_rows: loss_kind=%d", loss_kind);
    NKBK_CHECK_ARG(out_row_loss && workspace && labels, "nkbk_loss_rows: NULL out_row_loss/workspace/labels");
    if (B == 0) return NKBK_OK;
    NKBK_CHECK_ARG(logits != nullptr, "nkbk_loss_rows: NULL logits");
    if ((int64_t)workspace_bytes < nkbk_loss_workspace_bytes(B, T)) {
        set_error("nkbk_loss_rows: workspace %zu < %lld bytes", workspace_bytes, (long long)nkbk_loss_workspace_bytes(B, T));
        return NKBK_E_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = (B + K2_FWD_ROWS - 1) / K2_FWD_ROWS;
    K2FwdParams p;
    p.emb = nullptr; p.W = nullptr; p.bias = nullptr; p.labels = labels; p.class_weight = class_weight;
    p.out_logits = nullptr; p.out_probs = nullptr; p.dlogits = dlogits; p.loss_part = static_cast<float*>(workspace);
    p.out_pred = nullptr; p.cm_step = nullptr; p.row_loss = out_row_loss;
    p.counters = nullptr; p.n_counters = 0;
    p.B = B; p.D = 0; p.NC = NC; p.loss_kind = loss_kind; p.gamma = gamma; p.ignore_index = ignore_index; p.seg = seg;
    const size_t smem = (size_t)(K2_FWD_ROWS * (2 * NC + 6 * T)) * sizeof(float);
    if (dtype == NKBK_F32) k2_loss_on_logits<float><<<blocks, 32, smem, st>>>(p, static_cast<const float*>(logits), ld);
    else k2_loss_on_logits<__nv_bfloat16><<<blocks, 32, smem, st>>>(p, static_cast<const __nv_bfloat16*>(logits), ld);
    NKBK_CHECK_LAUNCH("k2_loss_on_logits");
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
                               const int32_t* seg_offsets, int T, int B, int D, const float* task_scale,
                               void* out_demb, int out_dtype, void* stream) {
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
        k2_heads_demb<float><<<B, 256, NC * sizeof(float), st>>>(dlogits, denom, task_scale, W_cat, seg, NC, D,
                                                                 static_cast<float*>(out_demb));
    else
        k2_heads_demb<__nv_bfloat16><<<B, 256, NC * sizeof(float), st>>>(dlogits, denom, task_scale, W_cat, seg, NC, D,
                                                                         static_cast<__nv_bfloat16*>(out_demb));
    NKBK_CHECK_LAUNCH("k2_heads_demb");
    return NKBK_OK;
}
