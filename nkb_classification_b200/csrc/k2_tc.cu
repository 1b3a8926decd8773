// K2 forward on the 5th-generation tensor cores, for bf16 embeddings (the
// autocast path): logits[B, NC] = emb[B, D] . W^T, accumulated in fp32 in TMEM,
// fused with the per-task softmax / CE / focal loss / dlogits epilogue.
//
//   CTA  = one 128-row tile of the batch, all classes (N = NC padded to 16/32/64/128/256)
//   warp 0   TMA producer: emb tile [128 x 64] + weight tile [N x 64] bf16 per k-block,
//            cp.async.bulk.tensor.2d, SWIZZLE_128B, 4-stage mbarrier ring
//   warp 1   TMEM allocation + MMA issuer: one elected lane issues
//            tcgen05.mma.cta_group::1.kind::f16 (M=128, N, K=16) x 4 per k-block,
//            tcgen05.commit releases the stage / signals the epilogue
//   warps 2-5  epilogue: tcgen05.ld 32x32b (one accumulator row per thread), bias,
//            then softmax / loss / dlogits for that row
//
// The contraction is tiny next to the ~250 flop/B ridge (N <= 64 classes): the kernel is
// bound by streaming emb once from HBM; tensor cores are used because it IS a contraction
// and they take the 2*D*NC FMAs per row off the FMA pipe.  fp32 embeddings stay on the
// exact FFMA path (k2_heads.cu): TF32/bf16 operands would break the 1e-5 fp32 bar.
#include <cuda.h>
#include <stdlib.h>

#include "k2_common.cuh"

namespace nkbk {

constexpr int TC_BM = 128;     // rows per CTA (UMMA M)
constexpr int TC_BK = 64;      // bf16 per k-block = one 128-byte swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr int TC_MAX_TASKS = 16;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TC_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TC_DONE_%=;\n\t"
        "bra TC_WAIT_%=;\n\t"
        "TC_DONE_%=:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO = 64 x 16 B),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.  Tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;          // leading byte offset (unused for swizzled K-major), 16-byte units
    d |= (uint64_t)64 << 32;         // stride byte offset: 1024 B between 8-row groups
    d |= (uint64_t)1 << 46;          // descriptor version
    d |= (uint64_t)2 << 61;          // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcSmem {  // carve-up of dynamic shared memory (after 1024-byte alignment)
    uint32_t a_off, b_off, lt_off, bar_off, total;   // the logits rows z[128][npad+1] alias the stage buffers
};
__host__ __device__ inline TcSmem tc_smem_layout(int npad, int stages, int T) {
    TcSmem L;
    L.a_off = 0;
    L.b_off = L.a_off + stages * TC_BM * 128;
    uint32_t end = L.b_off + stages * npad * 128;
    const uint32_t zbytes = TC_BM * (npad + 1) * 4;
    if (end < zbytes) end = zbytes;
    L.lt_off = (end + 15) & ~15u;
    L.bar_off = (L.lt_off + TC_BM * 2 * T * 4 + 15) & ~15u;
    L.total = L.bar_off + (2 * stages + 1) * 8 + 16;
    return L;
}

struct TcSplit {            // split-K bookkeeping (global workspace)
    float* partial;         // [ksplit][B][NPAD] fp32 partial accumulators
    unsigned int* counters; // [m tiles], zeroed by k2_tc_pack_weights
};

template <int NPAD, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) k2_tc_heads_forward(const __grid_constant__ CUtensorMap tm_emb,
                                                                   const __grid_constant__ CUtensorMap tm_w,
                                                                   const K2FwdParams p, const TcSplit sp) {
    extern __shared__ uint8_t tc_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~uintptr_t(1023));
    const int T = p.seg.T, NC = p.NC;
    const TcSmem L = tc_smem_layout(NPAD, STAGES, T);
    const uint32_t sA = tc_smem_u32(smem + L.a_off), sB = tc_smem_u32(smem + L.b_off);
    float* zs = reinterpret_cast<float*>(smem + L.a_off);     // [128][NPAD+1] logits (stage buffers are free by then)
    float* lt = reinterpret_cast<float*>(smem + L.lt_off);    // [128][2T] loss / denominator terms
    const uint32_t bars = tc_smem_u32(smem + L.bar_off);      // full[S], empty[S], tmem_full
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + L.bar_off + (2 * STAGES + 1) * 8);
    const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tmem_full = bars + 16 * STAGES;
    __shared__ int s_is_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * TC_BM;
    const int ks = gridDim.y, ksi = blockIdx.y;
    const int kblocks_all = p.D / TC_BK;
    const int kb_begin = (int)((int64_t)kblocks_all * ksi / ks), kb_end = (int)((int64_t)kblocks_all * (ksi + 1) / ks);
    const int kblocks = kb_end - kb_begin;   // >= 1 (the host keeps ks <= kblocks_all)
    constexpr uint32_t kTmemCols = NPAD < 32 ? 32 : NPAD;
    constexpr uint32_t kStageBytes = TC_BM * 128 + NPAD * 128;

    if (blockIdx.x == 0 && blockIdx.y == 0 && p.counters != nullptr)
        for (int i = threadIdx.x; i < p.n_counters; i += blockDim.x) p.counters[i] = 0u;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            tc_mbar_init(full0 + 8 * s, 1);
            tc_mbar_init(empty0 + 8 * s, 1);
        }
        tc_mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_emb)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_holder)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int i = 0; i < kblocks; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                tc_mbar_wait(empty0 + 8 * s, ph ^ 1);
                tc_mbar_expect_tx(full0 + 8 * s, kStageBytes);
                tc_tma_load_2d(sA + s * (TC_BM * 128), &tm_emb, full0 + 8 * s, (kb_begin + i) * TC_BK, m0);
                tc_tma_load_2d(sB + s * (NPAD * 128), &tm_w, full0 + 8 * s, (kb_begin + i) * TC_BK, 0);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3, M >> 4
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NPAD >> 3) << 17) |
                                   ((uint32_t)(TC_BM >> 4) << 24);
        if (lane == 0) {
            for (int i = 0; i < kblocks; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                tc_mbar_wait(full0 + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t adesc = tc_smem_desc(sA + s * (TC_BM * 128));
                const uint64_t bdesc = tc_smem_desc(sB + s * (NPAD * 128));
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)  // advance 32 bytes (= 2 x 16 B) along K inside the swizzle row
                    tc_mma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                tc_commit(empty0 + 8 * s);   // frees the stage once these MMAs have read it
            }
            tc_commit(tmem_full);            // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (split-K: partials, last CTA sums) -> softmax / loss / dlogits =====
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                    // row inside the tile
        const int row = m0 + r;
        const int e = threadIdx.x - 64;                 // 0..127 over the epilogue warps
        tc_mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* z = zs + r * (NPAD + 1);
        bool active = true;
        if (ks == 1) {
#pragma unroll
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (c0 + i < NC) z[c0 + i] = v[i] + __ldg(p.bias + c0 + i);
            }
        } else {
            float* mine = sp.partial + ((int64_t)ksi * p.B + row) * NPAD;
#pragma unroll
            for (int c0 = 0; c0 < NPAD; c0 += 16) {
                float v[16];
                tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
                if (row < p.B) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(mine + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (e == 0) s_is_last = (atomicAdd(&sp.counters[blockIdx.x], 1u) == (unsigned)ks - 1u);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            active = s_is_last != 0;
            if (active) {
                __threadfence();
                if (e == 0) sp.counters[blockIdx.x] = 0u;
                if (row < p.B) {
                    // all ks partials of four classes are requested together (one L2 round trip per group instead of
                    // one per (class, split)); the additions keep their fixed split order
                    const float* base = sp.partial + (int64_t)row * NPAD;
                    const int64_t kstride = (int64_t)p.B * NPAD;
#pragma unroll
                    for (int c0 = 0; c0 < NPAD; c0 += 4) {
                        if (c0 >= NC) break;
                        float4 v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < ks) v[k] = __ldcg(reinterpret_cast<const float4*>(base + k * kstride + c0));
                        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (k < ks) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
                        const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (c0 + i < NC) z[c0 + i] = sv[i] + __ldg(p.bias + c0 + i);
                    }
                }
            }
        }
        if (active) {
            float* my_lt = lt + r * 2 * T;
            for (int t = 0; t < T; ++t) {
                float loss_i = 0.f, den_i = 0.f;
                if (row < p.B) {
                    const int c0 = p.seg.off[t], C = p.seg.off[t + 1] - c0;
                    const float* zt = z + c0;
                    float mx = zt[0];
                    for (int j = 1; j < C; ++j) mx = fmaxf(mx, zt[j]);
                    float se = 0.f;
                    for (int j = 0; j < C; ++j) se += expf(zt[j] - mx);
                    const float lse = mx + logf(se);
                    const int64_t y = p.labels ? p.labels[(int64_t)row * T + t] : p.ignore_index;
                    const bool keep = p.labels != nullptr && y != p.ignore_index && y >= 0 && y < C;
                    float qv = 0.f;
                    if (keep) {
                        const float logpt = zt[y] - lse;
                        const float a = p.class_weight ? __ldg(p.class_weight + c0 + (int)y) : 1.f;
                        if (p.loss_kind == NKBK_LOSS_FOCAL) {
                            const float pt = expf(logpt);
                            const float om = 1.f - pt;
                            const float g = p.gamma;
                            float ft, dterm;
                            if (g == 0.f) { ft = 1.f; dterm = 0.f; }
                            else {
                                const float pw1 = (g == 1.f) ? 1.f : ((g == 2.f) ? om : powf(om, g - 1.f));
                                ft = pw1 * om;
                                dterm = g * pt * pw1 * logpt;
                            }
                            loss_i = -a * ft * logpt;
                            qv = a * (dterm - ft);
                            den_i = 1.f;
                        } else {
                            loss_i = -a * logpt;
                            qv = -a;
                            den_i = a;
                        }
                    }
                    const int64_t o = (int64_t)row * NC + c0;
                    for (int j = 0; j < C; ++j) {
                        const float zj = zt[j];
                        const float pj = expf(zj - lse);
                        if (p.out_logits) p.out_logits[o + j] = zj;
                        if (p.out_probs) p.out_probs[o + j] = pj;
                        if (p.dlogits) p.dlogits[o + j] = keep ? qv * ((j == (int)y ? 1.f : 0.f) - pj) : 0.f;
                    }
                }
                my_lt[t] = loss_i;
                my_lt[T + t] = den_i;
                // ---- K3 fused: argmax (first maximum, NaN maximal) + confusion count of this (row, task) ----
                if (row < p.B && (p.out_pred != nullptr || p.cm_step != nullptr)) {
                    const int c0 = p.seg.off[t], C = p.seg.off[t + 1] - c0;
                    const float* zt = z + c0;
                    float best = zt[0];
                    int bi = 0;
                    for (int j = 1; j < C; ++j) {
                        const float v = zt[j];
                        if (!(best != best) && (v > best || v != v)) { best = v; bi = j; }
                    }
                    if (p.out_pred) p.out_pred[(int64_t)row * T + t] = bi;
                    if (p.cm_step != nullptr && p.labels != nullptr) {
                        const int64_t y = p.labels[(int64_t)row * T + t];
                        long long bin = -1;
                        if (y >= 0 && y < C) {
                            int64_t off = 0;
                            for (int u = 0; u < t; ++u) {
                                const int64_t Cu = p.seg.off[u + 1] - p.seg.off[u];
                                off += Cu * Cu;
                            }
                            bin = off + y * C + bi;
                        }
                        k3_count_aggregated(p.cm_step, bin);   // 32 rows per warp: a trained model -> ~1 atomic per warp
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        // named barrier over the 4 epilogue warps, then a fixed-order column sum -> this row tile's partial
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (active && e < 2 * T) {
            float s = 0.f;
            for (int rr = 0; rr < TC_BM; ++rr) s += lt[rr * 2 * T + e];
            p.loss_part[(int64_t)blockIdx.x * 2 * T + e] = s;
        }
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// fp32 head weights -> bf16 [npad][D], rows >= NC zero (the tensor map's box always reads npad rows)
__global__ void __launch_bounds__(256) k2_tc_pack_weights(const float* __restrict__ W, int NC, int D, int npad,
                                                          __nv_bfloat16* __restrict__ out,
                                                          unsigned int* __restrict__ counters, int n_counters) {
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < n_counters; i += blockDim.x) counters[i] = 0u;
    const int64_t n = (int64_t)npad * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / D);
        out[i] = __float2bfloat16_rn(c < NC ? __ldg(W + i) : 0.f);
    }
}

typedef CUresult (*PFN_tmap_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmap_encode tmap_encoder() {
    static PFN_tmap_encode fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmap_encode>(ptr);
    }
    return fn;
}

static bool make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_tmap_encode enc = tmap_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int tc_npad(int NC) {
    if (NC <= 16) return 16;
    if (NC <= 32) return 32;
    if (NC <= 64) return 64;
    if (NC <= 128) return 128;
    if (NC <= 256) return 256;
    return 0;
}

static int tc_ksplit(int B, int D, int npad) {
    const int ctas_m = (B + TC_BM - 1) / TC_BM;
    const int kblocks = D / TC_BK;
    int ks = 148 / ctas_m;                       // fill the 148 SMs
    if (ks > 8) ks = 8;
    if (ks > kblocks) ks = kblocks;
    while (ks > 1 && (int64_t)ks * B * npad * 4 > (int64_t(64) << 20)) --ks;   // cap the partial buffer at 64 MB
    return ks < 1 ? 1 : ks;
}

// bf16 weights + split-K partials + per-tile counters, in floats
int64_t k2_tc_workspace_floats(int B, int D, int NC) {
    const int npad = tc_npad(NC);
    if (npad == 0 || D % TC_BK != 0) return 0;
    const int ks = tc_ksplit(B, D, npad);
    const int64_t wb = ((int64_t)npad * D * 2 + 1024 + 3) / 4;
    const int64_t part = ks > 1 ? (int64_t)ks * B * npad + 4 : 0;
    return wb + part + (B + TC_BM - 1) / TC_BM + 8;
}

template <int NPAD, int STAGES>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tw, const K2FwdParams& p, const TcSplit& sp, dim3 grid,
                     cudaStream_t st) {
    const TcSmem L = tc_smem_layout(NPAD, STAGES, p.seg.T);
    const int smem = (int)L.total + 1024;
    NKBK_CHECK_CUDA(cudaFuncSetAttribute(k2_tc_heads_forward<NPAD, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem));
    k2_tc_heads_forward<NPAD, STAGES><<<grid, TC_THREADS, smem, st>>>(ta, tw, p, sp);
    NKBK_CHECK_LAUNCH("k2_tc_heads_forward");
    return NKBK_OK;
}

constexpr int TC_PACK_CACHE = 16;
struct TcPackEntry {
    const void* wb = nullptr;
    const void* W = nullptr;
    int64_t version = 0;
    int NC = 0, D = 0, ctas = 0;
};
static TcPackEntry g_pack_cache[TC_PACK_CACHE];
static unsigned g_pack_next = 0;
static thread_local int64_t g_weights_version = 0;
void set_weights_version(int64_t v) { g_weights_version = v; }
int64_t take_weights_version() {
    const int64_t v = g_weights_version;
    g_weights_version = 0;
    return v;
}

int launch_k2_tc_forward(const K2FwdParams& p, void* ws_v, cudaStream_t st) {
    const int npad = tc_npad(p.NC);
    if (npad == 0 || p.seg.T > TC_MAX_TASKS || p.D % TC_BK != 0 || p.B < 1) return 0;
    if ((reinterpret_cast<uintptr_t>(p.emb) & 15) != 0) return 0;
    if (getenv("NKBK_DISABLE_TCGEN05") != nullptr) return 0;
    const int ks = tc_ksplit(p.B, p.D, npad);
    const int ctas = (p.B + TC_BM - 1) / TC_BM;
    float* ws = static_cast<float*>(ws_v);
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
    float* after_wb = ws + ((int64_t)npad * p.D * 2 + 1024 + 3) / 4;
    TcSplit sp;
    sp.partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(after_wb) + 15) & ~uintptr_t(15));
    sp.counters = reinterpret_cast<unsigned int*>(after_wb + (ks > 1 ? (int64_t)ks * p.B * npad + 4 : 0));
    CUtensorMap ta, tw;
    if (!make_tmap_bf16_2d(&ta, p.emb, (uint64_t)p.B, (uint64_t)p.D, TC_BM)) return 0;
    if (!make_tmap_bf16_2d(&tw, wb, (uint64_t)npad, (uint64_t)p.D, (uint32_t)npad)) return 0;
    // The bf16 copy of the weights lives in the caller's workspace.  When the caller has told us which version of the
    // weights this is (nkbk_heads_weights_version: e.g. torch's tensor version counter, bumped by every optimizer step)
    // and the workspace already holds that version, the re-pack is skipped: validation / inference loops pack once.
    bool packed = false;
    const int64_t ver = g_weights_version;
    g_weights_version = 0;                        // a version hint covers one call
    if (ver != 0)
        for (int i = 0; i < TC_PACK_CACHE; ++i)
            if (g_pack_cache[i].wb == wb && g_pack_cache[i].W == p.W && g_pack_cache[i].version == ver &&
                g_pack_cache[i].NC == p.NC && g_pack_cache[i].D == p.D && g_pack_cache[i].ctas >= ctas)
                packed = true;
    if (!packed) {
        const int64_t n = (int64_t)npad * p.D;
        int blocks = (int)((n + 255) / 256);
        if (blocks > 148 * 4) blocks = 148 * 4;
        k2_tc_pack_weights<<<blocks, 256, 0, st>>>(p.W, p.NC, p.D, npad, wb, sp.counters, ctas);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("launch of k2_tc_pack_weights failed: %s", cudaGetErrorString(e)); return NKBK_E_CUDA; }
        count_launch();
        if (ver != 0) {
            TcPackEntry& e2 = g_pack_cache[g_pack_next++ % TC_PACK_CACHE];
            e2.wb = wb; e2.W = p.W; e2.version = ver; e2.NC = p.NC; e2.D = p.D; e2.ctas = ctas;
        } else {
            for (int i = 0; i < TC_PACK_CACHE; ++i)
                if (g_pack_cache[i].wb == wb) g_pack_cache[i].wb = nullptr;   // overwritten with an unknown version
        }
    }
    dim3 grid(ctas, ks);
    int rc;
    switch (npad) {   // stage count: as deep as ~150 KB of operand buffers allows
        case 16: rc = launch_tc<16, 8>(ta, tw, p, sp, grid, st); break;
        case 32: rc = launch_tc<32, 7>(ta, tw, p, sp, grid, st); break;
        case 64: rc = launch_tc<64, 6>(ta, tw, p, sp, grid, st); break;
        case 128: rc = launch_tc<128, 4>(ta, tw, p, sp, grid, st); break;
        default: rc = launch_tc<256, 3>(ta, tw, p, sp, grid, st); break;
    }
    return rc == NKBK_OK ? ctas : rc;
}

}  // namespace nkbk

extern "C" void nkbk_heads_weights_version(int64_t version) { nkbk::set_weights_version(version); }
