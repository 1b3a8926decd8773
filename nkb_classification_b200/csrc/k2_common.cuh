// Definitions shared by k2_heads.cu (exact-fp32 FFMA path) and k2_tc.cu (tcgen05 path for bf16 embeddings).
#pragma once
#include "nkbk_common.cuh"

namespace nkbk {

constexpr int K2_MAX_TASKS = 64;
constexpr int K2_MAX_NC = 1024;
#ifndef NKBK_FWD_ROWS
#define NKBK_FWD_ROWS 4
#endif
constexpr int K2_FWD_ROWS = NKBK_FWD_ROWS;     // rows per warp of the forward / per CTA of the loss-on-logits kernel (2 | 4)
static_assert(K2_FWD_ROWS == 2 || K2_FWD_ROWS == 4, "the transposing reduction handles 32 or 64 partial sums");
constexpr int K2_FWD_NCB = 16;     // classes per pass
#ifndef NKBK_V3_WARPS
#define NKBK_V3_WARPS 4
#endif
#ifndef NKBK_V3_PD
#define NKBK_V3_PD 4
#endif
constexpr int K2_V3_WARPS = NKBK_V3_WARPS;  // forward v3: warps per CTA, each owning K2_FWD_ROWS rows over the full K
constexpr int K2_V3_PD = NKBK_V3_PD;        // forward v3: 128-column chunks of embedding loads in flight per warp
#ifndef NKBK_DW_WARPS
#define NKBK_DW_WARPS 8
#endif
constexpr int K2_DW_WARPS = NKBK_DW_WARPS;
constexpr int K2_DW_ROWS = 32 * NKBK_DW_WARPS;    // rows per dW chunk (32 per warp)
constexpr int K2_DW_COLS = 128;    // columns per dW CTA (4 per lane)
constexpr int K2_DW_NCB = 16;      // classes per dW pass

// K3 in the forward-only kernels: one confusion count per (row, task), warp-aggregated -- the lanes of the calling warp
// that hit the same bin are found with match.any and their leader adds the whole group with ONE 64-bit reduction, so a
// trained model (every row on the diagonal) issues one atomic per warp and bin, not one per row.  Call with the lanes
// that have a count converged (`bin` < 0 = this lane has none); exact integers, order independent.
__device__ __forceinline__ void k3_count_aggregated(unsigned long long* cm, long long bin) {
    const unsigned int act = __activemask();
    const unsigned int have = __ballot_sync(act, bin >= 0);
    if (bin < 0) return;
    const unsigned int peers = __match_any_sync(have, bin);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(cm + bin, (unsigned long long)__popc(peers));
}

struct K2Seg {
    int T;
    int off[K2_MAX_TASKS + 1];
};

struct K2FwdParams {
    const void* emb;
    const float* W;
    const float* bias;
    const int64_t* labels;
    const float* class_weight;
    float* out_logits;
    float* out_probs;
    float* dlogits;
    int32_t* out_pred;     // optional [B][T]: per-task argmax (K3 fused into the forward epilogue)
    unsigned long long* cm_step;  // optional confusion counts to add to (needs labels)
    float* row_loss;       // optional [B][T]: the loss term of every (row, task), 0 where ignored (nkbk_loss_rows)
    float* loss_part;      // [fwd_blocks][2T]
    unsigned int* counters;  // zeroed here for the dW kernel that follows
    int n_counters;
    int B, D, NC;
    int loss_kind;
    float gamma;
    int64_t ignore_index;
    K2Seg seg;
};

// k2_tc.cu: tcgen05 / TMEM / TMA forward for bf16 embeddings.  Returns the number of per-CTA loss partials it
// will write (0 = shape not supported, caller falls back to the FFMA kernel), < 0 on error.
int launch_k2_tc_forward(const K2FwdParams& p, void* wb_workspace, cudaStream_t st);
int64_t k2_tc_workspace_floats(int B, int D, int NC);
// nkbk_heads_weights_version: the hint covers exactly one heads call, whichever kernels end up serving it
void set_weights_version(int64_t v);
int64_t take_weights_version();

// k2_fused.cu: the whole training step of the heads (forward, loss, K3, dW / db, cross-CTA sum and -- by mode -- the
// finalize or the K4' exchange + finalize) as one persistent cooperative kernel.  Returns 1 when launched, 0 when the
// shape or device does not qualify (caller takes the multi-kernel path), < 0 on error.
enum { KF_MODE_SUMS = 0, KF_MODE_FINALIZE = 1, KF_MODE_PEER = 2 };
int launch_k2_fused(const K2FwdParams& f, int emb_dtype, float* reduce_buf, float* ws_floats, float* out_loss,
                    int64_t* cm_total, int64_t n_cm, int mode, cudaStream_t st);
int64_t k2_fused_workspace_floats(int B, int D, int NC, int T);
// which kernels served the calling thread's last heads call (nkbk_heads_last_path)
void set_heads_path(int bits);

}  // namespace nkbk
