// K1 fast path: A.Resize or A.LongestMaxSize + A.PadIfNeeded, full 32*JMAX column
// tiles, 16-byte aligned frame rows.  Same arithmetic as k1_general.cu (bit-exact with cv2),
// different data movement:
//
//   * every warp owns a band of output rows and a private shared-memory ring of
//     source-row slots; one elected lane streams the rows the band needs with
//     cp.async.bulk (TMA bulk copy, SASS UBLKCP) completing on per-slot mbarriers,
//     up to `nslot` rows ahead of the arithmetic -> DRAM latency is hidden by the
//     copy engine instead of by occupancy, and no registers are tied up;
//   * the list of rows to fetch (strictly increasing, <= 2 * rows_per_warp <= 32
//     entries) is built once per band with two ballots and lives one entry per lane;
//   * the two vertically-adjacent filtered rows are kept in two register arrays
//     whose roles flip instead of being copied when the lower tap becomes the
//     upper one;
//   * all per-(lane, j) constants (smem byte offset of the 2-pixel window, funnel
//     shift, packed 11-bit coefficients) are loop invariant, so one source row
//     costs 3 LDS + 2 SHF + 3 PRMT + 3 IDP.2A + 3 SHF per (lane, j);
//   * crops the ring cannot serve (frame rows not 16-byte aligned, boxes wider than
//     ~1500 px, invalid boxes) fall back, per CTA, to the direct-load band routine
//     of k1_general_impl.cuh -- same bits, one launch for the whole batch.
#include "k1_fast_impl.cuh"

// This translation unit is compiled twice: as is (validation / inference pipelines) and, through k1_fast_aug.cu, with
// the train-time augmentations fused (K1_FAST_AUG).
#ifndef K1_FAST_AUG
#define K1_FAST_AUG false
#define K1_FAST_LAUNCHER launch_k1_fast
#endif

namespace nkbk {

// `overlap`: enqueue as a programmatic dependent of the kernel in front (nkbk_k1_overlap_previous): K1 never calls
// griddepcontrol.wait -- it reads nothing that kernel writes -- so it starts as soon as that kernel has released its
// dependents (k2_fused_step does so at entry) and SMs are free.
template <typename Kern>
static void launch_k1_kernel(Kern kern, const K1Params& p, dim3 grid, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(K1_WARPS * 32);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    if (k1_overlap_previous()) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    cudaLaunchKernelEx(&cfg, kern, p);
}

template <int JMAX, bool AUG>
static void launch_fast_j(const K1Params& p, dim3 grid, cudaStream_t st, bool f32) {
    if (p.mode == NKBK_MODE_LETTERBOX && (p.padu[0] | p.padu[1] | p.padu[2]) == 0u) {   // PadIfNeeded(value = 0)
        if (f32) launch_k1_kernel(k1_crop_resize_normalize_tma<JMAX, float, 2, AUG>, p, grid, st);
        else launch_k1_kernel(k1_crop_resize_normalize_tma<JMAX, __nv_bfloat16, 2, AUG>, p, grid, st);
    } else if (p.mode == NKBK_MODE_LETTERBOX) {
        if (f32) launch_k1_kernel(k1_crop_resize_normalize_tma<JMAX, float, 1, AUG>, p, grid, st);
        else launch_k1_kernel(k1_crop_resize_normalize_tma<JMAX, __nv_bfloat16, 1, AUG>, p, grid, st);
    } else {
        if (f32) launch_k1_kernel(k1_crop_resize_normalize_tma<JMAX, float, 0, AUG>, p, grid, st);
        else launch_k1_kernel(k1_crop_resize_normalize_tma<JMAX, __nv_bfloat16, 0, AUG>, p, grid, st);
    }
}

// Returns false when no fast instantiation exists for this column-tile width.
bool K1_FAST_LAUNCHER(const K1Params& p, int jmax, dim3 grid, cudaStream_t st, bool f32) {
    switch (jmax) {
        case 4: launch_fast_j<4, K1_FAST_AUG>(p, grid, st, f32); return true;
        case 5: launch_fast_j<5, K1_FAST_AUG>(p, grid, st, f32); return true;
        case 6: launch_fast_j<6, K1_FAST_AUG>(p, grid, st, f32); return true;
        case 7: launch_fast_j<7, K1_FAST_AUG>(p, grid, st, f32); return true;
        case 8: launch_fast_j<8, K1_FAST_AUG>(p, grid, st, f32); return true;
        default: return false;
    }
}

}  // namespace nkbk
