// Shared pieces of the two K1 kernels (k1_fast.cu: TMA-staged fast path,
// k1_general.cu: direct-load general path).
#pragma once
#include "nkbk_common.cuh"
#include "k1_coef.h"

namespace nkbk {

constexpr int K1_WARPS = 4;            // warps per CTA, one band of output rows each
constexpr int K1F_RING_BYTES = 9216;   // per-warp shared-memory ring of staged source rows
constexpr int K1F_MAX_SLOTS = 8;

struct K1Params {
    const uint8_t* frames;
    const int64_t* frame_desc;
    const int32_t* boxes;
    const int32_t* frame_idx;
    int n, n_frames, mode, out_h, out_w, max_size;
    float m[3], d[3];
    float padf[3];       // normalised pad value per output channel
    uint32_t padu[3];    // raw pad value per output channel
    uint32_t sel[3];     // PRMT selectors per output channel (encode channel_swap)
    void* out;
    uint8_t* out_u8;
    int32_t* bad_count;
    int rows_per_warp;       // general kernel (<= 16)
    int rows_per_warp_fast;  // TMA kernel (<= 32)
    // ---- train-time augmentations on the resized (and padded) uint8 image, per crop (NULL = none) ----
    // applied in the reference's order: HorizontalFlip, VerticalFlip, RandomBrightnessContrast, CoarseDropout
    const int32_t* aug_flags;   // [n]  bit0 hflip, bit1 vflip, bit2 brightness/contrast, bits 8.. = number of holes
    const float* aug_alpha;     // [n]  f32(1 + contrast)
    const float* aug_beta;      // [n]  f32(brightness * max_value), added after the multiply
    const int32_t* aug_holes;   // [n][aug_max_holes][4]  x1, y1, x2, y2 in output pixels (exclusive ends)
    int aug_max_holes;
    uint32_t aug_fill[3];       // CoarseDropout fill value per output channel
};
constexpr int K1_AUG_HFLIP = 1, K1_AUG_VFLIP = 2, K1_AUG_BC = 4;
constexpr int K1_AUG_MAX_HOLES = 16;

// albumentations 1.x `_brightness_contrast_adjust_uint` with beta_by_max: the 256-entry LUT
// clip(f32(v) * alpha + beta, 0, 255).astype(uint8) evaluated per pixel (two rounded fp32 ops, truncation).
__host__ __device__ __forceinline__ uint32_t k1_brightness_contrast(uint32_t v, float alpha, float beta) {
#ifdef __CUDA_ARCH__
    float t = __fadd_rn(__fmul_rn((float)v, alpha), beta);
    t = fminf(fmaxf(t, 0.f), 255.f);
    return (uint32_t)__float2int_rz(t);
#else
    volatile float t = (float)v * alpha;
    t = t + beta;
    float u = t < 0.f ? 0.f : (t > 255.f ? 255.f : t);
    return (uint32_t)(int)u;
#endif
}

template <typename OutT>
__device__ __forceinline__ void store_out(OutT* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) {
    __stcs(p, v);
}
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    __stcs(reinterpret_cast<unsigned short*>(p), __bfloat16_as_ushort(__float2bfloat16_rn(v)));
}

struct CropGeom {
    int bx0, by0, bw, bh, fw, fh;
    int64_t f_off, pitch;
    bool ok;  // box non-empty and inside its frame, frame index valid, frame at least 2 pixels wide
};

__device__ __forceinline__ CropGeom load_geom(const K1Params& p, int crop) {
    CropGeom g;
    g.bx0 = __ldg(p.boxes + 4 * (int64_t)crop + 0);
    g.by0 = __ldg(p.boxes + 4 * (int64_t)crop + 1);
    const int bx1 = __ldg(p.boxes + 4 * (int64_t)crop + 2), by1 = __ldg(p.boxes + 4 * (int64_t)crop + 3);
    const int fi = __ldg(p.frame_idx + crop);
    g.ok = fi >= 0 && fi < p.n_frames;
    g.f_off = 0; g.pitch = 0; g.fh = 0; g.fw = 0;
    if (g.ok) {
        const int64_t* fd = p.frame_desc + 4 * (int64_t)fi;
        g.f_off = __ldg(fd + 0);
        g.fh = (int)__ldg(fd + 1);
        g.fw = (int)__ldg(fd + 2);
        g.pitch = __ldg(fd + 3);
    }
    g.bw = bx1 - g.bx0;
    g.bh = by1 - g.by0;
    g.ok = g.ok && g.bx0 >= 0 && g.by0 >= 0 && bx1 <= g.fw && by1 <= g.fh && g.bw >= 1 && g.bh >= 1 && g.fw >= 2;
    return g;
}

// The fast kernel stages, per source row, the 16-byte-aligned span that covers the box's pixels
// (plus the one pixel to its right that the last 2-tap window may touch) with one bulk async copy.
// That needs 16-byte aligned frame rows, and the span must leave room for >= 2 ring slots.
// Crops that fail the predicate are produced by the direct-load band routine inside the same kernel.
__device__ __forceinline__ bool fast_path_qualifies(const K1Params& p, const CropGeom& g, uint32_t& seg_start,
                                                    uint32_t& seg_bytes, uint32_t& slot_stride, int& nslot) {
    seg_start = 0; seg_bytes = 0; slot_stride = 0; nslot = 0;
    if (!g.ok) return false;
    if (((reinterpret_cast<uintptr_t>(p.frames) + (uintptr_t)g.f_off) & 15u) != 0 || (g.pitch & 15) != 0) return false;
    const uint32_t s = (uint32_t(g.bx0) * 3u) & ~15u;
    const uint32_t e_px = (uint32_t)min(g.bx0 + g.bw + 1, g.fw);
    const uint32_t e = (e_px * 3u + 15u) & ~15u;
    seg_start = s;
    seg_bytes = e - s;
    slot_stride = (seg_bytes + 16u + 127u) & ~127u;  // +16: the unconditional third-word read may run past the span
    nslot = min(K1F_MAX_SLOTS, (int)(K1F_RING_BYTES / slot_stride));
    if ((int64_t)g.bh * g.pitch >= (int64_t(1) << 31)) return false;  // row offsets are kept in 32 bits
    return nslot >= 2;
}

}  // namespace nkbk
