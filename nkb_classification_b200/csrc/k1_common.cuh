// Shared pieces of the two K1 kernels (k1_fast.cu: TMA-staged fast path,
// k1_general.cu: direct-load general path).
#pragma once
#include "nkbk_common.cuh"
#include "k1_coef.h"
#include <math.h>

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
    int fby_fast;            // TMA kernel: row blocks per crop; blockIdx.x = crop * fby_fast + row block (crop-major)
    unsigned long long* timing;   // debug (NKBK_K1_TIMING=1): per CTA {SM id, %globaltimer at entry, at exit}, else NULL
    // ---- train-time augmentations on the resized (and padded) uint8 image, per crop (NULL = none) ----
    // applied in the reference's order: HorizontalFlip, VerticalFlip, RandomBrightnessContrast, CoarseDropout
    const int32_t* aug_flags;   // [n]  bit0 hflip, bit1 vflip, bit2 brightness/contrast, bits 8.. = number of holes
    const float* aug_alpha;     // [n]  f32(1 + contrast)
    const float* aug_beta;      // [n]  f32(brightness * max_value), added after the multiply
    const int32_t* aug_holes;   // [n][aug_max_holes][4]  x1, y1, x2, y2 in output pixels (exclusive ends)
    int aug_max_holes;
    uint32_t aug_fill[3];       // CoarseDropout fill value per output channel
    const uint8_t* aug_hsv_lut; // [n][3][256]  HueSaturationValue look-up tables (hue, sat, val), read when bit 3 is set
    int aug_hsv_trunc_cols;     // output columns [0, this) take cv2's SIMD rounding (truncation), the rest its scalar one
};
constexpr int K1_AUG_HFLIP = 1, K1_AUG_VFLIP = 2, K1_AUG_BC = 4, K1_AUG_HSV = 8;
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

// ---- A.HueSaturationValue on uint8 RGB: cv2.cvtColor(RGB2HSV) -> three 256-entry LUTs -> cv2.cvtColor(HSV2RGB) ----
// RGB2HSV is OpenCV's 8-bit fixed-point path (hsv_shift = 12, division tables cvRound((255 << 12) / v) and
// cvRound((180 << 12) / (6 diff)), evaluated here as exact round-half-even integer divisions); HSV2RGB is its
// float path with the contraction OpenCV's build uses: tab2 = v * fma(-s, f, 1), tab3 = v * fma(-s, 1 - f, 1),
// then x * 255 -> uint8.  OpenCV converts that last value in two different ways: its vectorised body (the first
// (W / lanes) * lanes pixels of every image row, lanes = 32 with AVX2) TRUNCATES, the scalar tail of the row rounds
// to nearest even -- `trunc` selects which.  Checked exhaustively against cv2 4.13 (2^24 RGB triples; 180 * 2^16
// HSV triples, both roundings).
__host__ __device__ __forceinline__ int k1_div_rhe(int n, int d) {  // round-half-even(n / d), n, d > 0
    const int q = n / d, r = n - q * d, t = 2 * r;
    return q + ((t > d || (t == d && (q & 1))) ? 1 : 0);
}
// OpenCV's division tables: sdiv_table[v] = cvRound((255 << 12) / v), hdiv_table180[d] = cvRound((180 << 12) / (6 d)).
__host__ __device__ __forceinline__ int k1_hsv_sdiv(int v) { return v ? k1_div_rhe(255 << 12, v) : 0; }
__host__ __device__ __forceinline__ int k1_hsv_hdiv(int d) { return d ? k1_div_rhe((180 << 12) / 6, d) : 0; }

#ifdef __CUDA_ARCH__
__device__ __forceinline__ float k1_selp(float a, float b, uint32_t cond) {   // cond != 0 ? a : b, never a branch
    float r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.f32 %0, %1, %2, p;\n\t}" : "=f"(r) : "f"(a), "f"(b), "r"(cond));
    return r;
}
__device__ __forceinline__ int k1_selp_i(int a, int b, uint32_t cond) {
    int r;
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\tselp.s32 %0, %1, %2, p;\n\t}" : "=r"(r) : "r"(a), "r"(b), "r"(cond));
    return r;
}
#endif

// `tab`: [512] = sdiv_table | hdiv_table180 (the kernel keeps them in shared memory); on the host NULL = compute them.
__host__ __device__ __forceinline__ void k1_rgb2hsv(int r, int g, int b, int& h, int& s, int& v, const int* tab) {
    v = r > g ? r : g; v = v > b ? v : b;
    int vmin = r < g ? r : g; vmin = vmin < b ? vmin : b;
    const int diff = v - vmin;
#ifdef __CUDA_ARCH__
    const int sd = tab[v], hd = tab[256 + diff];          // device callers always pass the tables
#else
    const int sd = tab ? tab[v] : k1_hsv_sdiv(v);
    const int hd = tab ? tab[256 + diff] : k1_hsv_hdiv(diff);
#endif
    s = (diff * sd + (1 << 11)) >> 12;
#ifdef __CUDA_ARCH__
    // (selects, not branches: neighbouring pixels have different maxima)
    int hh = k1_selp_i(g - b, k1_selp_i(b - r + 2 * diff, r - g + 4 * diff, (uint32_t)(v == g)), (uint32_t)(v == r));
#else
    int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * diff) : (r - g + 4 * diff));
#endif
    hh = (hh * hd + (1 << 11)) >> 12;   // arithmetic shift, as in OpenCV
    h = hh < 0 ? hh + 180 : hh;
}
__host__ __device__ __forceinline__ void k1_hsv2rgb(int h, int s, int v, bool trunc, uint32_t& r, uint32_t& g,
                                                    uint32_t& b) {
    const float k255 = 0x1.010102p-8f;   // 1.f / 255.f
    const float hscale = 0x1.111112p-5f; // 6.f / 180.f
#ifdef __CUDA_ARCH__
    const float sf = __fmul_rn((float)s, k255), vf = __fmul_rn((float)v, k255), hf = __fmul_rn((float)h, hscale);
    // trunc(hf) == h / 30 for every h in 0..255 (checked exhaustively), so the sector comes from the integer and its
    // float form from the 2^23 trick instead of FRND + F2I
    const uint32_t sec6 = (uint32_t)(h * 2185) >> 16;
    const float pre = __fsub_rn(__int_as_float(0x4B000000 | (int)sec6), 8388608.f), fr = __fsub_rn(hf, pre);
    const float t1 = __fmul_rn(vf, __fsub_rn(1.f, sf));
    const float t2 = __fmul_rn(vf, __fmaf_rn(-sf, fr, 1.f));
    const float t3 = __fmul_rn(vf, __fmaf_rn(-sf, __fsub_rn(1.f, fr), 1.f));
#else
    volatile float sf = (float)s * k255, vf = (float)v * k255, hf = (float)h * hscale;
    volatile float pre = truncf(hf);
    volatile float fr = hf - pre;
    volatile float oms = 1.f - sf, omf = 1.f - fr;
    volatile float t1 = vf * oms;
    volatile float i2 = fmaf(-sf, fr, 1.f), i3 = fmaf(-sf, omf, 1.f);
    volatile float t2 = vf * i2, t3 = vf * i3;
#endif
#ifdef __CUDA_ARCH__
    // sector_data = {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0} -> (b, g, r) from tab0..3 (= vf, t1, t2, t3).  Even
    // sectors use t3 and odd ones t2 as their "moving" value tv, and every channel is then one of {vf, t1, tv} chosen
    // by the sector alone: two selects per channel on bit tests of (1 << sector).  Written with selp so that the lanes
    // of a warp (which sit in different sectors) never diverge -- as nested ?: the compiler emits a branch tree that
    // every warp walks completely.
    const uint32_t m = 1u << (sec6 >= 6u ? sec6 - 6u : sec6);         // 1 << sector (a table may hand out h >= 180)
    const float tv = k1_selp(t2, t3, m & 0x2Au);                        // odd sectors: t2
    const float bb = k1_selp(t1, k1_selp(vf, tv, m & 0x18u), m & 0x03u);   // t1: 0,1   vf: 3,4   tv: 2,5
    const float gg = k1_selp(vf, k1_selp(t1, tv, m & 0x30u), m & 0x06u);   // vf: 1,2   t1: 4,5   tv: 0,3
    const float rr = k1_selp(vf, k1_selp(t1, tv, m & 0x0Cu), m & 0x21u);   // vf: 0,5   t1: 2,3   tv: 1,4
    // x * 255 -> uint8 without F2I: adding 2^23 leaves round(x) (ties to even, as cvRound) or, with round-toward-zero,
    // floor(x) in the low mantissa bits; x >= -1 here and the final clamp maps the -1 a tiny negative x truncates to
    // (2^23 - 0.5 -> ...FF) back to 0.
    const float r255 = __fmul_rn(rr, 255.f), g255 = __fmul_rn(gg, 255.f), b255 = __fmul_rn(bb, 255.f);
    const float big = 8388608.f;
    const int ri = (trunc ? __float_as_int(__fadd_rz(r255, big)) : __float_as_int(__fadd_rn(r255, big))) - 0x4B000000;
    const int gi = (trunc ? __float_as_int(__fadd_rz(g255, big)) : __float_as_int(__fadd_rn(g255, big))) - 0x4B000000;
    const int bi = (trunc ? __float_as_int(__fadd_rz(b255, big)) : __float_as_int(__fadd_rn(b255, big))) - 0x4B000000;
#else
    int sector = (int)pre;
    sector = sector >= 6 ? sector - 6 : sector;
    const float t0 = vf, u1 = t1, u2 = t2, u3 = t3;
    const float bb = sector <= 1 ? u1 : (sector == 2 ? u3 : (sector == 5 ? u2 : t0));
    const float gg = sector == 0 ? u3 : (sector <= 2 ? t0 : (sector == 3 ? u2 : u1));
    const float rr = (sector == 0 || sector == 5) ? t0 : (sector == 1 ? u2 : (sector == 4 ? u3 : u1));
    volatile float r255 = rr * 255.f, g255 = gg * 255.f, b255 = bb * 255.f;
    const int ri = trunc ? (int)r255 : (int)rintf(r255), gi = trunc ? (int)g255 : (int)rintf(g255),
              bi = trunc ? (int)b255 : (int)rintf(b255);
#endif
#ifdef __CUDA_ARCH__
    // every tab value is <= vf <= fl(255 * k255) = 1 + 2^-23, so x * 255 < 255.5 rounds to at most 255: only the lower
    // clamp (the -1 above) is needed
    r = (uint32_t)max(ri, 0);
    g = (uint32_t)max(gi, 0);
    b = (uint32_t)max(bi, 0);
#else
    r = (uint32_t)(ri < 0 ? 0 : (ri > 255 ? 255 : ri));
    g = (uint32_t)(gi < 0 ? 0 : (gi > 255 ? 255 : gi));
    b = (uint32_t)(bi < 0 ? 0 : (bi > 255 ? 255 : bi));
#endif
}
// lut: [3][256] = hue, sat, val tables of this sample
template <typename LoadByte>
__host__ __device__ __forceinline__ void k1_hsv_shift(uint32_t& r, uint32_t& g, uint32_t& b, bool trunc, LoadByte ld,
                                                      const int* div_tab = nullptr) {
    int h, s, v;
    k1_rgb2hsv((int)r, (int)g, (int)b, h, s, v, div_tab);
    k1_hsv2rgb((int)ld(h), (int)ld(256 + s), (int)ld(512 + v), trunc, r, g, b);
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

// (a - m) * d and (b - m) * d: sub.rn.f32x2 + mul.rn.f32x2 (SASS FADD2 / FMUL2, sm_100); the scalars broadcast.
__device__ __forceinline__ void normalize2(float a, float b, float m, float d, float& ra, float& rb) {
    unsigned long long x, mm, dd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1, %1};" : "=l"(dd) : "f"(d));
    asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(mm));
    asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(dd));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(ra), "=f"(rb) : "l"(x));
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
    if (g.bx0 + 1 >= g.fw) return false;   // a 1-pixel box on the frame's last column: its window shifts left of the span
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
