// K1: fused YOLO-box crop + cv2-INTER_LINEAR-exact uint8 resize + Normalize +
// HWC->NCHW, one launch per batch.  See include/nkbk.h for the contract and
// DESIGN.md section 4 "K1" for the layout / roofline reasoning.
//
// This file holds the C-ABI entry point and the GENERAL kernel (letterbox,
// partial column tiles, uint8 side output).  The band routine both kernels
// share is in k1_general_impl.cuh; the TMA-staged fast kernel, which serves
// A.Resize pipelines with full column tiles, is in k1_fast.cu.
//
// Work decomposition (both kernels):
//   grid  = (crop, band-of-rows block, column tile)       block = 4 warps
//   warp  = one band of consecutive output rows of one crop
//   lane  = output columns  tile*32*JMAX + lane + 32*j,  j = 0..JMAX-1
//           -> every store instruction of a warp writes one full 128-byte line
//              of one channel plane (fp32), every source load instruction of a
//              warp touches one contiguous ~32*scale*3-byte span of one row.
// Per lane the horizontal tables (byte offset of the 2-pixel window, packed
// 11-bit coefficient pair) live in registers for the whole band.  The warp
// walks down its rows keeping the horizontally-filtered values (H >> 4, the
// 15-bit quantity OpenCV feeds its vertical pass) of the two source rows in
// registers, so a source row is fetched and filtered once per band even when
// consecutive output rows share it (up-scaling) or it moves from the lower
// to the upper tap (down-scaling by < 2).
//
// Integer pipeline per output sample (bit-exact with cv2):
//   6 interleaved source bytes (2 pixels) are cut out of 2-3 aligned 32-bit
//   words with two funnel shifts, one PRMT per channel pairs the two taps, one
//   IDP.2A (16-bit x 8-bit dot product) yields H = a0*p0 + a1*p1, >> 4.
//   Vertical: IMAD.HI with the coefficient pre-shifted by 16 gives
//   (b*(H>>4))>>16 per tap, +2, >> 2, int->float, then (v - mean255) * denom
//   as two separately rounded fp32 operations.
#include "k1_general_impl.cuh"
#include <stdlib.h>

namespace nkbk {

bool launch_k1_fast(const K1Params& p, int jmax, dim3 grid, cudaStream_t st, bool f32);      // k1_fast.cu
bool launch_k1_fast_aug(const K1Params& p, int jmax, dim3 grid, cudaStream_t st, bool f32);  // k1_fast_aug.cu

// One crop per blockIdx.x, one band of rows per warp.
template <int JMAX, typename OutT, bool GENERAL, bool WRITE_U8, bool AUG = false>
__global__ void __launch_bounds__(K1_WARPS * 32) k1_crop_resize_normalize(const K1Params p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // HueSaturationValue: OpenCV's two division tables, built once per CTA (4 integer divisions per thread)
    __shared__ int hsv_div_tab[AUG ? 512 : 1];
    __shared__ __align__(16) uint8_t hsv_lut_s[AUG ? K1_WARPS * 768 : 16];   // per warp: the current crop's 3 x 256 tables
    if (AUG && p.aug_hsv_lut != nullptr) {
        for (int i = threadIdx.x; i < 256; i += K1_WARPS * 32) {
            hsv_div_tab[i] = k1_hsv_sdiv(i);
            hsv_div_tab[256 + i] = k1_hsv_hdiv(i);
        }
        __syncthreads();
    }
    const int y_begin = (blockIdx.y * K1_WARPS + warp) * p.rows_per_warp;
    if (y_begin >= p.out_h) return;
    const int nrows = min(p.rows_per_warp, p.out_h - y_begin);
    const int ox0 = blockIdx.z * (32 * JMAX) + lane;
    for (int crop = blockIdx.x; crop < p.n; crop += gridDim.x) {
        const CropGeom g = load_geom(p, crop);
        k1_process_band<JMAX, OutT, GENERAL, WRITE_U8, AUG>(p, crop, g, y_begin, nrows, ox0,
                                                            blockIdx.y == 0 && blockIdx.z == 0 && warp == 0,
                                                            AUG ? hsv_div_tab : nullptr,
                                                            AUG ? hsv_lut_s + warp * 768 : nullptr);
    }
}

template <int JMAX, typename OutT>
static void launch_k1(const K1Params& p, dim3 grid, cudaStream_t st, bool general) {
    (void)general;
    if (p.aug_flags != nullptr) {   // train-time pipeline: flips / brightness-contrast / dropout holes fused in
        if (p.out_u8 != nullptr)
            k1_crop_resize_normalize<JMAX, OutT, true, true, true><<<grid, K1_WARPS * 32, 0, st>>>(p);
        else
            k1_crop_resize_normalize<JMAX, OutT, true, false, true><<<grid, K1_WARPS * 32, 0, st>>>(p);
        return;
    }
    if (p.out_u8 != nullptr)
        k1_crop_resize_normalize<JMAX, OutT, true, true><<<grid, K1_WARPS * 32, 0, st>>>(p);
    else
        k1_crop_resize_normalize<JMAX, OutT, true, false><<<grid, K1_WARPS * 32, 0, st>>>(p);
}

// albumentations' hue / sat / val tables from the per-sample shifts (float64, as numpy evaluates `ramp + shift`)
__global__ void k1_build_hsv_luts(const double* __restrict__ shift, const int32_t* __restrict__ flags, int n,
                                  uint8_t* __restrict__ lut) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // one thread per table entry
    if (i >= (int64_t)n * 768) return;
    const int s = (int)(i / 768), e = (int)(i - (int64_t)s * 768), c = e >> 8, v = e & 255;
    uint8_t out = (uint8_t)v;
    const double sh = shift[3 * s + c];
    if ((flags[s] & K1_AUG_HSV) && sh != 0.0) {
        double t = __dadd_rn((double)v, sh);
        if (c == 0) {   // numpy's float mod: fmod, then the divisor's sign
            t = fmod(t, 180.0);
            if (t != 0.0 && t < 0.0) t = __dadd_rn(t, 180.0);
        } else {
            t = t < 0.0 ? 0.0 : (t > 255.0 ? 255.0 : t);
        }
        out = (uint8_t)(int)t;   // truncation, as ndarray.astype(uint8) of an in-range value
    }
    lut[i] = out;
}

struct K1AugArgs {
    const int32_t* flags;
    const float* alpha;
    const float* beta;
    const int32_t* holes;
    int max_holes;
    const uint8_t* fill;     // host [3]
    const uint8_t* hsv_lut;  // device [n][3][256] or NULL
    int hsv_trunc_cols;
};

}  // namespace nkbk

using namespace nkbk;

static int k1_preprocess_impl(const void* frames_base, const int64_t* frame_desc, int n_frames, const int32_t* boxes,
                              const int32_t* frame_idx, int n, int mode, int out_h, int out_w, int max_size,
                              const uint8_t* pad_value, const float* mean255, const float* denom, int channel_swap,
                              void* out, int out_dtype, uint8_t* out_u8, int32_t* bad_count, const K1AugArgs* aug,
                              void* stream);

// debug timeline of the last TMA-kernel launch made with NKBK_K1_TIMING set (profiles/tools/k1_timeline.py)
static unsigned long long* g_k1_timing = nullptr;
static size_t g_k1_timing_cap = 0, g_k1_timing_n = 0;

extern "C" int64_t nkbk_debug_k1_timeline(uint64_t* out_host, int64_t max_ctas) {
    if (g_k1_timing == nullptr || out_host == nullptr) return 0;
    const int64_t n = (int64_t)g_k1_timing_n < max_ctas ? (int64_t)g_k1_timing_n : max_ctas;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(out_host, g_k1_timing, (size_t)n * 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    return n;
}

extern "C" int nkbk_preprocess_crops(const void* frames_base, const int64_t* frame_desc, int n_frames,
                                     const int32_t* boxes, const int32_t* frame_idx, int n, int mode, int out_h,
                                     int out_w, int max_size, const uint8_t* pad_value, const float* mean255,
                                     const float* denom, int channel_swap, void* out, int out_dtype, uint8_t* out_u8,
                                     int32_t* bad_count, void* stream) {
    return k1_preprocess_impl(frames_base, frame_desc, n_frames, boxes, frame_idx, n, mode, out_h, out_w, max_size,
                              pad_value, mean255, denom, channel_swap, out, out_dtype, out_u8, bad_count, nullptr,
                              stream);
}

extern "C" int nkbk_preprocess_crops_aug(const void* frames_base, const int64_t* frame_desc, int n_frames,
                                         const int32_t* boxes, const int32_t* frame_idx, int n, int mode, int out_h,
                                         int out_w, int max_size, const uint8_t* pad_value, const float* mean255,
                                         const float* denom, int channel_swap, const int32_t* aug_flags,
                                         const float* aug_alpha, const float* aug_beta, const int32_t* aug_holes,
                                         int max_holes, const uint8_t* hole_fill, const uint8_t* aug_hsv_lut,
                                         int hsv_trunc_cols, void* out, int out_dtype, uint8_t* out_u8,
                                         int32_t* bad_count, void* stream) {
    NKBK_CHECK_ARG(n <= 0 || aug_flags != nullptr, "nkbk_preprocess_crops_aug: NULL aug_flags");
    NKBK_CHECK_ARG(n <= 0 || (aug_alpha != nullptr && aug_beta != nullptr), "nkbk_preprocess_crops_aug: NULL aug_alpha / aug_beta");
    NKBK_CHECK_ARG(max_holes >= 0 && max_holes <= K1_AUG_MAX_HOLES, "nkbk_preprocess_crops_aug: max_holes=%d outside [0,%d]",
                   max_holes, K1_AUG_MAX_HOLES);
    NKBK_CHECK_ARG(max_holes == 0 || n <= 0 || aug_holes != nullptr, "nkbk_preprocess_crops_aug: NULL aug_holes");
    K1AugArgs a{aug_flags, aug_alpha, aug_beta, aug_holes, max_holes, hole_fill, aug_hsv_lut, hsv_trunc_cols};
    return k1_preprocess_impl(frames_base, frame_desc, n_frames, boxes, frame_idx, n, mode, out_h, out_w, max_size,
                              pad_value, mean255, denom, channel_swap, out, out_dtype, out_u8, bad_count, &a, stream);
}

extern "C" int nkbk_build_hsv_luts(const double* hsv_shift, const int32_t* aug_flags, int n, uint8_t* out_lut,
                                   void* stream) {
    NKBK_CHECK_ARG(n >= 0, "nkbk_build_hsv_luts: n=%d < 0", n);
    if (n == 0) return NKBK_OK;
    NKBK_CHECK_ARG(hsv_shift && aug_flags && out_lut, "nkbk_build_hsv_luts: NULL pointer");
    const int64_t total = (int64_t)n * 768;
    k1_build_hsv_luts<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(hsv_shift, aug_flags,
                                                                                                  n, out_lut);
    NKBK_CHECK_LAUNCH("k1_build_hsv_luts");
    return NKBK_OK;
}

static int k1_preprocess_impl(const void* frames_base, const int64_t* frame_desc, int n_frames, const int32_t* boxes,
                              const int32_t* frame_idx, int n, int mode, int out_h, int out_w, int max_size,
                              const uint8_t* pad_value, const float* mean255, const float* denom, int channel_swap,
                              void* out, int out_dtype, uint8_t* out_u8, int32_t* bad_count, const K1AugArgs* aug,
                              void* stream) {
    NKBK_CHECK_ARG(n >= 0, "nkbk_preprocess_crops: n=%d < 0", n);
    if (n == 0) return NKBK_OK;
    NKBK_CHECK_ARG(frames_base && frame_desc && boxes && frame_idx && out, "nkbk_preprocess_crops: NULL pointer");
    NKBK_CHECK_ARG(mean255 && denom, "nkbk_preprocess_crops: NULL mean255/denom");
    NKBK_CHECK_ARG(n_frames >= 1, "nkbk_preprocess_crops: n_frames=%d", n_frames);
    NKBK_CHECK_ARG(mode == NKBK_MODE_STRETCH || mode == NKBK_MODE_LETTERBOX, "nkbk_preprocess_crops: mode=%d", mode);
    NKBK_CHECK_ARG(out_dtype == NKBK_F32 || out_dtype == NKBK_BF16, "nkbk_preprocess_crops: out_dtype=%d", out_dtype);
    if (out_h < 1 || out_w < 1 || out_h > 16384 || out_w > 16384) {
        set_error("nkbk_preprocess_crops: output %dx%d outside [1,16384]", out_h, out_w);
        return NKBK_E_SHAPE;
    }
    if (mode == NKBK_MODE_LETTERBOX && (max_size < 1 || max_size > out_h || max_size > out_w)) {
        set_error("nkbk_preprocess_crops: LongestMaxSize(%d) does not fit PadIfNeeded(%d,%d)", max_size, out_h, out_w);
        return NKBK_E_UNSUPPORTED;
    }

    K1Params p;
    p.frames = static_cast<const uint8_t*>(frames_base);
    p.frame_desc = frame_desc;
    p.boxes = boxes;
    p.frame_idx = frame_idx;
    p.n = n; p.n_frames = n_frames; p.mode = mode; p.out_h = out_h; p.out_w = out_w; p.max_size = max_size;
    static const uint32_t kSel[3] = {0x0030u, 0x0041u, 0x0052u};  // (tap0, tap1) byte pairs of R, G, B
    for (int c = 0; c < 3; ++c) {
        p.m[c] = mean255[c];
        p.d[c] = denom[c];
        p.padu[c] = pad_value ? pad_value[c] : 0u;
        volatile float t = (float)p.padu[c] - mean255[c];  // two rounded ops, like the kernel
        p.padf[c] = t * denom[c];
        p.sel[c] = kSel[channel_swap ? 2 - c : c];
    }
    p.out = out; p.out_u8 = out_u8; p.bad_count = bad_count;
    p.aug_flags = nullptr; p.aug_alpha = nullptr; p.aug_beta = nullptr; p.aug_holes = nullptr; p.aug_max_holes = 0;
    p.aug_fill[0] = p.aug_fill[1] = p.aug_fill[2] = 0u;
    p.aug_hsv_lut = nullptr;
    p.aug_hsv_trunc_cols = 0;
    if (aug != nullptr) {
        p.aug_flags = aug->flags; p.aug_alpha = aug->alpha; p.aug_beta = aug->beta; p.aug_holes = aug->holes;
        p.aug_max_holes = aug->max_holes;
        p.aug_hsv_lut = aug->hsv_lut;
        p.aug_hsv_trunc_cols = aug->hsv_trunc_cols;
        for (int c = 0; c < 3; ++c) p.aug_fill[c] = aug->fill ? aug->fill[c] : 0u;
    }

    // rows per warp: split the height into blocks of <= 64 rows, 4 warps each
    const int nby = (out_h + 63) / 64;
    p.rows_per_warp = (out_h + nby * K1_WARPS - 1) / (nby * K1_WARPS);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool f32 = out_dtype == NKBK_F32;
    p.rows_per_warp_fast = p.rows_per_warp;
    p.fby_fast = 1;
    p.timing = nullptr;

    // ---- fast path: A.Resize or LongestMaxSize + PadIfNeeded, output width a whole number of 32*J column tiles,
    // no uint8 side output; the train-time augmentations have their own instantiation (k1_fast_aug.cu) ----
    // Crops whose frame rows are not 16-byte aligned or whose boxes are too wide for the shared-memory ring are
    // produced, per CTA, by the direct-load band routine inside the same launch.
    static const bool aug_direct = getenv("NKBK_K1_AUG_DIRECT") != nullptr;   // experiment knob: augmentations on the direct-load kernel
    if (out_u8 == nullptr && out_w % 32 == 0 && (aug == nullptr || !aug_direct)) {
        const int cols = out_w / 32;
        int fj = 0;
        for (int j = 8; j >= 4; --j)
            if (cols % j == 0) { fj = j; break; }
        // the TMA kernel amortises its per-band prologue over up to 32 rows per warp (128 per CTA)
        // ... but small batches need the parallelism more.  Measured (rows per warp 32 / 24 / 16 / 12 / 8 / 6, K1 in us):
        // 512 crops 83.0 / 79.9 / 78.4 / 80.0 / 80.6 / 86.5; 1024 crops 144.9 / 142.0 / 141.9 / 146.7 / 150.3 / 163.9;
        // 256 images of 256^2 47.1 / 47.5 / 41.7 / 45.9 / 45.8 / 48.4 -- 16 rows until there are ~4 bands per resident
        // warp, 8 rows only for batches that cannot even fill half the warp slots with 16-row bands.
        const int64_t tiles = (int64_t)n * (cols / (fj ? fj : 1));
        int rmax = 32;
        if (tiles * ((out_h + 31) / 32) < 148 * 16 * 4) rmax = 16;
        if (tiles * ((out_h + 15) / 16) < 148 * 16 / 2) rmax = 8;
        if (const char* e = getenv("NKBK_K1_RMAX")) {   // experiment knob: rows per warp of the TMA kernel (4 .. 32)
            const int v = atoi(e);
            if (v >= 4 && v <= 32) rmax = v;
        }
        const int fby = (out_h + rmax * K1_WARPS - 1) / (rmax * K1_WARPS);
        p.rows_per_warp_fast = (out_h + fby * K1_WARPS - 1) / (fby * K1_WARPS);
        p.fby_fast = fby;
        if (fj != 0 && cols / fj <= 65535 && (int64_t)n * fby < (int64_t(1) << 31)) {
            dim3 fgrid((unsigned)((int64_t)n * fby), 1u, (unsigned)(cols / fj));
            static const bool want_timing = getenv("NKBK_K1_TIMING") != nullptr;   // debug: per-CTA timeline of the launch
            if (want_timing) {
                const size_t ctas = (size_t)fgrid.x * fgrid.z;
                if (g_k1_timing_cap < ctas) {
                    if (g_k1_timing) cudaFree(g_k1_timing);
                    if (cudaMalloc(&g_k1_timing, ctas * 3 * sizeof(unsigned long long)) != cudaSuccess) g_k1_timing = nullptr;
                    g_k1_timing_cap = g_k1_timing ? ctas : 0;
                }
                if (g_k1_timing) {
                    cudaMemsetAsync(g_k1_timing, 0, ctas * 3 * sizeof(unsigned long long), st);
                    p.timing = g_k1_timing;
                    g_k1_timing_n = ctas;
                }
            }
            if (aug != nullptr ? launch_k1_fast_aug(p, fj, fgrid, st, f32) : launch_k1_fast(p, fj, fgrid, st, f32)) {
                NKBK_CHECK_LAUNCH("k1_crop_resize_normalize_tma");
                return NKBK_OK;   // crops the TMA path cannot take are produced in-kernel by the direct-load routine
            }
        }
    }

    // ---- general path: letterbox, partial column tiles, uint8 side output ----
    // column tile: 32*JMAX columns, JMAX in {4,7,8}; least padded wins, ties -> wider
    int best_j = 8, best_cost = 1 << 30;
    const int cands[3] = {8, 7, 4};
    for (int i = 0; i < 3; ++i) {
        const int w = 32 * cands[i];
        const int cost = (out_w + w - 1) / w * w;
        if (cost < best_cost) { best_cost = cost; best_j = cands[i]; }
    }
    const int ntx = (out_w + 32 * best_j - 1) / (32 * best_j);
    if (nby > 65535 || ntx > 65535) {
        set_error("nkbk_preprocess_crops: output %dx%d too large", out_h, out_w);
        return NKBK_E_SHAPE;
    }
    dim3 grid((unsigned)n, (unsigned)nby, (unsigned)ntx);
    const bool general = true;
    switch (best_j) {
        case 4: f32 ? launch_k1<4, float>(p, grid, st, general) : launch_k1<4, __nv_bfloat16>(p, grid, st, general); break;
        case 7: f32 ? launch_k1<7, float>(p, grid, st, general) : launch_k1<7, __nv_bfloat16>(p, grid, st, general); break;
        default: f32 ? launch_k1<8, float>(p, grid, st, general) : launch_k1<8, __nv_bfloat16>(p, grid, st, general); break;
    }
    NKBK_CHECK_LAUNCH("k1_crop_resize_normalize");
    return NKBK_OK;
}

extern "C" int nkbk_debug_axis_table(int dsize, int ssize, int horizontal, int32_t* src_index, int32_t* coef0,
                                     int32_t* coef1) {
    NKBK_CHECK_ARG(dsize >= 1 && ssize >= 1 && src_index && coef0 && coef1, "nkbk_debug_axis_table: bad argument");
    const double sc = axis_scale(dsize, ssize);
    for (int d = 0; d < dsize; ++d) {
        int s, c0, c1;
        axis_coef(d, sc, ssize, horizontal != 0, s, c0, c1);
        src_index[d] = s; coef0[d] = c0; coef1[d] = c1;
    }
    return NKBK_OK;
}

extern "C" int nkbk_debug_letterbox(int h, int w, int max_size, int out_h, int out_w, int32_t* out4) {
    NKBK_CHECK_ARG(h >= 1 && w >= 1 && max_size >= 1 && out4, "nkbk_debug_letterbox: bad argument");
    int nh, nw, top = 0, left = 0;
    if (!letterbox_geometry(h, w, max_size, out_h, out_w, nh, nw, top, left)) {
        set_error("nkbk_debug_letterbox: %dx%d at max_size %d does not fit %dx%d", h, w, max_size, out_h, out_w);
        return NKBK_E_UNSUPPORTED;
    }
    out4[0] = nh; out4[1] = nw; out4[2] = top; out4[3] = left;
    return NKBK_OK;
}

extern "C" int nkbk_debug_brightness_contrast_lut(float alpha, float beta, uint8_t* lut256) {
    NKBK_CHECK_ARG(lut256 != nullptr, "nkbk_debug_brightness_contrast_lut: NULL output");
    for (int v = 0; v < 256; ++v) lut256[v] = (uint8_t)k1_brightness_contrast((uint32_t)v, alpha, beta);
    return NKBK_OK;
}

extern "C" int nkbk_debug_hsv_shift(const uint8_t* rgb_in, int64_t n, const uint8_t* lut768, int trunc,
                                    uint8_t* rgb_out) {
    NKBK_CHECK_ARG(n >= 0 && (n == 0 || (rgb_in && lut768 && rgb_out)), "nkbk_debug_hsv_shift: bad argument");
    for (int64_t i = 0; i < n; ++i) {
        uint32_t r = rgb_in[3 * i], g = rgb_in[3 * i + 1], b = rgb_in[3 * i + 2];
        k1_hsv_shift(r, g, b, trunc != 0, [&](int k) { return (uint32_t)lut768[k]; }, nullptr);
        rgb_out[3 * i] = (uint8_t)r; rgb_out[3 * i + 1] = (uint8_t)g; rgb_out[3 * i + 2] = (uint8_t)b;
    }
    return NKBK_OK;
}
