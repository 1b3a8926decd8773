/* CPU oracle, plain C restatement of the per-crop preprocessing path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): never linked into or
 * called from the shipped library.  Built by oracle/Makefile (and by
 * __graft_entry__.build()) into oracle/_build/liboracle_preprocess.so and
 * driven from tests/ and bench.py's cpu_baseline leg through ctypes.
 *
 * Follows:
 *   nkb_classification/dataset.py:398-409  (crop = frame[y0:y1, x0:x1])
 *   nkb_classification/dataset.py:96-102   (np.array(crop) -> A.Compose)
 *   configs/singletask_config.py:203-219   (LongestMaxSize + PadIfNeeded + Normalize + ToTensorV2)
 *   metrics/det_cls_val.py:86-109          (A.Resize | letterbox, Normalize, ToTensorV2)
 * and OpenCV's 8-bit INTER_LINEAR fixed-point arithmetic (imgproc resize.cpp;
 * spec in SURVEY.md section 9.1), albumentations 1.x Normalize (section 9.2) and
 * letterbox geometry (section 9.3).
 *
 * Compile without FMA contraction (-ffp-contract=off): the coordinate
 * (dx+0.5)*scale-0.5 must be a separately rounded multiply and subtract.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define COEF_SCALE 2048.0f

static void axis_tables(int dsize, int ssize, int horizontal, int *s, int *c0, int *c1) {
    double scale = 1.0 / ((double)dsize / (double)ssize);
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int si = (int)floorf(f);
        f -= (float)si;
        if (horizontal) {
            if (si < 0) { si = 0; f = 0.f; }
            if (si >= ssize - 1) { si = ssize - 1; f = 0.f; }
        }
        s[d] = si;
        c0[d] = (int)lrintf((1.f - f) * COEF_SCALE); /* round half even */
        c1[d] = (int)lrintf(f * COEF_SCALE);
    }
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Python-3 round(): half to even, on a double */
static long py3round(double x) { return lrint(x); /* default FE_TONEAREST = half-even */ }

/* src: crop top-left pointer, pitch in bytes, sh x sw pixels, 3 channels.
 * dst: dh x dw x 3 uint8, dst_pitch bytes. swap: read channel 2-c. */
static void resize_linear_u8(const uint8_t *src, long pitch, int sh, int sw, uint8_t *dst, long dst_pitch, int dh,
                             int dw, int swap) {
    int *sx = (int *)malloc(sizeof(int) * 3 * (size_t)dw);
    int *ax0 = sx + dw, *ax1 = sx + 2 * dw;
    int *sy = (int *)malloc(sizeof(int) * 3 * (size_t)dh);
    int *by0 = sy + dh, *by1 = sy + 2 * dh;
    axis_tables(dw, sw, 1, sx, ax0, ax1);
    axis_tables(dh, sh, 0, sy, by0, by1);
    for (int y = 0; y < dh; ++y) {
        int r0 = clampi(sy[y], 0, sh - 1), r1 = clampi(sy[y] + 1, 0, sh - 1);
        const uint8_t *p0 = src + (long)r0 * pitch, *p1 = src + (long)r1 * pitch;
        for (int x = 0; x < dw; ++x) {
            int xa = sx[x], xb = xa + 1 < sw ? xa + 1 : sw - 1;
            for (int c = 0; c < 3; ++c) {
                int cs = swap ? 2 - c : c;
                int h0 = p0[xa * 3 + cs] * ax0[x] + p0[xb * 3 + cs] * ax1[x];
                int h1 = p1[xa * 3 + cs] * ax0[x] + p1[xb * 3 + cs] * ax1[x];
                int v = (((by0[y] * (h0 >> 4)) >> 16) + ((by1[y] * (h1 >> 4)) >> 16) + 2) >> 2;
                dst[(long)y * dst_pitch + x * 3 + c] = (uint8_t)v;
            }
        }
    }
    free(sx);
    free(sy);
}

/* One crop.  mode 0 = stretch to out_h x out_w, 1 = letterbox (LongestMaxSize(max_size)
 * + centred constant pad).  out_u8: out_h*out_w*3 HWC (may be NULL);
 * out_f32: 3*out_h*out_w CHW (may be NULL).  Returns 0, or -1 on a bad box. */
int oracle_preprocess_crop(const uint8_t *frame, long pitch, int fh, int fw, const int *box, int mode, int out_h,
                           int out_w, int max_size, const uint8_t *pad, const float *mean255, const float *denom,
                           int swap, uint8_t *out_u8, float *out_f32) {
    int x0 = box[0], y0 = box[1], x1 = box[2], y1 = box[3];
    if (x0 < 0 || y0 < 0 || x1 > fw || y1 > fh || x1 <= x0 || y1 <= y0) return -1;
    int w = x1 - x0, h = y1 - y0;
    uint8_t *u8 = out_u8 ? out_u8 : (uint8_t *)malloc((size_t)out_h * out_w * 3);
    const uint8_t *src = frame + (long)y0 * pitch + (long)x0 * 3;
    if (mode == 0) {
        if (w == out_w && h == out_h) {
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x)
                    for (int c = 0; c < 3; ++c)
                        u8[((long)y * out_w + x) * 3 + c] = src[(long)y * pitch + x * 3 + (swap ? 2 - c : c)];
        } else {
            resize_linear_u8(src, pitch, h, w, u8, (long)out_w * 3, out_h, out_w, swap);
        }
    } else {
        int longest = w > h ? w : h;
        double scale = (double)max_size / (double)longest;
        int new_h = h, new_w = w;
        if (scale != 1.0) {
            new_h = (int)py3round(h * scale);
            new_w = (int)py3round(w * scale);
        }
        if (new_h > out_h || new_w > out_w || new_h < 1 || new_w < 1) {
            if (!out_u8) free(u8);
            return -2;
        }
        int top = (int)((out_h - new_h) / 2.0), left = (int)((out_w - new_w) / 2.0);
        for (long i = 0; i < (long)out_h * out_w; ++i)
            for (int c = 0; c < 3; ++c) u8[i * 3 + c] = pad[c];
        uint8_t *dst = u8 + ((long)top * out_w + left) * 3;
        if (new_h == h && new_w == w) {
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x)
                    for (int c = 0; c < 3; ++c)
                        dst[((long)y * out_w + x) * 3 + c] = src[(long)y * pitch + x * 3 + (swap ? 2 - c : c)];
        } else {
            resize_linear_u8(src, pitch, h, w, dst, (long)out_w * 3, new_h, new_w, swap);
        }
    }
    if (out_f32) {
        long plane = (long)out_h * out_w;
        for (long i = 0; i < plane; ++i)
            for (int c = 0; c < 3; ++c) {
                float v = (float)u8[i * 3 + c];
                v = v - mean255[c];
                v = v * denom[c];
                out_f32[c * plane + i] = v;
            }
    }
    if (!out_u8) free(u8);
    return 0;
}

/* A batch: frames described by (byte offset from base, h, w, pitch) rows. */
int oracle_preprocess_batch(const uint8_t *base, const int64_t *frame_desc, const int *boxes, const int *frame_idx,
                            int n, int mode, int out_h, int out_w, int max_size, const uint8_t *pad,
                            const float *mean255, const float *denom, int swap, uint8_t *out_u8, float *out_f32) {
    long plane3 = (long)out_h * out_w * 3;
    for (int i = 0; i < n; ++i) {
        const int64_t *fd = frame_desc + 4 * (long)frame_idx[i];
        int rc = oracle_preprocess_crop(base + fd[0], (long)fd[3], (int)fd[1], (int)fd[2], boxes + 4 * (long)i, mode,
                                        out_h, out_w, max_size, pad, mean255, denom, swap,
                                        out_u8 ? out_u8 + plane3 * i : NULL, out_f32 ? out_f32 + plane3 * i : NULL);
        if (rc) return rc;
    }
    return 0;
}
