// Per-axis source index + 11-bit coefficients of OpenCV's 8-bit INTER_LINEAR,
// and the albumentations letterbox geometry -- shared by the K1 kernel prologue
// (device) and the host-only nkbk_debug_* exports used by the CPU parity tests.
//
// Arithmetic contract: the coordinate is evaluated in
// double as (d + 0.5) * scale - 0.5 with scale = 1 / (dsize / ssize) (two
// roundings), rounded to float32; the fraction is a float32 subtraction; the
// horizontal clamp zeroes the fraction, the vertical one keeps it; the
// coefficients are rint(c * 2048).  No step may be contracted into an FMA, so
// the device side spells every operation with a round-to-nearest intrinsic.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define NKBK_HD __host__ __device__ __forceinline__
#else
#define NKBK_HD inline
#endif

namespace nkbk {

NKBK_HD double axis_scale(int dsize, int ssize) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(1.0, __ddiv_rn((double)dsize, (double)ssize));
#else
    volatile double inv = (double)dsize / (double)ssize;
    return 1.0 / inv;
#endif
}

// d: destination index; returns source index s and coefficients c0 (for s) and
// c1 (for s + 1).  horizontal: clamp s into [0, ssize-1] and zero the fraction.
NKBK_HD void axis_coef(int d, double scale, int ssize, bool horizontal, int& s, int& c0, int& c1) {
#if defined(__CUDA_ARCH__)
    double t = __dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    float f = __double2float_rn(t);
    float fl = floorf(f);
    int si = (int)fl;
    f = __fsub_rn(f, fl);
#else
    volatile double t0 = ((double)d + 0.5) * scale;
    volatile double t = t0 - 0.5;
    volatile float f0 = (float)t;
    float fl = floorf(f0);
    int si = (int)fl;
    volatile float f1 = f0 - fl;
    float f = f1;
#endif
    if (horizontal) {
        if (si < 0) { si = 0; f = 0.f; }
        if (si >= ssize - 1) { si = ssize - 1; f = 0.f; }
    }
#if defined(__CUDA_ARCH__)
    c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    c1 = __float2int_rn(__fmul_rn(f, 2048.f));
#else
    volatile float omf = 1.f - f;
    c0 = (int)lrintf(omf * 2048.f);
    c1 = (int)lrintf(f * 2048.f);
#endif
    s = si;
}

// LongestMaxSize(max_size) + centred PadIfNeeded(out_h, out_w).
// Returns false when the resized crop would not fit the output canvas.
NKBK_HD bool letterbox_geometry(int h, int w, int max_size, int out_h, int out_w, int& new_h, int& new_w, int& top,
                                int& left) {
    int longest = h > w ? h : w;
#if defined(__CUDA_ARCH__)
    double scale = __ddiv_rn((double)max_size, (double)longest);
#else
    volatile double scale = (double)max_size / (double)longest;
#endif
    new_h = h;
    new_w = w;
    if (scale != 1.0) {
#if defined(__CUDA_ARCH__)
        new_h = (int)rint(__dmul_rn((double)h, scale));  // round half even == Python 3 round()
        new_w = (int)rint(__dmul_rn((double)w, scale));
#else
        volatile double ph = (double)h * scale, pw = (double)w * scale;
        new_h = (int)rint(ph);
        new_w = (int)rint(pw);
#endif
    }
    if (new_h < 1 || new_w < 1 || new_h > out_h || new_w > out_w) return false;
    top = (int)((double)(out_h - new_h) / 2.0);   // int() truncation of a non-negative value
    left = (int)((double)(out_w - new_w) / 2.0);
    return true;
}

}  // namespace nkbk
