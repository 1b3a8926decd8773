// The direct-load ("general") K1 band routine, shared by k1_general.cu (its own kernel: letterbox, partial
// tiles, uint8 side output) and k1_fast.cu (in-kernel fallback for crops the TMA path cannot take).
#pragma once
#include "k1_common.cuh"

namespace nkbk {

// Raw words of one source row for this lane's JMAX two-pixel windows, loaded one
// row ahead of use so the DRAM/L2 latency hides behind the vertical pass.
template <int JMAX>
struct RawRow {
    uint32_t w0[JMAX], w1[JMAX], w2[JMAX];
    uint32_t mis;  // (row address & 3): byte misalignment of the row the words came from
    int row;       // source row held, -1 = none
};

// Issue the loads of source row `rowp` (pointer to the crop's first byte of that row).
// Every word read contains at least one byte of the row's pixels, so aligned 32-bit
// loads never leave the mapped frame; the third word is fetched only when the
// 6-byte window straddles into it ((offset & 3) == 3).
template <int JMAX>
__device__ __forceinline__ void issue_row(RawRow<JMAX>& R, const uint8_t* __restrict__ rowp,
                                          const uint32_t (&xo)[JMAX], int row) {
    const uint32_t mis = uint32_t(reinterpret_cast<uintptr_t>(rowp)) & 3u;
    const uint8_t* base = rowp - mis;
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
        const uint32_t p = xo[j] + mis;
        const uint32_t* al = reinterpret_cast<const uint32_t*>(base + (p & ~3u));
        R.w0[j] = __ldg(al);
        R.w1[j] = __ldg(al + 1);
        R.w2[j] = 0u;
        if ((p & 3u) == 3u) R.w2[j] = __ldg(al + 2);
    }
    R.mis = mis;
    R.row = row;
}

// Horizontal pass: raw words -> H >> 4 (the 15-bit values OpenCV's vertical pass consumes).
template <int JMAX>
__device__ __forceinline__ void convert_row(uint32_t (&H)[JMAX][3], const RawRow<JMAX>& R, const uint32_t (&xo)[JMAX],
                                            const uint32_t (&cf)[JMAX], const uint32_t sel0, const uint32_t sel1,
                                            const uint32_t sel2) {
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
        const uint32_t k8 = (xo[j] + R.mis) << 3;                    // shf.wrap uses the low 5 bits: (offset & 3) * 8
        const uint32_t lo = __funnelshift_r(R.w0[j], R.w1[j], k8);   // bytes o .. o+3
        const uint32_t hi = __funnelshift_r(R.w1[j], R.w2[j], k8);   // bytes o+4 .. o+7
        const uint32_t t01 = __byte_perm(lo, hi, (sel0 & 0xFFu) | ((sel1 & 0xFFu) << 8));   // taps of channels 0 and 1
        H[j][0] = __dp2a_lo(cf[j], t01, 0u) >> 4;
        H[j][1] = __dp2a_hi(cf[j], t01, 0u) >> 4;
        H[j][2] = __dp2a_lo(cf[j], __byte_perm(lo, hi, sel2), 0u) >> 4;
    }
}

// GENERAL = false: stretch mode and out_w a multiple of the 32*JMAX column tile -> every lane
// writes every column, no border handling, no uint8 side output.  GENERAL = true: everything.
// One warp produces output rows [y_begin, y_begin + nrows) (nrows <= 32) of `crop`, columns ox0 + 32*j.
// `count_bad`: this warp is the one that reports an invalid box to bad_count.
// AUG = true (general variant only): the per-crop train-time augmentations of K1Params::aug_* are applied to the
// resized + padded uint8 pixel before Normalize; flips mirror the store address, so a band of source-order rows
// lands on the mirrored rows / columns of the output.
template <int JMAX, typename OutT, bool GENERAL, bool WRITE_U8, bool AUG = false>
__device__ __forceinline__ void k1_process_band(const K1Params& p, const int crop, const CropGeom& g, const int y_begin,
                                                const int nrows, const int ox0, const bool count_bad,
                                                const int* hsv_div_tab = nullptr, uint8_t* hsv_lut_smem = nullptr) {
    static_assert(GENERAL || !WRITE_U8, "uint8 side output only in the general variant");
    static_assert(GENERAL || !AUG, "augmentations only in the general variant");
    const int lane = threadIdx.x & 31;
    const int bx0 = g.bx0, by0 = g.by0, bw = g.bw, bh = g.bh, fw = g.fw;
    const int64_t f_off = g.f_off, pitch = g.pitch;
    bool ok = g.ok;
    int dw = p.out_w, dh = p.out_h, top = 0, left = 0;
    if (GENERAL && ok && p.mode == NKBK_MODE_LETTERBOX)
        ok = letterbox_geometry(bh, bw, p.max_size, p.out_h, p.out_w, dh, dw, top, left);

    const int64_t plane = (int64_t)p.out_h * p.out_w;
    OutT* const out_crop = reinterpret_cast<OutT*>(p.out) + (int64_t)crop * 3 * plane;
    uint32_t wmask = (1u << JMAX) - 1u;  // columns this lane writes at all
    if (GENERAL) {
        wmask = 0;
#pragma unroll
        for (int j = 0; j < JMAX; ++j) wmask |= uint32_t(ox0 + 32 * j < p.out_w) << j;
    }

    // ---- augmentation parameters of this crop ----
    int aflags = 0, nholes = 0;
    float a_alpha = 1.f, a_beta = 0.f;
    const int32_t* holes = nullptr;
    if (AUG) {
        aflags = __ldg(p.aug_flags + crop);
        nholes = min(aflags >> 8, p.aug_max_holes);
        if (aflags & K1_AUG_BC) { a_alpha = __ldg(p.aug_alpha + crop); a_beta = __ldg(p.aug_beta + crop); }
        holes = p.aug_holes + (int64_t)crop * p.aug_max_holes * 4;
    }
    const bool hflip = AUG && (aflags & K1_AUG_HFLIP), vflip = AUG && (aflags & K1_AUG_VFLIP);
    const int xstep = hflip ? -32 : 32;                       // destination column step between this lane's j's
    const int xd0 = hflip ? p.out_w - 1 - ox0 : ox0;          // destination column of j = 0
    // the colour ops on one pixel, in the reference's order: RandomBrightnessContrast, HueSaturationValue, then the
    // CoarseDropout fill.  px is in OUTPUT channel order (RGB after the optional BGR swap).
    // this crop's hue / sat / val tables: staged once per band in the warp's 768 bytes of shared memory
    const uint8_t* hsv_lut = nullptr;
    if (AUG && (aflags & K1_AUG_HSV)) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.aug_hsv_lut + (int64_t)crop * 768);
        uint32_t* dst = reinterpret_cast<uint32_t*>(hsv_lut_smem);
        __syncwarp();   // the previous crop's readers are done
#pragma unroll
        for (int i = 0; i < 6; ++i) dst[lane + 32 * i] = __ldg(src + lane + 32 * i);
        __syncwarp();
        hsv_lut = hsv_lut_smem;
    }
    auto augment3 = [&](uint32_t (&px)[3], bool in_hole, int xd) {   // xd: destination column of the pixel
        if (!AUG) return;
        if (aflags & K1_AUG_BC) {
#pragma unroll
            for (int c = 0; c < 3; ++c) px[c] = k1_brightness_contrast(px[c], a_alpha, a_beta);
        }
        if (hsv_lut != nullptr)
            k1_hsv_shift(px[0], px[1], px[2], xd < p.aug_hsv_trunc_cols,
                         [&](int i) { return (uint32_t)hsv_lut[i]; }, hsv_div_tab);
        if (in_hole) { px[0] = p.aug_fill[0]; px[1] = p.aug_fill[1]; px[2] = p.aug_fill[2]; }
    };
    const int trunc_cols = AUG ? p.aug_hsv_trunc_cols : 0;
    const uint32_t fill0 = p.aug_fill[0], fill1 = p.aug_fill[1], fill2 = p.aug_fill[2];
    auto augment_row = [&](uint32_t (&P)[JMAX][3], uint32_t hm) {
        if (aflags & K1_AUG_BC) {
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) P[j][c] = k1_brightness_contrast(P[j][c], a_alpha, a_beta);
        }
        if (hsv_lut != nullptr) {
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
                k1_hsv_shift(P[j][0], P[j][1], P[j][2], xd0 + xstep * j < trunc_cols,
                             [&](int i) { return (uint32_t)hsv_lut[i]; }, hsv_div_tab);
        }
        if (hm != 0u) {
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
                if (hm >> j & 1) { P[j][0] = fill0; P[j][1] = fill1; P[j][2] = fill2; }
        }
    };
    // bit j set: destination column of (lane, j) on destination row yd lies inside a CoarseDropout hole
    auto hole_mask = [&](int yd) -> uint32_t {
        uint32_t m = 0;
        if constexpr (AUG) {
            for (int h = 0; h < nholes; ++h) {
                const int x1 = __ldg(holes + 4 * h + 0), y1 = __ldg(holes + 4 * h + 1);
                const int x2 = __ldg(holes + 4 * h + 2), y2 = __ldg(holes + 4 * h + 3);
                if (yd < y1 || yd >= y2) continue;
#pragma unroll
                for (int j = 0; j < JMAX; ++j) {
                    const int xd = xd0 + xstep * j;
                    m |= uint32_t(xd >= x1 && xd < x2) << j;
                }
            }
        }
        return m;
    };

    if (!ok) {
        // empty / out-of-frame box: emit the normalised pad value, count it once per crop
        if (p.bad_count != nullptr && count_bad && lane == 0) atomicAdd(p.bad_count, 1);
        for (int yy = 0; yy < nrows; ++yy) {
            OutT* o = out_crop + (int64_t)(y_begin + yy) * p.out_w + ox0;
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
                if (wmask >> j & 1) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) store_out<OutT>(o + c * plane + 32 * j, p.padf[c]);
                    if (WRITE_U8) {
                        uint8_t* u = p.out_u8 + (((int64_t)crop * p.out_h + y_begin + yy) * p.out_w + ox0 + 32 * j) * 3;
                        u[0] = (uint8_t)p.padu[0]; u[1] = (uint8_t)p.padu[1]; u[2] = (uint8_t)p.padu[2];
                    }
                }
        }
        return;
    }

    // ---- horizontal tables, one entry per (lane, j), kept in registers ----
    uint32_t xo[JMAX], cf[JMAX];
    uint32_t vmask = (1u << JMAX) - 1u;  // columns that receive resized pixels (the rest of wmask is border)
    {
        if (GENERAL) vmask = 0;
        const double sxs = axis_scale(dw, bw);
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const int ox = ox0 + 32 * j;
            const int dx = ox - left;
            const bool v = !GENERAL || (ox < p.out_w && dx >= 0 && dx < dw);
            int s = 0, c0 = 0, c1 = 0;
            if (v) axis_coef(dx, sxs, bw, true, s, c0, c1);
            int px = bx0 + s;
            uint32_t c = uint32_t(c0) | (uint32_t(c1) << 16);
            if (px + 1 >= fw) {  // window would leave the frame row: shift it left, weight moves to tap 1
                px -= 1;
                c = uint32_t(c0) << 16;
            }
            if (!v) { px = 0; c = 0u; }
            xo[j] = uint32_t(px) * 3u;
            cf[j] = c;
            if (GENERAL) vmask |= uint32_t(v) << j;
        }
    }

    // ---- vertical tables: lane l holds row y_begin + l of this band ----
    int my_r0 = -1, my_r1 = -1;
    uint32_t my_b0 = 0, my_b1 = 0;
    if (lane < nrows) {
        const int dy = y_begin + lane - top;
        if (dy >= 0 && dy < dh) {
            int s, c0, c1;
            axis_coef(dy, axis_scale(dh, bh), bh, false, s, c0, c1);
            my_r0 = min(max(s, 0), bh - 1);
            my_r1 = min(max(s + 1, 0), bh - 1);
            my_b0 = uint32_t(c0) << 16;  // pre-shifted: umulhi(b << 16, h) == (b * h) >> 16
            my_b1 = uint32_t(c1) << 16;
        }
    }

    const uint8_t* const src0 = p.frames + f_off + (int64_t)by0 * pitch;
    const uint32_t sel0 = p.sel[0], sel1 = p.sel[1], sel2 = p.sel[2];
    const float m0 = p.m[0], m1 = p.m[1], m2 = p.m[2];
    const float d0 = p.d[0], d1 = p.d[1], d2 = p.d[2];

    uint32_t Ha[JMAX][3], Hb[JMAX][3];
    int ia = -1, ib = -1;
#pragma unroll
    for (int j = 0; j < JMAX; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) Ha[j][c] = Hb[j][c] = 0u;

    // ---- source-row stream with one row of look-ahead ----
    // The rows a band needs form a strictly increasing sequence (r0, r1 of each output row that
    // are not already held).  `fetch_next` scans the (row, tap) candidates in order of need and
    // issues the loads of the next unseen row; it is called right after the previous row has
    // been converted, i.e. a whole vertical pass before the data is consumed.
    RawRow<JMAX> pf;
    pf.row = -1;
    pf.mis = 0;
#pragma unroll
    for (int j = 0; j < JMAX; ++j) pf.w0[j] = pf.w1[j] = pf.w2[j] = 0u;
    int kq = 0, fetched_max = -1;
    auto fetch_next = [&]() {
        pf.row = -1;
        while (kq < 2 * nrows) {
            const int rc = __shfl_sync(0xffffffffu, (kq & 1) ? my_r1 : my_r0, kq >> 1);
            ++kq;
            if (rc > fetched_max) {
                fetched_max = rc;
                issue_row<JMAX>(pf, src0 + (int64_t)rc * pitch, xo, rc);
                break;
            }
        }
    };
    fetch_next();

    for (int yy = 0; yy < nrows; ++yy) {
        const int r0 = __shfl_sync(0xffffffffu, my_r0, yy);
        const int r1 = __shfl_sync(0xffffffffu, my_r1, yy);
        const uint32_t b0 = __shfl_sync(0xffffffffu, my_b0, yy);
        const uint32_t b1 = __shfl_sync(0xffffffffu, my_b1, yy);
        const int y = AUG && vflip ? p.out_h - 1 - (y_begin + yy) : y_begin + yy;   // destination row
        OutT* o = out_crop + (int64_t)y * p.out_w + xd0;
        uint8_t* u = WRITE_U8 ? p.out_u8 + (((int64_t)crop * p.out_h + y) * p.out_w + xd0) * 3 : nullptr;
        const uint32_t hmask = hole_mask(y);

        if (GENERAL && r0 < 0) {  // letterbox border row
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
                if (wmask >> j & 1) {
                    if (AUG) {
                        const float mm[3] = {m0, m1, m2}, dd[3] = {d0, d1, d2};
                        uint32_t pv[3] = {p.padu[0], p.padu[1], p.padu[2]};
                        augment3(pv, hmask >> j & 1, xd0 + xstep * j);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            store_out<OutT>(o + c * plane + xstep * j, __fmul_rn(__fsub_rn((float)pv[c], mm[c]), dd[c]));
                            if (WRITE_U8) u[3 * xstep * j + c] = (uint8_t)pv[c];
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; ++c) store_out<OutT>(o + c * plane + 32 * j, p.padf[c]);
                        if (WRITE_U8) {
                            u[96 * j + 0] = (uint8_t)p.padu[0]; u[96 * j + 1] = (uint8_t)p.padu[1];
                            u[96 * j + 2] = (uint8_t)p.padu[2];
                        }
                    }
                }
            continue;
        }

        if (r0 != ia) {
            if (r0 == ib) {
#pragma unroll
                for (int j = 0; j < JMAX; ++j)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Ha[j][c] = Hb[j][c];
            } else {  // pf.row == r0 by construction of the stream
                convert_row<JMAX>(Ha, pf, xo, cf, sel0, sel1, sel2);
                fetch_next();
            }
            ia = r0;
        }
        if (r1 != ib) {
            if (r1 == ia) {
#pragma unroll
                for (int j = 0; j < JMAX; ++j)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Hb[j][c] = Ha[j][c];
            } else {
                convert_row<JMAX>(Hb, pf, xo, cf, sel0, sel1, sel2);
                fetch_next();
            }
            ib = r1;
        }

        // the row's pixels first, then each colour op over the whole row under ONE warp-uniform test (a test per
        // element costs a BSSY / BRA / BSYNC triple each), then Normalize + stores
        uint32_t P[JMAX][3];
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t t0 = __umulhi(b0, Ha[j][c]);
                const uint32_t t1 = __umulhi(b1, Hb[j][c]);
                P[j][c] = (t0 + t1 + 2u) >> 2;
                if (GENERAL && !(vmask >> j & 1)) P[j][c] = p.padu[c];
            }
        }
        if (AUG) augment_row(P, hmask);
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const float f0 = __fmul_rn(__fsub_rn((float)P[j][0], m0), d0);
            const float f1 = __fmul_rn(__fsub_rn((float)P[j][1], m1), d1);
            const float f2 = __fmul_rn(__fsub_rn((float)P[j][2], m2), d2);
            if (!GENERAL || (wmask >> j & 1)) {
                const int xs = AUG ? xstep * j : 32 * j;
                store_out<OutT>(o + xs, f0);
                store_out<OutT>(o + plane + xs, f1);
                store_out<OutT>(o + 2 * plane + xs, f2);
                if (WRITE_U8) {
                    u[3 * xs + 0] = (uint8_t)P[j][0]; u[3 * xs + 1] = (uint8_t)P[j][1]; u[3 * xs + 2] = (uint8_t)P[j][2];
                }
            }
        }
    }
}

}  // namespace nkbk
