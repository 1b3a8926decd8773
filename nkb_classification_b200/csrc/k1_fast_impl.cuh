// The TMA-staged K1 kernel template (see k1_fast.cu for the description); instantiated without the train-time
// augmentations in k1_fast.cu and with them in k1_fast_aug.cu.
#pragma once
#include "k1_general_impl.cuh"

namespace nkbk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Source rows are re-read by overlapping boxes of the same frame (and by the neighbouring band of the same box):
// keep them in L2 (evict_last) while the 3x larger output stream goes through with evict-first stores.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "NKBK_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NKBK_DONE_%=;\n\t"
        "bra NKBK_WAIT_%=;\n\t"
        "NKBK_DONE_%=:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// (((b0 * h0) >> 16) + ((b1 * h1) >> 16) + 2) >> 2 with b pre-shifted by 16.  (Folding the "+2" and the first tap into
// the IMAD.HI addends through mad.hi.u32 was tried: the 64-bit addend pairs cost more register moves than the adds
// they save -- 215 vs 177 instructions per output row.)
__device__ __forceinline__ uint32_t vtap(uint32_t b0, uint32_t h0, uint32_t b1, uint32_t h1) {
    return (__umulhi(b0, h0) + __umulhi(b1, h1) + 2u) >> 2;
}

// (The ">> 4" of the horizontal pass was tried as mul.hi.u32 x, 2^28 -- IMAD.HI on the FMA pipe instead of SHF on the more
// heavily loaded ALU pipe: K1 went from 497 to 538 us, bf16 from 508 to 550 us.  IMAD.HI costs more than a shift.)

#ifndef K1F_MIN_BLOCKS
#define K1F_MIN_BLOCKS 4
#endif
#ifndef K1F_MIN_BLOCKS_AUG
#define K1F_MIN_BLOCKS_AUG 3   // 168 registers, no spills (4 blocks -> 128 registers spill in the colour ops: 1.66 ms vs the direct kernel's 1.38)
#endif

// LB = true: A.LongestMaxSize + centred A.PadIfNeeded (the geometry of every val / train pipeline in the reference's
// configs): the resized crop covers columns [left, left + dw) and rows [top, top + dh) of the output; lanes / rows
// outside it emit the pad value and never touch the source.
// AUG = true: the per-crop train-time augmentations (K1Params::aug_*) on the resized + padded uint8 pixel before
// Normalize, exactly as in the general kernel (k1_general_impl.cuh): flips mirror the store address, brightness /
// contrast and HueSaturationValue run per pixel, CoarseDropout holes become a per-row column mask.
// LBM = 0: A.Resize; 1: letterbox; 2: letterbox with PadIfNeeded(value = 0) -- every pipeline in the reference's configs:
// border columns already carry zero weights, so their pixel comes out 0 = the pad value and the per-element select of
// the pad value (two ALU instructions on the pipe this kernel loads most) is dropped.
template <int JMAX, typename OutT, int LBM, bool AUG = false>
__global__ void __launch_bounds__(K1_WARPS * 32, AUG ? K1F_MIN_BLOCKS_AUG : K1F_MIN_BLOCKS)
k1_crop_resize_normalize_tma(const K1Params p) {
    constexpr bool LB = LBM != 0;        // letterbox geometry
    constexpr bool PADSEL = LBM == 1;    // border columns need the pad value selected in (it is not zero)
    __shared__ __align__(128) uint8_t ring[K1_WARPS][K1F_RING_BYTES];
    __shared__ __align__(8) uint64_t bars[K1_WARPS][K1F_MAX_SLOTS];
    __shared__ int fetch_rows[K1_WARPS][64];
    __shared__ uint32_t htab[2][JMAX][32];   // the CTA's horizontal tables: window offset (bit 31: valid), coefficients
    // HueSaturationValue: OpenCV's two division tables (built once per CTA) + per warp the current crop's 3 x 256 tables
    __shared__ int hsv_div_tab[AUG ? 512 : 1];
    __shared__ __align__(16) uint8_t hsv_lut_s[AUG ? K1_WARPS * 768 : 16];
    __shared__ __align__(16) uint8_t bc_lut_s[AUG ? K1_WARPS * 256 : 16];   // per warp: the crop's brightness / contrast table
    if (AUG && p.aug_hsv_lut != nullptr) {
        for (int i = threadIdx.x; i < 256; i += K1_WARPS * 32) {
            hsv_div_tab[i] = k1_hsv_sdiv(i);
            hsv_div_tab[256 + i] = k1_hsv_hdiv(i);
        }
        __syncthreads();
    }

    // crop-major block order: the row blocks of a crop -- and the crops of a frame, which the caller lists together -- are
    // resident at the same time, so source rows shared by overlapping boxes (the lower half of one box is the upper half
    // of another) are fetched from HBM once and served from L2 afterwards.  (Row-block-major order walked the frames once
    // per row block: 520 MB of DRAM reads per launch for 305 MB of distinct source bytes.)
    const int crop = blockIdx.x / p.fby_fast;
    const int yblk = blockIdx.x - crop * p.fby_fast;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    struct K1Stamp {   // debug timeline: entry stamp by thread 0, exit stamp = the latest warp exit (atomicMax)
        unsigned long long* t;
        __device__ static unsigned long long now() {
            unsigned long long v;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
            return v;
        }
        __device__ K1Stamp(unsigned long long* base) : t(base) {
            if (t != nullptr && threadIdx.x == 0) {
                unsigned int sm;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
                t[0] = sm;
                t[1] = now();
            }
        }
        __device__ ~K1Stamp() {
            if (t != nullptr && (threadIdx.x & 31) == 0) atomicMax(t + 2, now());
        }
    } stamp(p.timing == nullptr ? nullptr : p.timing + 3 * ((size_t)blockIdx.z * gridDim.x + blockIdx.x));
    const int band = yblk * K1_WARPS + warp;
    const int y_begin = band * p.rows_per_warp_fast;
    const bool active = y_begin < p.out_h;   // (a warp past the last row still helps with the CTA's horizontal tables)
    const int nrows = active ? min(p.rows_per_warp_fast, p.out_h - y_begin) : 0;   // <= 32: one output row per lane
    const int ox0 = blockIdx.z * (32 * JMAX) + lane;

    const CropGeom g = load_geom(p, crop);
    uint32_t seg_start, seg_bytes, slot_stride;
    int nslot;
    int dw = p.out_w, dh = p.out_h, top = 0, left = 0;
    bool fast = fast_path_qualifies(p, g, seg_start, seg_bytes, slot_stride, nslot);
    if (LB && fast) fast = letterbox_geometry(g.bh, g.bw, p.max_size, p.out_h, p.out_w, dh, dw, top, left);
    if (!fast) {   // (uniform over the CTA: it depends on the crop only)
        // unaligned frame rows, a box wider than the ring, an invalid box or a letterbox that does not fit: same
        // arithmetic, direct loads
        if (active)
            k1_process_band<JMAX, OutT, true, false, AUG>(p, crop, g, y_begin, nrows, ox0,
                                                          yblk == 0 && blockIdx.z == 0 && warp == 0,
                                                          AUG ? hsv_div_tab : nullptr, AUG ? hsv_lut_s + warp * 768 : nullptr);
        return;
    }

    // ---- per-warp barriers ----
    const uint32_t bar0 = smem_u32(&bars[warp][0]);
    const uint32_t ring0 = smem_u32(&ring[warp][0]);
    if (active && lane == 0) {
        for (int s = 0; s < nslot; ++s) mbar_init(bar0 + 8u * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }

    // ---- vertical tables: lane l holds output row y_begin + l ----
    int my_r0 = -1, my_r1 = -1;
    uint32_t my_b0 = 0, my_b1 = 0;
    if (lane < nrows) {
        const int dy = y_begin + lane - top;
        if (!LB || (dy >= 0 && dy < dh)) {   // LB: border rows keep r0 = -1 and never enter the fetch list
            int s, c0, c1;
            axis_coef(dy, axis_scale(dh, g.bh), g.bh, false, s, c0, c1);
            my_r0 = min(max(s, 0), g.bh - 1);
            my_r1 = min(max(s + 1, 0), g.bh - 1);
            my_b0 = uint32_t(c0) << 16;  // pre-shifted: umulhi(b << 16, h) == (b * h) >> 16
            my_b1 = uint32_t(c1) << 16;
        }
    }

    // ---- fetch list: the strictly increasing sequence of source rows this band consumes ----
    int nfetch;
    {
        int prev_r1 = __shfl_up_sync(0xffffffffu, my_r1, 1);
        if (lane == 0) prev_r1 = -1;
        const bool new0 = lane < nrows && my_r0 > prev_r1;
        const bool new1 = lane < nrows && my_r1 > my_r0 && my_r1 > prev_r1;
        const uint32_t m0 = __ballot_sync(0xffffffffu, new0), m1 = __ballot_sync(0xffffffffu, new1);
        const uint32_t lt = (1u << lane) - 1u;
        const int pos0 = __popc(m0 & lt) + __popc(m1 & lt);
        if (new0) fetch_rows[warp][pos0] = my_r0;
        if (new1) fetch_rows[warp][pos0 + (new0 ? 1 : 0)] = my_r1;
        nfetch = __popc(m0) + __popc(m1);
    }
    __syncwarp();

    const uint8_t* const src_seg = p.frames + g.f_off + (int64_t)g.by0 * g.pitch + seg_start;
    const uint32_t pitch32 = (uint32_t)g.pitch;  // qualifying crops have bh * pitch < 2^31 (fast_path_qualifies)
    const uint64_t l2pol = l2_policy_evict_last();
    // lane 0: start the bulk copy of fetch number k into the slot at (slot_addr, bar_addr)
    auto issue = [&](int k, uint32_t slot_addr, uint32_t bar_addr) {
        const uint32_t row = (uint32_t)fetch_rows[warp][k];
        mbar_expect_tx(bar_addr, seg_bytes);
        bulk_g2s(slot_addr, src_seg + row * pitch32, seg_bytes, bar_addr, l2pol);
    };
    if (lane == 0) {
        const int pre = min(nslot, nfetch);
        for (int k = 0; k < pre; ++k) issue(k, ring0 + slot_stride * k, bar0 + 8u * k);
    }

    // (the first bulk copies are in flight: the horizontal tables below are computed under their latency -- the prologue's
    // dependent chain geometry loads -> coefficient arithmetic -> first source row is what a CTA slot idles on, 2.6 us per
    // band before this reordering)
    // ---- horizontal tables: smem byte offset of the window, funnel shift, packed coefficients ----
    // They depend on the crop and the column only, not on the band: the CTA's four warps share them -- warp w computes
    // the columns j = w, w + 4 and everybody reads all of them back (two of seven double-precision coefficient sets per
    // warp instead of seven: the prologue is what makes short bands expensive).
    uint32_t soa[JMAX], sk8[JMAX], cf[JMAX];
    uint32_t vmask = (1u << JMAX) - 1u;   // LB: columns of this lane that receive resized pixels (the rest is border)
    {
        const double sxs = axis_scale(dw, g.bw);
        for (int j = warp; j < JMAX; j += K1_WARPS) {
            int s = 0, c0 = 0, c1 = 0;
            const int dx = ox0 + 32 * j - left;
            const bool valid = !LB || (dx >= 0 && dx < dw);
            if (valid) axis_coef(dx, sxs, g.bw, true, s, c0, c1);
            int px = g.bx0 + s;
            uint32_t c = uint32_t(c0) | (uint32_t(c1) << 16);
            if (px + 1 >= g.fw) {  // window would leave the frame row: shift it left, weight moves to tap 1
                px -= 1;
                c = uint32_t(c0) << 16;
            }
            if (LB && !valid) c = 0u;   // border column: reads the box's first window, weights zero (value replaced below)
            const uint32_t so = uint32_t(px) * 3u - seg_start;
            htab[0][j][lane] = so | (uint32_t(valid) << 31);   // (so < the ring size)
            htab[1][j][lane] = c;
        }
        __syncthreads();
        if (!active) return;
        if (LB) vmask = 0;
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const uint32_t so = htab[0][j][lane];
            if (LB) vmask |= (so >> 31) << j;
            soa[j] = so & 0x7ffffffcu;
            sk8[j] = (so & 3u) * 8u;
            cf[j] = htab[1][j][lane];
        }
    }

    const uint32_t sel01 = (p.sel[0] & 0xFFu) | ((p.sel[1] & 0xFFu) << 8), sel2 = p.sel[2];
    const float m0f = p.m[0], m1f = p.m[1], m2f = p.m[2];
    const float d0f = p.d[0], d1f = p.d[1], d2f = p.d[2];
    const int64_t plane = (int64_t)p.out_h * p.out_w;
    OutT* o = reinterpret_cast<OutT*>(p.out) + (int64_t)crop * 3 * plane + (int64_t)y_begin * p.out_w + ox0;
    const int out_w = p.out_w;

    // ---- augmentation parameters of this crop (AUG only) ----
    int aflags = 0, nholes = 0;
    float a_alpha = 1.f, a_beta = 0.f;
    const int32_t* holes = nullptr;
    const uint8_t* hsv_lut = nullptr;
    if (AUG) {
        aflags = __ldg(p.aug_flags + crop);
        nholes = min(aflags >> 8, p.aug_max_holes);
        holes = p.aug_holes + (int64_t)crop * p.aug_max_holes * 4;
        if (aflags & K1_AUG_BC) {
            // albumentations applies brightness / contrast as a 256-entry table through cv2.LUT; so does this warp: eight
            // entries per lane, then one byte load per channel instead of six instructions
            a_alpha = __ldg(p.aug_alpha + crop); a_beta = __ldg(p.aug_beta + crop);
            uint8_t* t = bc_lut_s + warp * 256;
#pragma unroll
            for (int i = 0; i < 8; ++i) t[lane + 32 * i] = (uint8_t)k1_brightness_contrast((uint32_t)(lane + 32 * i), a_alpha, a_beta);
        }
        if (aflags & K1_AUG_HSV) {   // this crop's hue / sat / val tables -> the warp's 768 bytes of shared memory
            const uint32_t* src = reinterpret_cast<const uint32_t*>(p.aug_hsv_lut + (int64_t)crop * 768);
            uint32_t* dst = reinterpret_cast<uint32_t*>(hsv_lut_s + warp * 768);
#pragma unroll
            for (int i = 0; i < 6; ++i) dst[lane + 32 * i] = __ldg(src + lane + 32 * i);
            hsv_lut = hsv_lut_s + warp * 768;
        }
        __syncwarp();
    }
    const uint8_t* const bc_lut = bc_lut_s + (AUG ? warp * 256 : 0);
    const bool hflip = AUG && (aflags & K1_AUG_HFLIP), vflip = AUG && (aflags & K1_AUG_VFLIP);
    const int xstep = hflip ? -32 : 32;                       // destination column step between this lane's j's
    const int xd0 = hflip ? p.out_w - 1 - ox0 : ox0;          // destination column of j = 0
    const int trunc_cols = AUG ? p.aug_hsv_trunc_cols : 0;
    OutT* const out_crop = reinterpret_cast<OutT*>(p.out) + (int64_t)crop * 3 * plane;
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
    // the colour ops over one row of this lane's pixels, each under ONE warp-uniform test, in the reference's order:
    // RandomBrightnessContrast, HueSaturationValue, CoarseDropout fill.  P is in OUTPUT channel order.
    auto augment_row = [&](uint32_t (&P)[JMAX][3], uint32_t hm) {
        if (aflags & K1_AUG_BC) {
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) P[j][c] = bc_lut[P[j][c]];
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
                if (hm >> j & 1) { P[j][0] = p.aug_fill[0]; P[j][1] = p.aug_fill[1]; P[j][2] = p.aug_fill[2]; }
        }
    };
    // AUG: Normalize + store one row of augmented pixels at destination row pointer od (column xd0, step xstep)
    auto store_row = [&](OutT* od, const uint32_t (&P)[JMAX][3]) {
        const float mf[3] = {m0f, m1f, m2f}, df[3] = {d0f, d1f, d2f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            OutT* oc = od + c * plane;
#pragma unroll
            for (int j = 0; j < JMAX; j += 2) {
                if (j + 1 < JMAX) {
                    float ra, rb;
                    normalize2((float)P[j][c], (float)P[j + 1][c], mf[c], df[c], ra, rb);
                    store_out<OutT>(oc + xstep * j, ra);
                    store_out<OutT>(oc + xstep * (j + 1), rb);
                } else {
                    store_out<OutT>(oc + xstep * j, __fmul_rn(__fsub_rn((float)P[j][c], mf[c]), df[c]));
                }
            }
        }
    };

    uint32_t HA[JMAX][3], HB[JMAX][3];
    int iA = -1, iB = -1;
    int consumed = 0;
    uint32_t cslot = 0, cparity = 0, cur_slot = ring0, cur_bar = bar0;
    const uint32_t ring_end = ring0 + slot_stride * (uint32_t)nslot;

    // wait for the next staged row, run the horizontal pass into H, refill the slot
    auto consume_into = [&](uint32_t (&H)[JMAX][3]) {
        mbar_wait(cur_bar, cparity);
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const uint32_t a = cur_slot + soa[j];
            uint32_t w0, w1, w2;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a));
            asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(a));
            asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(a));
            const uint32_t lo = __funnelshift_r(w0, w1, sk8[j]);  // bytes o .. o+3
            const uint32_t hi = __funnelshift_r(w1, w2, sk8[j]);  // bytes o+4 .. o+7
            // one PRMT pairs the taps of TWO channels (bytes 0,1 and 2,3); dp2a.lo / dp2a.hi pick the pair
            const uint32_t t01 = __byte_perm(lo, hi, sel01);
            H[j][0] = __dp2a_lo(cf[j], t01, 0u) >> 4;
            H[j][1] = __dp2a_hi(cf[j], t01, 0u) >> 4;
            H[j][2] = __dp2a_lo(cf[j], __byte_perm(lo, hi, sel2), 0u) >> 4;
        }
        __syncwarp();  // every lane has read the slot before it is overwritten
        if (lane == 0 && consumed + nslot < nfetch) issue(consumed + nslot, cur_slot, cur_bar);
        ++consumed;
        cur_slot += slot_stride;
        cur_bar += 8u;
        if (cur_slot == ring_end) { cur_slot = ring0; cur_bar = bar0; cparity ^= 1u; }
    };
    (void)cslot;

    for (int yy = 0; yy < nrows; ++yy) {
        const int r0 = __shfl_sync(0xffffffffu, my_r0, yy);
        const int r1 = __shfl_sync(0xffffffffu, my_r1, yy);
        const uint32_t b0 = __shfl_sync(0xffffffffu, my_b0, yy);
        const uint32_t b1 = __shfl_sync(0xffffffffu, my_b1, yy);
        // AUG: destination row (vertical flip mirrors it), its pointer and its CoarseDropout mask
        const int yd = AUG ? (vflip ? p.out_h - 1 - (y_begin + yy) : y_begin + yy) : 0;
        OutT* const od = AUG ? out_crop + (int64_t)yd * out_w + xd0 : nullptr;
        const uint32_t hmask = AUG ? hole_mask(yd) : 0u;
        if (AUG && LB && r0 < 0) {   // letterbox border row under augmentation: the pad value goes through the colour ops
            uint32_t P[JMAX][3];
#pragma unroll
            for (int j = 0; j < JMAX; ++j) { P[j][0] = p.padu[0]; P[j][1] = p.padu[1]; P[j][2] = p.padu[2]; }
            augment_row(P, hmask);
            store_row(od, P);
            continue;
        }
        if (LB && r0 < 0) {   // letterbox border row: the normalised pad value, no source row involved
#pragma unroll
            for (int j = 0; j < JMAX; ++j) {
                store_out<OutT>(o + 32 * j, p.padf[0]);
                store_out<OutT>(o + plane + 32 * j, p.padf[1]);
                store_out<OutT>(o + 2 * plane + 32 * j, p.padf[2]);
            }
            o += out_w;
            continue;
        }
        uint32_t PA[AUG ? JMAX : 1][3];   // AUG: this row's resized (and padded) pixels
        auto vertical = [&](const uint32_t (&Ht)[JMAX][3], const uint32_t (&Hb)[JMAX][3]) {
            if constexpr (AUG) {   // only the pixels here: the colour ops and the stores follow ONCE below (this lambda is
                                   // inlined four times; four copies of the colour ops thrash the instruction cache)
#pragma unroll
                for (int j = 0; j < JMAX; ++j)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        PA[j][c] = vtap(b0, Ht[j][c], b1, Hb[j][c]);
                        if (PADSEL && !(vmask >> j & 1)) PA[j][c] = p.padu[c];
                    }
                return;
            }
            // Normalize two columns per instruction: (v - m) * d as FADD2 + FMUL2 (packed fp32, IEEE round-to-nearest
            // per half: the same two separately rounded operations, half the issue slots)
            const float mf[3] = {m0f, m1f, m2f}, df[3] = {d0f, d1f, d2f};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                OutT* oc = o + c * plane;
#pragma unroll
                for (int j = 0; j < JMAX; j += 2) {
                    uint32_t va = vtap(b0, Ht[j][c], b1, Hb[j][c]);
                    if (PADSEL && !(vmask >> j & 1)) va = p.padu[c];
                    if (j + 1 < JMAX) {
                        uint32_t vb = vtap(b0, Ht[j + 1][c], b1, Hb[j + 1][c]);
                        if (PADSEL && !(vmask >> (j + 1) & 1)) vb = p.padu[c];
                        float ra, rb;
                        normalize2((float)va, (float)vb, mf[c], df[c], ra, rb);
                        store_out<OutT>(oc + 32 * j, ra);
                        store_out<OutT>(oc + 32 * (j + 1), rb);
                    } else {
                        store_out<OutT>(oc + 32 * j, __fmul_rn(__fsub_rn((float)va, mf[c]), df[c]));
                    }
                }
            }
        };
        // Ht must hold source row r0, Hb row r1 (r1 == r0 only when the tap is clamped at an edge)
        auto row_step = [&](uint32_t (&Ht)[JMAX][3], uint32_t (&Hb)[JMAX][3], int& it, int& ib) {
            if (r0 != it) { consume_into(Ht); it = r0; }
            if (r1 != r0) {
                if (r1 != ib) { consume_into(Hb); ib = r1; }
                vertical(Ht, Hb);
            } else {
                vertical(Ht, Ht);
            }
        };
        // roles flip (no register copies) when the lower tap of the previous row is this row's upper tap
        const bool a_is_top = (r0 == iA) || (r0 != iB);
        if (a_is_top) row_step(HA, HB, iA, iB);
        else row_step(HB, HA, iB, iA);
        if constexpr (AUG) {
            augment_row(PA, hmask);
            store_row(od, PA);
        }
        o += out_w;
    }
}


}  // namespace nkbk
