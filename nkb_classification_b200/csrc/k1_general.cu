// K1: fused YOLO-box crop + cv2-INTER_LINEAR-exact uint8 resize + Normalize +
// HWC->NCHW, one launch per batch.  See include/nkbk.h for the contract and
// DESIGN.md section "K1" for the layout / roofline reasoning.
//
// Work decomposition (v1, "row walker"):
//   grid  = (crop, band-of-rows block, column tile)       block = 4 warps
//   warp  = one band of `rows_per_warp` consecutive output rows of one crop
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
//   (b*(H>>4))>>16 with the "+2" and the second tap folded into the addend,
//   then >> 2, int->float, (v - mean255) * denom as two rounded fp32 ops.
#include "k1_common.cuh"

namespace nkbk {

bool launch_k1_fast(const K1Params& p, int jmax, dim3 grid, cudaStream_t st, bool f32);  // k1_fast.cu

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
        H[j][0] = __dp2a_lo(cf[j], __byte_perm(lo, hi, sel0), 0u) >> 4;
        H[j][1] = __dp2a_lo(cf[j], __byte_perm(lo, hi, sel1), 0u) >> 4;
        H[j][2] = __dp2a_lo(cf[j], __byte_perm(lo, hi, sel2), 0u) >> 4;
    }
}

// GENERAL = false: stretch mode and out_w a multiple of the 32*JMAX column tile -> every lane
// writes every column, no border handling, no uint8 side output.  GENERAL = true: everything.
template <int JMAX, typename OutT, bool GENERAL, bool WRITE_U8>
__device__ __forceinline__ void k1_process_crop(const K1Params& p, const int crop) {
    static_assert(GENERAL || !WRITE_U8, "uint8 side output only in the general variant");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int band = blockIdx.y * K1_WARPS + warp;
    const int y_begin = band * p.rows_per_warp;
    if (y_begin >= p.out_h) return;
    const int nrows = min(p.rows_per_warp, p.out_h - y_begin);
    const int ox0 = blockIdx.z * (32 * JMAX) + lane;

    // ---- crop geometry (warp-uniform) ----
    const CropGeom g = load_geom(p, crop);
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

    if (!ok) {
        // empty / out-of-frame box: emit the normalised pad value, count it once per crop
        if (p.bad_count != nullptr && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0)
            atomicAdd(p.bad_count, 1);
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
        const int y = y_begin + yy;
        OutT* o = out_crop + (int64_t)y * p.out_w + ox0;
        uint8_t* u = WRITE_U8 ? p.out_u8 + (((int64_t)crop * p.out_h + y) * p.out_w + ox0) * 3 : nullptr;

        if (GENERAL && r0 < 0) {  // letterbox border row
#pragma unroll
            for (int j = 0; j < JMAX; ++j)
                if (wmask >> j & 1) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) store_out<OutT>(o + c * plane + 32 * j, p.padf[c]);
                    if (WRITE_U8) {
                        u[96 * j + 0] = (uint8_t)p.padu[0]; u[96 * j + 1] = (uint8_t)p.padu[1];
                        u[96 * j + 2] = (uint8_t)p.padu[2];
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

#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            uint32_t px[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t t0 = __umulhi(b0, Ha[j][c]);
                const uint32_t t1 = __umulhi(b1, Hb[j][c]);
                px[c] = (t0 + t1 + 2u) >> 2;
                if (GENERAL && !(vmask >> j & 1)) px[c] = p.padu[c];
            }
            const float f0 = __fmul_rn(__fsub_rn((float)px[0], m0), d0);
            const float f1 = __fmul_rn(__fsub_rn((float)px[1], m1), d1);
            const float f2 = __fmul_rn(__fsub_rn((float)px[2], m2), d2);
            if (!GENERAL || (wmask >> j & 1)) {
                store_out<OutT>(o + 32 * j, f0);
                store_out<OutT>(o + plane + 32 * j, f1);
                store_out<OutT>(o + 2 * plane + 32 * j, f2);
                if (WRITE_U8) {
                    u[96 * j + 0] = (uint8_t)px[0]; u[96 * j + 1] = (uint8_t)px[1]; u[96 * j + 2] = (uint8_t)px[2];
                }
            }
        }
    }
}

// Normal mode: one crop per blockIdx.x.  Fix-up mode (skip_fast, small grid behind the TMA kernel): the
// block's threads test 128 crops at a time against the fast-path predicate in parallel and the block then
// produces only the crops the TMA kernel left out -- a few microseconds when there are none.
template <int JMAX, typename OutT, bool GENERAL, bool WRITE_U8>
__global__ void __launch_bounds__(K1_WARPS * 32) k1_crop_resize_normalize(const K1Params p) {
    if (!p.skip_fast) {
        for (int crop = blockIdx.x; crop < p.n; crop += gridDim.x)
            k1_process_crop<JMAX, OutT, GENERAL, WRITE_U8>(p, crop);
        return;
    }
    __shared__ int todo[K1_WARPS * 32];
    __shared__ int ntodo;
    for (int base = blockIdx.x * (K1_WARPS * 32); base < p.n; base += gridDim.x * (K1_WARPS * 32)) {
        if (threadIdx.x == 0) ntodo = 0;
        __syncthreads();
        const int crop = base + threadIdx.x;
        if (crop < p.n) {
            const CropGeom g = load_geom(p, crop);
            uint32_t a, b, c;
            int d;
            if (!fast_path_qualifies(p, g, a, b, c, d)) todo[atomicAdd(&ntodo, 1)] = crop;
        }
        __syncthreads();
        const int nt = ntodo;
        for (int i = 0; i < nt; ++i) k1_process_crop<JMAX, OutT, GENERAL, WRITE_U8>(p, todo[i]);
        __syncthreads();
    }
}

template <int JMAX, typename OutT>
static void launch_k1(const K1Params& p, dim3 grid, cudaStream_t st, bool general) {
    (void)general;
    if (p.out_u8 != nullptr)
        k1_crop_resize_normalize<JMAX, OutT, true, true><<<grid, K1_WARPS * 32, 0, st>>>(p);
    else
        k1_crop_resize_normalize<JMAX, OutT, true, false><<<grid, K1_WARPS * 32, 0, st>>>(p);
}

}  // namespace nkbk

using namespace nkbk;

extern "C" int nkbk_preprocess_crops(const void* frames_base, const int64_t* frame_desc, int n_frames,
                                     const int32_t* boxes, const int32_t* frame_idx, int n, int mode, int out_h,
                                     int out_w, int max_size, const uint8_t* pad_value, const float* mean255,
                                     const float* denom, int channel_swap, void* out, int out_dtype, uint8_t* out_u8,
                                     int32_t* bad_count, void* stream) {
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

    // rows per warp: split the height into blocks of <= 64 rows, 4 warps each
    const int nby = (out_h + 63) / 64;
    p.rows_per_warp = (out_h + nby * K1_WARPS - 1) / (nby * K1_WARPS);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool f32 = out_dtype == NKBK_F32;
    p.skip_fast = 0;
    p.rows_per_warp_fast = p.rows_per_warp;

    // ---- fast path: A.Resize, output width a whole number of 32*J column tiles, no uint8 side output ----
    // Crops whose frame rows are not 16-byte aligned or whose boxes are too wide for the shared-memory ring are
    // left untouched by the TMA kernel and produced by the general kernel in a small-grid fix-up pass.
    if (mode == NKBK_MODE_STRETCH && out_u8 == nullptr && out_w % 32 == 0) {
        const int cols = out_w / 32;
        int fj = 0;
        for (int j = 8; j >= 4; --j)
            if (cols % j == 0) { fj = j; break; }
        // the TMA kernel amortises its per-band prologue over up to 32 rows per warp (128 per CTA)
        const int fby = (out_h + 32 * K1_WARPS - 1) / (32 * K1_WARPS);
        p.rows_per_warp_fast = (out_h + fby * K1_WARPS - 1) / (fby * K1_WARPS);
        if (fj != 0 && cols / fj <= 65535 && fby <= 65535) {
            dim3 fgrid((unsigned)n, (unsigned)fby, (unsigned)(cols / fj));
            if (launch_k1_fast(p, fj, fgrid, st, f32)) {
                NKBK_CHECK_LAUNCH("k1_crop_resize_normalize_tma");
                p.skip_fast = 1;
            }
        }
    }

    // ---- general path (everything), or the fix-up pass behind the fast kernel ----
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
    const int gx = p.skip_fast ? ((n + 127) / 128 < 148 ? (n + 127) / 128 : 148) : n;
    dim3 grid((unsigned)gx, (unsigned)nby, (unsigned)ntx);
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
