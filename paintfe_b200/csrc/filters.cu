// Box blur, motion blur, median and vignette.
// Reference: src/ops/effects/blur.rs:144-318, src/ops/effects/noise.rs:357-410,
// src/ops/effects/stylize.rs:170-191 (+ apply_per_pixel, src/ops/effects.rs:53-100).
// All four are HBM-light (8 algorithmic bytes per pixel); box and median are integer and
// bit-exact by construction, motion blur sums integers in f32 (exact), vignette is strict f32.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace {

__device__ __forceinline__ uint4 unpack4(uint32_t v) {
    return make_uint4(v & 255u, (v >> 8) & 255u, (v >> 16) & 255u, v >> 24);
}
__device__ __forceinline__ void add4(uint4 &s, uint32_t v) {
    s.x += v & 255u; s.y += (v >> 8) & 255u; s.z += (v >> 16) & 255u; s.w += v >> 24;
}
__device__ __forceinline__ void sub4(uint4 &s, uint32_t v) {
    s.x -= v & 255u; s.y -= (v >> 8) & 255u; s.z -= (v >> 16) & 255u; s.w -= v >> 24;
}
// (sum + d/2) / d, blur.rs:269.  The divisor is the same for every pixel, so the quotient is one IMAD.HI against
// magic = floor(2^32 / d) + 1: exact for every numerator below 256 d as long as 256 d^2 <= 2^32, i.e. d <= 4096
// (checked exhaustively on the host for the divisors the tests use); magic == 0 selects the plain division.
__device__ __forceinline__ uint32_t box_out(const uint4 &s, uint32_t d, uint32_t magic) {
    const uint32_t h = d / 2;
    if (magic) return pfe_pack(__umulhi(s.x + h, magic), __umulhi(s.y + h, magic), __umulhi(s.z + h, magic), __umulhi(s.w + h, magic));
    return pfe_pack((s.x + h) / d, (s.y + h) / d, (s.z + h) / d, (s.w + h) / d);
}

// ---- box blur H pass: block = 32 rows x 128 output columns; lane = row, warp = 32-column run.
// The clamped input tile is staged with coalesced loads; each thread slides a running window
// sum along its run (sums are windows over clamped indices, identical to the reference's
// incremental add/remove at blur.rs:272-278).
constexpr int BOX_COLS = 128;
__global__ void __launch_bounds__(128) box_h_kernel(const uint32_t *src, uint32_t *dst, int w, int h, int r, uint32_t magic) {
    extern __shared__ uint32_t sm[];
    const int tw = BOX_COLS + 2 * r;       // tile width
    const int pitch = tw | 1;              // odd pitch: lanes (rows) hit distinct banks
    uint32_t *tin = sm;
    uint32_t *tout = sm + 32 * pitch;      // 32 x (BOX_COLS+1)
    const int x0 = blockIdx.x * BOX_COLS, y0 = blockIdx.y * 32;
    for (int idx = threadIdx.x; idx < 32 * tw; idx += blockDim.x) {
        int ry = idx / tw, cx = idx - ry * tw;
        int y = min(y0 + ry, h - 1), x = pfe_clampi(x0 - r + cx, 0, w - 1);
        tin[ry * pitch + cx] = __ldg(src + (size_t)y * w + x);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t *row = tin + lane * pitch + warp * 32;  // window for output c starts at row[c]
    const uint32_t d = 2u * (uint32_t)r + 1u;
    uint4 s = make_uint4(0, 0, 0, 0);
    for (int k = 0; k < (int)d; k++) add4(s, row[k]);
    for (int c = 0; c < 32; c++) {
        tout[lane * (BOX_COLS + 1) + warp * 32 + c] = box_out(s, d, magic);
        sub4(s, row[c]);
        add4(s, row[c + (int)d]);  // within the padded tile (+1 column slack below)
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * BOX_COLS; idx += blockDim.x) {
        int ry = idx / BOX_COLS, cx = idx - ry * BOX_COLS;
        int y = y0 + ry, x = x0 + cx;
        if (y < h && x < w) dst[(size_t)y * w + x] = tout[ry * (BOX_COLS + 1) + cx];
    }
}

// ---- box blur V pass: lanes along x (coalesced), each thread slides down a band of rows.
constexpr int BOX_BAND = 128;
__global__ void __launch_bounds__(128) box_v_kernel(const uint32_t *hb, const uint32_t *src, const uint8_t *mask,
                                                    uint32_t *dst, int w, int h, int r, uint32_t magic) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const int y0 = blockIdx.y * BOX_BAND, y1 = min(y0 + BOX_BAND, h);
    const uint32_t d = 2u * (uint32_t)r + 1u;
    uint4 s = make_uint4(0, 0, 0, 0);
    for (int k = -r; k <= r; k++) add4(s, __ldg(hb + (size_t)pfe_clampi(y0 + k, 0, h - 1) * w + x));
    for (int y = y0; y < y1; y++) {
        size_t o = (size_t)y * w + x;
        dst[o] = (mask && mask[o] == 0) ? src[o] : box_out(s, d, magic);            // blur.rs:298-306
        sub4(s, __ldg(hb + (size_t)pfe_clampi(y - r, 0, h - 1) * w + x));
        add4(s, __ldg(hb + (size_t)pfe_clampi(y + r + 1, 0, h - 1) * w + x));
    }
}

// ---- motion blur, blur.rs:144-210 ---------------------------------------------------------
// `v.round() as i32` (half away from zero) for |v| < 2^22 without FRND / F2I, which run on the quarter-rate XU pipe
// (the first version of this kernel spent 87 % of its time there): a round-toward-zero add of 0.5 to |v| cannot step
// over an integer, and a second one against 2^23 leaves floor() of that in the low mantissa bits.
__device__ __forceinline__ int round_half_away_i32(float v) {
    const int n = (int)(__float_as_uint(__fadd_rz(__fadd_rz(fabsf(v), 0.5f), 8388608.0f)) & 0x007FFFFFu);
    return v < 0.0f ? -n : n;
}
// SMALL: 2*steps+1 <= 257 samples, so the channel sums fit 16-bit lanes and are accumulated as packed integers (the
// reference adds the u8 values as f32, which is exact - integer sums are the same numbers, with no I2F per sample).
template <bool SMALL>
__global__ void __launch_bounds__(256) motion_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst,
                                                     int w, int h, int steps, float dx, float dy, float inv_steps) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    if (mask && mask[o] == 0) { dst[o] = src[o]; return; }
    const float fx = (float)x, fy = (float)y;
    float sr, sg, sb, sa;
    if (SMALL) {
        uint32_t rb = 0, ga = 0;  // (r | b << 16), (g | a << 16)
        for (int i = -steps; i <= steps; i++) {
            // (x as f32 + i as f32 * dx).round() as i32, then clamp (:196-199)
            const int sx = pfe_clampi(round_half_away_i32(fx + (float)i * dx), 0, w - 1);
            const int sy = pfe_clampi(round_half_away_i32(fy + (float)i * dy), 0, h - 1);
            const uint32_t v = __ldg(src + (size_t)sy * w + sx);
            rb += v & 0x00FF00FFu;
            ga += (v >> 8) & 0x00FF00FFu;
        }
        sr = (float)(rb & 0xFFFFu); sb = (float)(rb >> 16); sg = (float)(ga & 0xFFFFu); sa = (float)(ga >> 16);
    } else {
        sr = sg = sb = sa = 0.f;
        for (int i = -steps; i <= steps; i++) {
            float px = roundf(fx + (float)i * dx), py = roundf(fy + (float)i * dy);
            int sx = pfe_clampi(__float2int_rz(px), 0, w - 1), sy = pfe_clampi(__float2int_rz(py), 0, h - 1);
            uint32_t v = __ldg(src + (size_t)sy * w + sx);
            sr += (float)(v & 255u); sg += (float)((v >> 8) & 255u); sb += (float)((v >> 16) & 255u); sa += (float)(v >> 24);
        }
    }
    dst[o] = pfe_pack(pfe_round_u8(sr * inv_steps), pfe_round_u8(sg * inv_steps), pfe_round_u8(sb * inv_steps),
                      pfe_round_u8(sa * inv_steps));
}

// ---- median, noise.rs:357-410 --------------------------------------------------------------
// sorted[len/2] per channel over the clamp-to-edge (2r+1)^2 window.  Four kernels, all exact:
//   r <= 2    forgetful selection in registers (median_small_kernel);
//   r <= 32   column histograms (median_hist_kernel): cost per pixel grows with r, not with r^2;
//   r <= 127  bitwise bisection over the window staged in shared memory (median_kernel);
//   larger    the same bisection straight from global memory with 32-bit counts (median_global_kernel).

// One compare-exchange of two pixels held as (R | B << 16, G | A << 16): afterwards a <= b in every 16-bit lane.
// VIMNMX.U16x2 is native on sm_100; the 8-bit SIMD min / max / compares are emulated with ~6 instructions each.
struct Px2 { uint32_t lo, hi; };
__device__ __forceinline__ void cex(Px2 &a, Px2 &b) {
    const uint32_t mnl = __vminu2(a.lo, b.lo), mxl = __vmaxu2(a.lo, b.lo), mnh = __vminu2(a.hi, b.hi), mxh = __vmaxu2(a.hi, b.hi);
    a.lo = mnl; b.lo = mxl; a.hi = mnh; b.hi = mxh;
}
__device__ __forceinline__ Px2 widen(uint32_t v) { return Px2{v & 0x00FF00FFu, (v >> 8) & 0x00FF00FFu}; }

// Forgetful selection: of any k >= M + 2 of the N = 2M + 1 window values, the smallest has at least M + 1 values above
// it and the largest M + 1 below it, so neither is the median.  A working set of M + 2 values loses both each round and
// gains the next window value; after M - 1 rounds three values are left and the middle one is sorted[N / 2].  The
// compare-exchange network is data-oblivious, so all four channels run through it at once in 16-bit lanes.
template <int R>
__global__ void __launch_bounds__(256) median_small_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h) {
    constexpr int SIDE = 2 * R + 1, N = SIDE * SIDE, M = N / 2, K = M + 2;
    constexpr int tw = 32 + 2 * R, th = 8 + 2 * R;
    __shared__ uint32_t sm[th * tw];
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    for (int idx = threadIdx.x; idx < tw * th; idx += blockDim.x) {
        int ty = idx / tw, tx = idx - ty * tw;
        sm[idx] = __ldg(src + (size_t)pfe_clampi(y0 - R + ty, 0, h - 1) * w + pfe_clampi(x0 - R + tx, 0, w - 1));
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    if (mask && mask[o] == 0) { dst[o] = src[o]; return; }
    const uint32_t *win = sm + ly * tw + lx;
    Px2 a[K];
#pragma unroll
    for (int i = 0; i < K; i++) a[i] = widen(win[(i / SIDE) * tw + (i % SIDE)]);
    // round with k live values a[0..k-1]: one pass bubbles the maximum up to a[k-1], one pass back bubbles the minimum
    // down to a[0]; a[k-1] is dropped, a[0] is overwritten by the next window element: k-1 live values remain
#pragma unroll
    for (int k = K; k > 3; k--) {
#pragma unroll
        for (int i = 0; i + 1 < k; i++) cex(a[i], a[i + 1]);
#pragma unroll
        for (int i = k - 3; i >= 0; i--) cex(a[i], a[i + 1]);
        constexpr int first_new = K;
        const int e = first_new + (K - k);  // window elements K .. N-1, one per round
        a[0] = widen(win[(e / SIDE) * tw + (e % SIDE)]);
    }
    cex(a[0], a[1]);
    cex(a[1], a[2]);
    cex(a[0], a[1]);
    dst[o] = a[1].lo | (a[1].hi << 8);
}

// Column-histogram median (the idea of Perreault & Hebert's constant-time median, arranged for one CTA per image
// strip).  A CTA owns a strip of 32 output columns and walks down a segment of rows; warp c handles channel c and
// keeps, for each of the strip's 32 + 2r source columns, a histogram of that column's 2r+1 window rows: 16 coarse bins
// (value >> 4) and 256 fine bins, 16-bit counts, two per word.  Moving down one row removes one pixel from and adds one
// pixel to every column histogram.  A lane then sums the coarse histograms of its 2r+1 columns, finds the coarse bin
// that holds rank len/2, sums that bin's 16 fine counters over the same columns and finds the value.  Column stride
// 140 words: neighbouring lanes start 12 banks apart, so the quarter-warp LDS.128 of the coarse pass are conflict free.
constexpr int kMedHistTW = 32, kMedHistStride = 140, kMedHistMaxR = 32;

__global__ void __launch_bounds__(128) median_hist_kernel(const uint32_t *src, const uint8_t *mask, uint8_t *dst, int w, int h,
                                                         int r, int seg_rows) {
    extern __shared__ __align__(16) uint32_t hist_sm[];
    const int nc = kMedHistTW + 2 * r;
    const int lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
    uint32_t *hist = hist_sm + (size_t)ch * nc * kMedHistStride;  // this warp's (= this channel's) column histograms
    const int x0 = blockIdx.x * kMedHistTW;
    const int ys = blockIdx.y * seg_rows, ye = min(ys + seg_rows, h);
    const uint32_t shift = 8u * (uint32_t)ch;
    const uint32_t target = (uint32_t)((2 * r + 1) * (2 * r + 1) / 2);

    for (int i = lane; i < nc * kMedHistStride; i += 32) hist[i] = 0u;
    __syncwarp();
    // +-1 on the 16-bit lane of fine bin v and of coarse bin v >> 4.  Adding 0xFFFF to the low lane subtracts one from
    // it and carries into the high lane, adding 0xFFFF0000 on top takes that carry back out: -1 on the low lane is
    // += 0xFFFFFFFF... written as one add of (0xFFFF + 0xFFFF0000) = 0xFFFFFFFF, i.e. a plain 32-bit decrement, which is
    // right as long as the low lane is >= 1 (it is: the pixel being removed was added before).
    auto bump = [&](uint32_t *col, uint32_t v, bool add) {
        const uint32_t lo = add ? 1u : 0xFFFFFFFFu, hi = add ? 0x00010000u : 0xFFFF0000u;
        col[8 + (v >> 1)] += (v & 1u) ? hi : lo;
        col[v >> 5] += ((v >> 4) & 1u) ? hi : lo;
    };
    auto row_pixels = [&](int yy, bool add) {
        const uint32_t *row = src + (size_t)pfe_clampi(yy, 0, h - 1) * w;
        for (int c = lane; c < nc; c += 32) {
            const uint32_t v = (__ldg(row + pfe_clampi(x0 - r + c, 0, w - 1)) >> shift) & 255u;
            bump(hist + (size_t)c * kMedHistStride, v, add);
        }
    };
    for (int yy = ys - r; yy <= ys + r; yy++) row_pixels(yy, true);
    __syncwarp();

    const uint32_t *base = hist + (size_t)lane * kMedHistStride;
    for (int y = ys; y < ye; y++) {
        const int x = x0 + lane;
        // ---- coarse level: 16 bins summed over the lane's 2r+1 columns (each 16-bit sum <= (2r+1)^2 <= 4225)
        uint32_t acc[8];
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0u;
        for (int j = 0; j <= 2 * r; j++) {
            const uint4 a = *reinterpret_cast<const uint4 *>(base + (size_t)j * kMedHistStride);
            const uint4 b = *reinterpret_cast<const uint4 *>(base + (size_t)j * kMedHistStride + 4);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        }
        uint32_t cum = 0, bin = 0, below = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const uint32_t cnt = (acc[b >> 1] >> (16 * (b & 1))) & 0xFFFFu;
            cum += cnt;
            const bool le = cum <= target;  // bins wholly at or below the rank
            bin += le ? 1u : 0u;
            below += le ? cnt : 0u;
        }
        // ---- fine level: the 16 counters of coarse bin `bin`, same columns
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0u;
        const uint32_t *fbase = base + 8 + bin * 8;
        for (int j = 0; j <= 2 * r; j++) {
            const uint4 a = *reinterpret_cast<const uint4 *>(fbase + (size_t)j * kMedHistStride);
            const uint4 b = *reinterpret_cast<const uint4 *>(fbase + (size_t)j * kMedHistStride + 4);
            acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
            acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
        }
        const uint32_t target2 = target - below;
        uint32_t fcum = 0, fbin = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            fcum += (acc[b >> 1] >> (16 * (b & 1))) & 0xFFFFu;
            fbin += fcum <= target2 ? 1u : 0u;
        }
        if (x < w) {
            const size_t o = (size_t)y * w + x;
            uint32_t v = bin * 16 + fbin;
            if (mask && mask[o] == 0) v = (__ldg(src + o) >> shift) & 255u;
            dst[o * 4 + ch] = (uint8_t)v;
        }
        __syncwarp();  // every lane has read this row's histograms
        if (y + 1 < ye) {
            row_pixels(y - r, false);
            row_pixels(y + r + 1, true);
            __syncwarp();
        }
    }
}

// Bitwise bisection: sorted[len/2] per channel == the largest v with #(x < v) <= len/2, found in 8 steps on all four
// channels at once with packed byte compares; per-channel counts live in 16-bit lanes (window <= 255x255).
__device__ __forceinline__ uint32_t median_select(const uint32_t *win, int pitch, int side, uint32_t target) {
    uint32_t cur = 0;
    const uint32_t tgt_lo = target | (target << 16);
#pragma unroll 1
    for (int bit = 7; bit >= 0; bit--) {
        const uint32_t trial = cur | (0x01010101u << bit);
        uint32_t c02 = 0, c13 = 0;  // counts for channels (0,2) and (1,3) in 16-bit lanes
        for (int yy = 0; yy < side; yy++) {
            const uint32_t *rowp = win + yy * pitch;
            for (int xx = 0; xx < side; xx++) {
                uint32_t lt = __vsetltu4(rowp[xx], trial);  // 1 per byte where x < trial
                c02 += lt & 0x00FF00FFu;
                c13 += (lt >> 8) & 0x00FF00FFu;
            }
        }
        // keep the trial bit in channel c iff count_c <= target
        uint32_t k02 = __vsetleu2(c02, tgt_lo), k13 = __vsetleu2(c13, tgt_lo);  // 1 per halfword
        uint32_t keep = (k02 & 0x00010001u) | ((k13 & 0x00010001u) << 8);       // 1 per byte
        cur |= (keep << bit) & (0x01010101u << bit);
    }
    return cur;
}

constexpr int MED_BX = 32, MED_BY = 8;
__global__ void __launch_bounds__(MED_BX *MED_BY) median_kernel(const uint32_t *src, const uint8_t *mask,
                                                                uint32_t *dst, int w, int h, int r) {
    extern __shared__ uint32_t sm[];
    const int tw = MED_BX + 2 * r, th = MED_BY + 2 * r;
    const int x0 = blockIdx.x * MED_BX, y0 = blockIdx.y * MED_BY;
    for (int idx = threadIdx.x; idx < tw * th; idx += blockDim.x) {
        int ty = idx / tw, tx = idx - ty * tw;
        sm[idx] = __ldg(src + (size_t)pfe_clampi(y0 - r + ty, 0, h - 1) * w + pfe_clampi(x0 - r + tx, 0, w - 1));
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    if (mask && mask[o] == 0) { dst[o] = src[o]; return; }
    const int side = 2 * r + 1;
    dst[o] = median_select(sm + ly * tw + lx, tw, side, (uint32_t)(side * side / 2));
}

// Large-radius fallback: window read straight from global memory (L1/L2 cached).
__global__ void __launch_bounds__(256) median_global_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst,
                                                            int w, int h, int r) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    if (mask && mask[o] == 0) { dst[o] = src[o]; return; }
    const int side = 2 * r + 1;
    const uint32_t target = (uint32_t)side * side / 2;
    uint32_t cur = 0;
    for (int bit = 7; bit >= 0; bit--) {
        const uint32_t trial = cur | (0x01010101u << bit);
        uint32_t c[4] = {0, 0, 0, 0};
        for (int dy = -r; dy <= r; dy++) {
            const uint32_t *rowp = src + (size_t)pfe_clampi(y + dy, 0, h - 1) * w;
            for (int dx = -r; dx <= r; dx++) {
                uint32_t v = __ldg(rowp + pfe_clampi(x + dx, 0, w - 1));
                c[0] += (v & 255u) < (trial & 255u);
                c[1] += ((v >> 8) & 255u) < ((trial >> 8) & 255u);
                c[2] += ((v >> 16) & 255u) < ((trial >> 16) & 255u);
                c[3] += (v >> 24) < (trial >> 24);
            }
        }
        for (int ch = 0; ch < 4; ch++)
            if (c[ch] <= target) cur |= (1u << bit) << (8 * ch);
    }
    dst[o] = cur;
}

// ---- vignette, stylize.rs:170-191 ------------------------------------------------------------
// VEC consecutive pixels per thread (16-byte loads and stores when VEC = 4); the two divisions by image-wide constants
// go through pfe_fast_div (dist <= a few max_dist, soft >= 0.01: nowhere near the exponent extremes).
__device__ __forceinline__ uint32_t vignette_px(uint32_t v, int x, int y, float amount, float soft, float cx, float cy, float max_dist) {
    float dx = (float)x - cx, dy = (float)y - cy;
    float dist = pfe_fast_div(sqrtf(dx * dx + dy * dy), max_dist);
    float q = fminf(pfe_fast_div(dist, soft), 1.0f);
    float vf = pfe_clampf(1.0f - (amount * (q * q)), 0.0f, 1.0f);  // powf(2.0) == x*x exactly
    return pfe_pack(pfe_round_u8(pfe_u8_to_f32(v & 255u) * vf), pfe_round_u8(pfe_u8_to_f32((v >> 8) & 255u) * vf),
                    pfe_round_u8(pfe_u8_to_f32((v >> 16) & 255u) * vf), v >> 24);
}
template <int VEC>
__global__ void __launch_bounds__(256) vignette_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w,
                                                       int h, float amount, float soft, float cx, float cy,
                                                       float max_dist) {
    const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * VEC, y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    uint32_t v[VEC];
    if constexpr (VEC == 4) {
        const uint4 q = *reinterpret_cast<const uint4 *>(src + o);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
        v[0] = src[o];
    }
#pragma unroll
    for (int k = 0; k < VEC; k++)
        if (!(mask && mask[o + k] == 0)) v[k] = vignette_px(v[k], x + k, y, amount, soft, cx, cy, max_dist);
    if constexpr (VEC == 4) *reinterpret_cast<uint4 *>(dst + o) = make_uint4(v[0], v[1], v[2], v[3]);
    else dst[o] = v[0];
}

int check(pfe_ctx *ctx, const void *src, const void *dst, uint32_t w, uint32_t h, const char *what) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !dst || !w || !h) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, what);
    if (w > 0x7FFFFFFFu / 4 || h > 0x7FFFFFFFu / 4) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, what);
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    return PFE_OK;
}

int copy_through(pfe_ctx *ctx, const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h) {
    if (src != dst) PFE_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)w * h * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return PFE_OK;
}

// The same algorithm with 8-bit column counters (a column holds at most 2r+1 <= 65 pixels of a value): a column is
// 16 coarse + 256 fine bytes = 68 words, half the shared memory of the 16-bit layout, so twice the CTAs fit an SM - the
// kernel is bound by shared-memory latency, not by issue.  Window sums need 16 bits: a lane adds up to K = 255 / (2r+1)
// columns byte-wise in a 32-bit word (no carry can cross a byte), then widens the four bytes into two 16-bit pairs.
constexpr int kMedHist8Stride = 68;  // words; 17 x 16 bytes: eight consecutive lanes hit eight different bank groups
__global__ void __launch_bounds__(128) median_hist8_kernel(const uint32_t *src, const uint8_t *mask, uint8_t *dst, int w, int h,
                                                          int r, int seg_rows) {
    extern __shared__ __align__(16) uint32_t hist_sm[];
    const int nc = kMedHistTW + 2 * r;
    const int lane = threadIdx.x & 31, ch = threadIdx.x >> 5;
    uint32_t *hist = hist_sm + (size_t)ch * nc * kMedHist8Stride;  // this warp's (= this channel's) column histograms
    const int x0 = blockIdx.x * kMedHistTW;
    const int ys = blockIdx.y * seg_rows, ye = min(ys + seg_rows, h);
    const uint32_t shift = 8u * (uint32_t)ch;
    const int side = 2 * r + 1, K = 255 / side;
    const uint32_t target = (uint32_t)(side * side / 2);

    for (int i = lane; i < nc * kMedHist8Stride; i += 32) hist[i] = 0u;
    __syncwarp();
    // +-1 on the byte of fine bin v and of coarse bin v >> 4: a byte being decremented is >= 1, so no borrow leaves it
    auto bump = [&](uint32_t *col, uint32_t v, bool add) {
        const uint32_t f = 1u << (8u * (v & 3u)), c = 1u << (8u * ((v >> 4) & 3u));
        col[4 + (v >> 2)] += add ? f : 0u - f;
        col[v >> 6] += add ? c : 0u - c;
    };
    auto row_pixels = [&](int yy, bool add) {
        const uint32_t *row = src + (size_t)pfe_clampi(yy, 0, h - 1) * w;
        for (int c = lane; c < nc; c += 32) {
            const uint32_t v = (__ldg(row + pfe_clampi(x0 - r + c, 0, w - 1)) >> shift) & 255u;
            bump(hist + (size_t)c * kMedHist8Stride, v, add);
        }
    };
    for (int yy = ys - r; yy <= ys + r; yy++) row_pixels(yy, true);
    __syncwarp();

    const uint32_t *base = hist + (size_t)lane * kMedHist8Stride;
    // 16 byte counters at `p` of each of the lane's 2r+1 columns -> s[2k] = bins 4k (low half) and 4k+2, s[2k+1] = bins 4k+1 and 4k+3
    auto window16 = [&](const uint32_t *p, uint32_t (&s)[8]) {
#pragma unroll
        for (int k = 0; k < 8; k++) s[k] = 0u;
        for (int j0 = 0; j0 < side; j0 += K) {
            uint32_t a0 = 0u, a1 = 0u, a2 = 0u, a3 = 0u;
            const int j1 = min(j0 + K, side);
#pragma unroll 4
            for (int j = j0; j < j1; j++) {
                const uint4 q = *reinterpret_cast<const uint4 *>(p + (size_t)j * kMedHist8Stride);
                a0 += q.x; a1 += q.y; a2 += q.z; a3 += q.w;
            }
            s[0] += a0 & 0x00FF00FFu; s[1] += (a0 >> 8) & 0x00FF00FFu;
            s[2] += a1 & 0x00FF00FFu; s[3] += (a1 >> 8) & 0x00FF00FFu;
            s[4] += a2 & 0x00FF00FFu; s[5] += (a2 >> 8) & 0x00FF00FFu;
            s[6] += a3 & 0x00FF00FFu; s[7] += (a3 >> 8) & 0x00FF00FFu;
        }
    };
    auto count = [](const uint32_t (&s)[8], int b) -> uint32_t {  // bin b of the 16
        return (s[2 * (b >> 2) + (b & 1)] >> (16 * ((b >> 1) & 1))) & 0xFFFFu;
    };
    for (int y = ys; y < ye; y++) {
        const int x = x0 + lane;
        uint32_t s[8];
        window16(base, s);  // coarse level
        uint32_t cum = 0, bin = 0, below = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const uint32_t cnt = count(s, b);
            cum += cnt;
            const bool le = cum <= target;  // bins wholly at or below the rank
            bin += le ? 1u : 0u;
            below += le ? cnt : 0u;
        }
        window16(base + 4 + bin * 4, s);  // fine level: the 16 counters of coarse bin `bin`
        const uint32_t target2 = target - below;
        uint32_t fcum = 0, fbin = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) {
            fcum += count(s, b);
            fbin += fcum <= target2 ? 1u : 0u;
        }
        if (x < w) {
            const size_t o = (size_t)y * w + x;
            uint32_t v = bin * 16 + fbin;
            if (mask && mask[o] == 0) v = (__ldg(src + o) >> shift) & 255u;
            dst[o * 4 + ch] = (uint8_t)v;
        }
        __syncwarp();  // every lane has read this row's histograms
        if (y + 1 < ye) {
            row_pixels(y - r, false);
            row_pixels(y + r + 1, true);
            __syncwarp();
        }
    }
}

}  // namespace

extern "C" int pfe_dev_box_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius,
                                const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "box_blur: bad args"));
    if (radius < 0.5f || radius != radius) return copy_through(ctx, src, dst, w, h);      // blur.rs:234
    if (src == dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "box_blur: in-place not supported");
    const float c = ceilf(radius);
    if (c > 20000.0f) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "box_blur: radius too large");
    const int r = (int)c;
    const size_t smem = ((size_t)32 * ((BOX_COLS + 2 * r) | 1) + 32 * (BOX_COLS + 1) + 64) * 4;
    if (smem > 220 * 1024) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "box_blur: radius too large for the H tile");
    const uint32_t dwin = 2u * (uint32_t)r + 1u;
    const uint32_t magic = dwin <= 4096u ? (uint32_t)((1ull << 32) / dwin) + 1u : 0u;
    void *hb;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, (size_t)w * h * 4, &hb));
    if (smem > 48 * 1024) PFE_CUDA(ctx, cudaFuncSetAttribute(box_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFE_KERNEL(ctx, "box_h", box_h_kernel<<<dim3(pfe_div_up(w, BOX_COLS), pfe_div_up(h, 32)), 128, smem, ctx->stream>>>(
        (const uint32_t *)src, (uint32_t *)hb, (int)w, (int)h, r, magic));
    PFE_LAUNCHED(ctx);
    PFE_KERNEL(ctx, "box_v", box_v_kernel<<<dim3(pfe_div_up(w, 128), pfe_div_up(h, BOX_BAND)), 128, 0, ctx->stream>>>(
        (const uint32_t *)hb, (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, r, magic));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_motion_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float angle_deg,
                                   float distance, const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "motion_blur: bad args"));
    if (distance < 1.0f || distance != distance) return copy_through(ctx, src, dst, w, h);  // blur.rs:150
    if (src == dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "motion_blur: in-place not supported");
    if (distance > 1.0e6f) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "motion_blur: distance too large");
    // host transcendental constants, exactly as the reference computes them (blur.rs:159-163)
    const float angle = angle_deg * (3.14159265358979323846f / 180.0f);  // f32::to_radians
    const int steps = (int)ceilf(distance);
    const float dx = cosf(angle), dy = sinf(angle);
    const float inv_steps = 1.0f / (float)(steps * 2 + 1);
    const dim3 grid(pfe_div_up(w, 32), pfe_div_up(h, 8));
    if (steps <= 128 && w < (1u << 21) && h < (1u << 21))
        PFE_KERNEL(ctx, "motion", motion_kernel<true><<<grid, 256, 0, ctx->stream>>>(
            (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, steps, dx, dy, inv_steps));
    else
        PFE_KERNEL(ctx, "motion", motion_kernel<false><<<grid, 256, 0, ctx->stream>>>(
            (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, steps, dx, dy, inv_steps));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_median(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius,
                              const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "median: bad args"));
    if (src == dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "median: in-place not supported");
    const int r = radius < 1 ? 1 : (int)std::min<uint32_t>(radius, 1u << 20);   // noise.rs:364
    if (r > 20000) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "median: radius above 20000");  // (2r+1)^2 leaves 32-bit counts
    // tuning aid: PFE_MEDIAN_KERNEL=bisect forces the bisection kernels (A/B against the oracle and for timing)
    const char *force = getenv("PFE_MEDIAN_KERNEL");
    const bool bisect_only = force && strcmp(force, "bisect") == 0;
    const dim3 tiles(pfe_div_up(w, 32), pfe_div_up(h, 8));
    if (r <= 2 && !bisect_only) {
        if (r == 1) PFE_KERNEL(ctx, "median_small", median_small_kernel<1><<<tiles, 256, 0, ctx->stream>>>((const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h));
        else PFE_KERNEL(ctx, "median_small", median_small_kernel<2><<<tiles, 256, 0, ctx->stream>>>((const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h));
    } else if (r <= kMedHistMaxR && !bisect_only) {
        const bool wide = force && strcmp(force, "hist16") == 0;  // the 16-bit counter layout (A/B)
        const size_t smem = (size_t)4 * (kMedHistTW + 2 * r) * (wide ? kMedHistStride : kMedHist8Stride) * sizeof(uint32_t);
        PFE_CUDA(ctx, cudaFuncSetAttribute(wide ? median_hist_kernel : median_hist8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // row segments: enough CTAs for every SM several times over, long enough that building the first window
        // (2r+1 rows) stays a small fraction of a segment
        const unsigned strips = pfe_div_up(w, kMedHistTW);
        unsigned seg = std::max<unsigned>(64u, 8u * (unsigned)(2 * r + 1));
        while (seg > 64u && (uint64_t)strips * pfe_div_up(h, seg) < (uint64_t)ctx->sm_count * 8) seg /= 2;
        if (wide)
            PFE_KERNEL(ctx, "median_hist", median_hist_kernel<<<dim3(strips, pfe_div_up(h, seg)), 128, smem, ctx->stream>>>(
                (const uint32_t *)src, mask, dst, (int)w, (int)h, r, (int)seg));
        else
            PFE_KERNEL(ctx, "median_hist", median_hist8_kernel<<<dim3(strips, pfe_div_up(h, seg)), 128, smem, ctx->stream>>>(
                (const uint32_t *)src, mask, dst, (int)w, (int)h, r, (int)seg));
    } else if (r <= 127 && (size_t)(MED_BX + 2 * r) * (MED_BY + 2 * r) * 4 <= 160 * 1024) {
        const size_t smem = (size_t)(MED_BX + 2 * r) * (MED_BY + 2 * r) * 4;
        if (smem > 48 * 1024) PFE_CUDA(ctx, cudaFuncSetAttribute(median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PFE_KERNEL(ctx, "median", median_kernel<<<tiles, MED_BX * MED_BY, smem, ctx->stream>>>(
            (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, r));
    } else {
        // any radius the reference accepts (it sorts whatever window it is given, noise.rs:403): 32-bit counts
        PFE_KERNEL(ctx, "median_global", median_global_kernel<<<tiles, 256, 0, ctx->stream>>>(
            (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, r));
    }
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_vignette(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, float softness,
                                const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "vignette: bad args"));
    const float fw = (float)w, fh = (float)h;
    const float cx = fw / 2.0f, cy = fh / 2.0f;
    const float max_dist = sqrtf(cx * cx + cy * cy);
    const float soft = fmaxf(softness, 0.01f);
    if (w % 4 == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0)
        PFE_KERNEL(ctx, "vignette", vignette_kernel<4><<<dim3(pfe_div_up(w, 128), pfe_div_up(h, 8)), 256, 0, ctx->stream>>>(
            (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, amount, soft, cx, cy, max_dist));
    else
        PFE_KERNEL(ctx, "vignette", vignette_kernel<1><<<dim3(pfe_div_up(w, 32), pfe_div_up(h, 8)), 256, 0, ctx->stream>>>(
            (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, amount, soft, cx, cy, max_dist));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}
