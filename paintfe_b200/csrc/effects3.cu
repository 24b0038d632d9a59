// Widened scope, part 2: every remaining effect of src/ops/effects/ that the reference pins with a
// golden (tests/visual_filters.rs): ink, oil painting, colour filter (effects/artistic.rs), contours
// (effects/contours.rs), crystallize, dents (effects/distort.rs), halftone (effects/stylize.rs), bokeh,
// zoom blur (effects/blur.rs), grid, canvas border, drop shadow, outline (effects/render.rs), pixel drag
// and RGB displace (effects/glitch.rs).  All are strict-f32 / integer restatements and bit-exact;
// transcendental constants (halftone / pixel-drag cos and sin) are computed on the host with libm as
// the reference does.  The drop shadow's blur is the library's Gaussian, so in the default (FMA) mode it
// inherits that kernel's <= 1 level tolerance and is bit-exact with PFE_GAUSS_EXACT.
#include <cstring>

#include "fx_common.cuh"

namespace {

__device__ __forceinline__ float ch(uint32_t v, int c) { return (float)((v >> (8 * c)) & 255u); }

// ---- ink_core, artistic.rs:31-99 ---------------------------------------------------------------
__device__ __forceinline__ float ink_lum(const uint32_t *src, int w, int h, int x, int y) {
    const uint32_t v = px_clamped(src, w, h, x, y);
    return 0.2126f * ch(v, 0) + 0.7152f * ch(v, 1) + 0.0722f * ch(v, 2);
}
__global__ void __launch_bounds__(256) ink_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                  float edge_strength, float threshold) {
    PFE_PIXEL_XY()
    const float l00 = ink_lum(src, w, h, x - 1, y - 1), l10 = ink_lum(src, w, h, x, y - 1), l20 = ink_lum(src, w, h, x + 1, y - 1);
    const float l01 = ink_lum(src, w, h, x - 1, y), l21 = ink_lum(src, w, h, x + 1, y);
    const float l02 = ink_lum(src, w, h, x - 1, y + 1), l12 = ink_lum(src, w, h, x, y + 1), l22 = ink_lum(src, w, h, x + 1, y + 1);
    const float gx = -l00 - 2.0f * l01 - l02 + l20 + 2.0f * l21 + l22;
    const float gy = -l00 - 2.0f * l10 - l20 + l02 + 2.0f * l12 + l22;
    const float edge = sqrtf(gx * gx + gy * gy) * edge_strength / 100.0f;
    const uint32_t val = edge > threshold ? 0u : 255u;
    dst[o] = pfe_pack(val, val, val, src[o] >> 24);
}

// ---- oil_painting_core, artistic.rs:123-217 ----------------------------------------------------
// One histogram per thread in shared memory, one 64-bit word per bin: count (10 bits, <= 441) and the
// three colour sums (18 bits each, <= 441*255) packed so a sample is a single 64-bit add.
constexpr int OIL_BX = 32, OIL_BY = 4;
__global__ void __launch_bounds__(OIL_BX *OIL_BY) oil_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w,
                                                              int h, int r, uint32_t nl) {
    extern __shared__ unsigned long long oil_sm[];
    const int tw = OIL_BX + 2 * r, th = OIL_BY + 2 * r, nt = OIL_BX * OIL_BY;
    unsigned long long *bins = oil_sm;                       // [nl][nt]
    uint32_t *tile = (uint32_t *)(oil_sm + (size_t)nl * nt);  // [th][tw]
    const int x0 = blockIdx.x * OIL_BX, y0 = blockIdx.y * OIL_BY;
    for (int idx = threadIdx.x; idx < tw * th; idx += nt) {
        const int ty = idx / tw, tx = idx - ty * tw;
        tile[idx] = px_clamped(src, w, h, x0 - r + tx, y0 - r + ty);
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    const uint32_t cv = tile[(ly + r) * tw + lx + r];
    if (mask && mask[o] == 0) { dst[o] = cv; return; }
    unsigned long long *mine = bins + threadIdx.x;
    for (uint32_t i = 0; i < nl; i++) mine[(size_t)i * nt] = 0ull;
    for (int dy = 0; dy <= 2 * r; dy++)
        for (int dx = 0; dx <= 2 * r; dx++) {
            const uint32_t p = tile[(ly + dy) * tw + lx + dx];
            const uint32_t pr = p & 255u, pg = (p >> 8) & 255u, pb = (p >> 16) & 255u;
            const uint32_t bin = min((pr + pg + pb) / 3u * nl / 256u, nl - 1u);
            mine[(size_t)bin * nt] += (1ull << 54) | ((unsigned long long)pr << 36) | ((unsigned long long)pg << 18) | pb;
        }
    uint32_t max_count = 0;
    unsigned long long best = 0ull;
    for (uint32_t i = 0; i < nl; i++) {
        const unsigned long long b = mine[(size_t)i * nt];
        const uint32_t c = (uint32_t)(b >> 54);
        if (c > max_count) { max_count = c; best = b; }
    }
    const uint32_t sr = (uint32_t)(best >> 36) & 0x3FFFFu, sg = (uint32_t)(best >> 18) & 0x3FFFFu, sb = (uint32_t)best & 0x3FFFFu;
    dst[o] = pfe_pack(sr / max_count, sg / max_count, sb / max_count, cv >> 24);
}

// ---- color_filter_core, artistic.rs:266-307 ----------------------------------------------------
__device__ __forceinline__ float cf_blend(int mode, float s, float f) {
    switch (mode) {
    case 0: return s * f;
    case 1: return 1.0f - (1.0f - s) * (1.0f - f);
    case 2: return s < 0.5f ? 2.0f * s * f : 1.0f - 2.0f * (1.0f - s) * (1.0f - f);
    default: return f < 0.5f ? s - (1.0f - 2.0f * f) * s * (1.0f - s) : s + (2.0f * f - 1.0f) * (sqrtf(s) - s);
    }
}
__global__ void __launch_bounds__(256) color_filter_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                           float f0, float f1, float f2, float intensity, int mode) {
    PFE_PIXEL_XY()
    const uint32_t v = src[o];
    const float fc[3] = {f0, f1, f2};
    uint32_t out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float s = ch(v, c) / 255.0f;
        out[c] = pfe_round_u8((s * (1.0f - intensity) + cf_blend(mode, s, fc[c]) * intensity) * 255.0f);
    }
    dst[o] = pfe_pack(out[0], out[1], out[2], v >> 24);
}

// ---- contours_core, contours.rs:56-112 ---------------------------------------------------------
__global__ void __launch_bounds__(256) contours_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                       float inv_scale, float freq, float half_lw, float lr, float lg, float lb,
                                                       float la, uint32_t seed, uint32_t oct, float blend) {
    PFE_PIXEL_XY()
    const uint32_t v = src[o];
    const float noise_val = turbulence_2d((float)x * inv_scale, (float)y * inv_scale, seed, oct, 0.5f);
    const float level = noise_val * freq;
    const float dist = fabsf(level - roundf(level)) / freq;
    const float edge = half_lw * inv_scale * 0.5f;
    const float line_alpha = dist < edge ? 1.0f : (dist < edge * 2.0f ? 1.0f - (dist - edge) / edge : 0.0f);
    const float alpha = line_alpha * la * blend;
    dst[o] = pfe_pack(pfe_round_u8(ch(v, 0) * (1.0f - alpha) + lr * alpha), pfe_round_u8(ch(v, 1) * (1.0f - alpha) + lg * alpha),
                      pfe_round_u8(ch(v, 2) * (1.0f - alpha) + lb * alpha), v >> 24);
}

// ---- crystallize_core, distort.rs:26-169 -------------------------------------------------------
__global__ void crystal_seeds_kernel(float2 *seeds, int cells_x, int cells_y, float cs, uint32_t seed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cells_x * cells_y) return;
    const int cy = i / cells_x, cx = i - cy * cells_x;
    const float jx = hash_f32((uint32_t)cx, (uint32_t)cy, seed), jy = hash_f32((uint32_t)cx, (uint32_t)cy, seed + 77u);
    seeds[i] = make_float2((float)cx * cs + jx * cs, (float)cy * cs + jy * cs);
}
__device__ __forceinline__ uint32_t crystal_cell(const float2 *seeds, int cells_x, int cells_y, float cs, int x, int y) {
    const int gcx = __float2int_rz((float)x / cs), gcy = __float2int_rz((float)y / cs);
    const float px = (float)x + 0.5f, py = (float)y + 0.5f;
    float best = 3.40282347e+38f;
    uint32_t best_idx = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            const int nx = gcx + dx, ny = gcy + dy;
            if (nx < 0 || ny < 0 || nx >= cells_x || ny >= cells_y) continue;
            const uint32_t idx = (uint32_t)(ny * cells_x + nx);
            const float2 s = __ldg(seeds + idx);
            const float d = (px - s.x) * (px - s.x) + (py - s.y) * (py - s.y);
            if (d < best) { best = d; best_idx = idx; }
        }
    return best_idx;
}
// Cell sums are integers (the reference's f64 sums of u8 values are exact), accumulated with one atomic per
// warp-level group of pixels that share a cell.
__global__ void __launch_bounds__(256) crystal_accum_kernel(const uint32_t *src, const float2 *seeds, unsigned long long *sums,
                                                            uint32_t *counts, int w, int h, float cs, int cells_x, int cells_y) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool valid = x < w && y < h;
    uint32_t idx = 0xFFFFFFFFu, v = 0;
    if (valid) { idx = crystal_cell(seeds, cells_x, cells_y, cs, x, y); v = src[(size_t)y * w + x]; }
    const unsigned grp = __match_any_sync(0xffffffffu, idx);
    const uint32_t s0 = __reduce_add_sync(grp, v & 255u), s1 = __reduce_add_sync(grp, (v >> 8) & 255u);
    const uint32_t s2 = __reduce_add_sync(grp, (v >> 16) & 255u), s3 = __reduce_add_sync(grp, v >> 24);
    if (valid && (int)(threadIdx.x & 31) == __ffs(grp) - 1) {
        atomicAdd(sums + (size_t)idx * 4, (unsigned long long)s0);
        atomicAdd(sums + (size_t)idx * 4 + 1, (unsigned long long)s1);
        atomicAdd(sums + (size_t)idx * 4 + 2, (unsigned long long)s2);
        atomicAdd(sums + (size_t)idx * 4 + 3, (unsigned long long)s3);
        atomicAdd(counts + idx, (uint32_t)__popc(grp));
    }
}
__global__ void crystal_avg_kernel(const unsigned long long *sums, const uint32_t *counts, uint32_t *avg, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t out[4] = {0, 0, 0, 0};
    if (counts[i] > 0) {
        const double inv = 1.0 / (double)counts[i];
#pragma unroll
        for (int c = 0; c < 4; c++) out[c] = (uint32_t)fmin(fmax(round((double)sums[(size_t)i * 4 + c] * inv), 0.0), 255.0);
    }
    avg[i] = pfe_pack(out[0], out[1], out[2], out[3]);
}
__global__ void __launch_bounds__(256) crystal_assign_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                             const float2 *seeds, const uint32_t *avg, float cs, int cells_x,
                                                             int cells_y) {
    PFE_PIXEL_XY()
    dst[o] = __ldg(avg + crystal_cell(seeds, cells_x, cells_y, cs, x, y));
}

// ---- dents_core, distort.rs:248-310 ------------------------------------------------------------
__device__ __forceinline__ float rem_euclid_f(float a, float b) {
    const float r = fmodf(a, b);
    return r < 0.0f ? r + fabsf(b) : r;
}
__global__ void __launch_bounds__(256) dents_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                    float scale, float inv_scale, float amount, uint32_t seed, uint32_t oct,
                                                    float roughness, int pinch, int wrap) {
    PFE_PIXEL_XY()
    float nx = turbulence_2d((float)x * inv_scale, (float)y * inv_scale, seed, oct, roughness) * 2.0f - 1.0f;
    float ny = turbulence_2d((float)x * inv_scale, (float)y * inv_scale, seed + 9999u, oct, roughness) * 2.0f - 1.0f;
    if (pinch) {
        const float cx = (float)w * 0.5f, cy = (float)h * 0.5f;
        const float dx = (float)x - cx, dy = (float)y - cy;
        const float dist = fmaxf(sqrtf(dx * dx + dy * dy), 1.0f);
        const float factor = (1.0f - dist / fmaxf(cx, cy)) * 0.5f;
        nx = nx + dx / dist * factor;
        ny = ny + dy / dist * factor;
    }
    float sx = (float)x + nx * amount * scale, sy = (float)y + ny * amount * scale;
    if (wrap) { sx = rem_euclid_f(sx, (float)w); sy = rem_euclid_f(sy, (float)h); }
    dst[o] = bilinear_round(src, w, h, sx, sy);
}

// ---- halftone_core, stylize.rs:242-277 ---------------------------------------------------------
__global__ void __launch_bounds__(256) halftone_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                       float ds, float cos_a, float sin_a, int shape) {
    PFE_PIXEL_XY()
    const uint32_t v = src[o];
    const float lum = (0.2126f * ch(v, 0) + 0.7152f * ch(v, 1) + 0.0722f * ch(v, 2)) / 255.0f;
    const float fx = (float)x * cos_a + (float)y * sin_a;
    const float fy = -((float)x) * sin_a + (float)y * cos_a;
    const float qx = fx / ds, qy = fy / ds;
    const float cx = fabsf(qx - truncf(qx)) - 0.5f, cy = fabsf(qy - truncf(qy)) - 0.5f;
    float thr;
    switch (shape) {
    case 0: thr = sqrtf(cx * cx + cy * cy) * 2.0f; break;
    case 1: thr = fmaxf(fabsf(cx), fabsf(cy)) * 2.0f; break;
    case 2: thr = fabsf(cx) + fabsf(cy); break;
    default: thr = fabsf(cy) * 2.0f; break;
    }
    const uint32_t val = thr < lum ? 255u : 0u;
    dst[o] = pfe_pack(val, val, val, v >> 24);
}

// ---- bokeh_blur_core, blur.rs:22-115 -----------------------------------------------------------
// The disc is one horizontal span per kernel row.  Pass 1 builds, per image row, the prefix sums of the
// clamp-extended row (E[j] = sum of src[clamp(i - r)] for i < j, j in [0, w + 2r]); pass 2 sums each
// span as a difference of two prefix entries: O(r) per pixel like the reference's sliding sums, and the
// same integers.
__global__ void __launch_bounds__(256) bokeh_prefix_kernel(const uint32_t *src, uint4 *E, int w, int h, int r) {
    const int lane = threadIdx.x & 31, y = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (y >= h) return;
    const int L = w + 2 * r;
    const uint32_t *row = src + (size_t)y * w;
    uint4 *e = E + (size_t)y * (L + 1);
    if (lane == 0) e[0] = make_uint4(0, 0, 0, 0);
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int base = 0; base < L; base += 32) {
        const int j = base + lane;
        const uint32_t v = j < L ? __ldg(row + pfe_clampi(j - r, 0, w - 1)) : 0u;
        uint32_t a0 = v & 255u, a1 = (v >> 8) & 255u, a2 = (v >> 16) & 255u, a3 = v >> 24;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            const uint32_t t0 = __shfl_up_sync(0xffffffffu, a0, s), t1 = __shfl_up_sync(0xffffffffu, a1, s);
            const uint32_t t2 = __shfl_up_sync(0xffffffffu, a2, s), t3 = __shfl_up_sync(0xffffffffu, a3, s);
            if (lane >= s) { a0 += t0; a1 += t1; a2 += t2; a3 += t3; }
        }
        a0 += c0; a1 += c1; a2 += c2; a3 += c3;
        if (j < L) e[j + 1] = make_uint4(a0, a1, a2, a3);
        c0 = __shfl_sync(0xffffffffu, a0, 31); c1 = __shfl_sync(0xffffffffu, a1, 31);
        c2 = __shfl_sync(0xffffffffu, a2, 31); c3 = __shfl_sync(0xffffffffu, a3, 31);
    }
}
__global__ void __launch_bounds__(256) bokeh_gather_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                           const uint4 *E, const int2 *spans, int ns, int r, float inv_count) {
    PFE_PIXEL_XY()
    const size_t pitch = (size_t)w + 2 * r + 1;
    unsigned long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
    for (int s = 0; s < ns; s++) {
        const int2 sp = __ldg(spans + s);
        const uint4 *e = E + (size_t)pfe_clampi(y + sp.x, 0, h - 1) * pitch;
        const uint4 hi = __ldg(e + (x + sp.y + r + 1)), lo = __ldg(e + (x - sp.y + r));
        t0 += hi.x - lo.x; t1 += hi.y - lo.y; t2 += hi.z - lo.z; t3 += hi.w - lo.w;
    }
    dst[o] = pfe_pack(pfe_round_u8(__ull2float_rn(t0) * inv_count), pfe_round_u8(__ull2float_rn(t1) * inv_count),
                      pfe_round_u8(__ull2float_rn(t2) * inv_count), pfe_round_u8(__ull2float_rn(t3) * inv_count));
}

// ---- zoom_blur_core, blur.rs:322-427 -----------------------------------------------------------
__device__ __forceinline__ int round_i32(float v) {  // `.round() as i32`: saturating, NaN -> 0
    const float r = roundf(v);
    return (r != r) ? 0 : __float2int_rz(fminf(fmaxf(r, -2147483648.0f), 2147483520.0f));
}
struct ZoomParams { float cx, cy, s, inv_n, max_dist, tint[4], tint_strength; uint32_t n; };
__global__ void __launch_bounds__(256) zoom_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                   ZoomParams p) {
    PFE_PIXEL_XY()
    const float dx = (float)x - p.cx, dy = (float)y - p.cy;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const float nm1 = (float)(p.n - 1);
    for (uint32_t i = 0; i < p.n; i++) {
        const float t = 1.0f - p.s * ((float)i / nm1);
        const uint32_t v = px_clamped(src, w, h, round_i32(p.cx + dx * t), round_i32(p.cy + dy * t));
        s0 += ch(v, 0); s1 += ch(v, 1); s2 += ch(v, 2); s3 += ch(v, 3);
    }
    float v[4] = {s0 * p.inv_n, s1 * p.inv_n, s2 * p.inv_n, s3 * p.inv_n};
    if (p.tint_strength > 0.001f) {
        const float dist = sqrtf(dx * dx + dy * dy);
        const float t = fmaxf(1.0f - dist / p.max_dist, 0.0f) * p.tint_strength;
#pragma unroll
        for (int c = 0; c < 4; c++) v[c] = v[c] + (p.tint[c] * 255.0f - v[c]) * t;
    }
    dst[o] = pfe_pack(pfe_round_u8(v[0]), pfe_round_u8(v[1]), pfe_round_u8(v[2]), pfe_round_u8(v[3]));
}

// ---- grid_core / canvas_border_core, render.rs:52-92, :114-165 ---------------------------------
__global__ void __launch_bounds__(256) grid_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                   uint32_t cw, uint32_t chh, uint32_t lw, uint32_t color, int style, float t) {
    PFE_PIXEL_XY()
    const uint32_t v = src[o];
    const bool draw = style == 0 ? (((uint32_t)x % cw) < lw || ((uint32_t)y % chh) < lw)
                                 : ((((uint32_t)x / cw) + ((uint32_t)y / chh)) % 2u == 0u);
    if (!draw) { dst[o] = v; return; }
    uint32_t out[4];
#pragma unroll
    for (int c = 0; c < 4; c++) out[c] = pfe_round_u8(ch(v, c) * (1.0f - t) + ch(color, c) * t);
    dst[o] = pfe_pack(out[0], out[1], out[2], out[3]);
}
__global__ void __launch_bounds__(256) border_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                     uint32_t bw, uint32_t color) {
    PFE_PIXEL_XY()
    const uint32_t xu = (uint32_t)x, yu = (uint32_t)y;
    const bool border = xu < bw || yu < bw || xu >= (uint32_t)w - bw || yu >= (uint32_t)h - bw;
    dst[o] = border ? color : src[o];
}

// ---- shadow_core, render.rs:220-352 ------------------------------------------------------------
// step 1 (:232-246): the source alpha, offset; written as a plane (for the spread pass) or replicated
// into the RGBA image the blur consumes (:299-302).
__global__ void __launch_bounds__(256) shadow_alpha_kernel(const uint32_t *src, uint8_t *plane, uint32_t *argba, int w, int h,
                                                           int ox, int oy) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const long long sx = (long long)x - ox, sy = (long long)y - oy;
    uint32_t a = 0;
    if (sx >= 0 && sx < w && sy >= 0 && sy < h) a = src[(size_t)sy * w + sx] >> 24;
    const size_t o = (size_t)y * w + x;
    if (plane) plane[o] = (uint8_t)a;
    else argba[o] = a * 0x01010101u;
}
// step 2 (:249-295): separable max over a window truncated at the image border
__global__ void __launch_bounds__(256) shadow_spread_kernel(const uint8_t *in, uint8_t *out_plane, uint32_t *argba, int w, int h,
                                                            int r, int vertical) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    uint32_t m = 0;
    if (!vertical) {
        const int a = max(x - r, 0), b = (int)min((long long)x + r, (long long)w - 1);
        for (int s = a; s <= b; s++) m = max(m, (uint32_t)in[(size_t)y * w + s]);
    } else {
        const int a = max(y - r, 0), b = (int)min((long long)y + r, (long long)h - 1);
        for (int s = a; s <= b; s++) m = max(m, (uint32_t)in[(size_t)s * w + x]);
    }
    const size_t o = (size_t)y * w + x;
    if (out_plane) out_plane[o] = (uint8_t)m;
    else argba[o] = m * 0x01010101u;
}
// step 4 (:313-349): shadow under, source over, straight alpha out
__global__ void __launch_bounds__(256) shadow_composite_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w,
                                                               int h, const uint32_t *blur, uint32_t color, float opacity) {
    PFE_PIXEL_XY()
    const uint32_t v = src[o];
    const float shadow_a = ((float)(blur[o] & 255u) / 255.0f) * opacity * (ch(color, 3) / 255.0f);
    const float src_a = ch(v, 3) / 255.0f;
    const float out_a = src_a + shadow_a * (1.0f - src_a);
    uint32_t out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float shadow_c = ch(color, c) / 255.0f, src_c = ch(v, c) / 255.0f;
        const float out_c = out_a > 0.0f ? (src_c * src_a + shadow_c * shadow_a * (1.0f - src_a)) / out_a : 0.0f;
        out[c] = pfe_round_u8(out_c * 255.0f);
    }
    dst[o] = pfe_pack(out[0], out[1], out[2], pfe_round_u8(out_a * 255.0f));
}

// ---- outline_core, render.rs:403-572 -----------------------------------------------------------
// bounding box of alpha > 0 (:421-436): bb = {min_x, min_y, max_x, max_y}
__global__ void alpha_bbox_kernel(const uint32_t *src, uint32_t w, uint32_t h, uint32_t *bb) {
    uint32_t mnx = 0xFFFFFFFFu, mny = 0xFFFFFFFFu, mxx = 0, mxy = 0;
    bool any = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)w * h; i += (size_t)gridDim.x * blockDim.x) {
        if ((src[i] >> 24) > 0) {
            const uint32_t y = (uint32_t)(i / w), x = (uint32_t)(i - (size_t)y * w);
            mnx = min(mnx, x); mny = min(mny, y); mxx = max(mxx, x); mxy = max(mxy, y);
            any = true;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    any = __any_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && any) {
        atomicMin(&bb[0], mnx); atomicMin(&bb[1], mny); atomicMax(&bb[2], mxx); atomicMax(&bb[3], mxy);
    }
}
__device__ __forceinline__ float outline_cov(float distance, float radius, int aa) {
    if (aa) {
        const float t = pfe_clampf((radius + 0.5f - distance) / 1.0f, 0.0f, 1.0f);
        return t * t * (3.0f - 2.0f * t);
    }
    return distance <= radius ? 1.0f : 0.0f;
}
__global__ void __launch_bounds__(256) outline_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                      const uint32_t *bb, float radius, int sr, uint32_t color, int mode, int aa) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    const uint32_t v = src[o];
    const uint32_t b0 = bb[0], b1 = bb[1], b2 = bb[2], b3 = bb[3];
    // nothing opaque (:437-439), outside the processed box (:440-443, :500-502) or deselected: unchanged
    const long long pad = (long long)sr + 1;
    if (b0 > b2 || x < (long long)b0 - pad || x > (long long)b2 + pad || y < (long long)b1 - pad || y > (long long)b3 + pad ||
        (mask && mask[o] == 0)) {
        dst[o] = v;
        return;
    }
    // nearest filled / empty sample in the (2 sr + 1)^2 window (:466-494): the smallest squared distance
    int best_f = 0x7FFFFFFF, best_e = 0x7FFFFFFF;
    const int ya = max(y - sr, 0), yb = min(y + sr, h - 1), xa = max(x - sr, 0), xb = min(x + sr, w - 1);
    for (int sy = ya; sy <= yb; sy++) {
        const int dy2 = (sy - y) * (sy - y);
        const uint32_t *row = src + (size_t)sy * w;
        for (int sx = xa; sx <= xb; sx++) {
            const int d = (sx - x) * (sx - x) + dy2;
            if ((__ldg(row + sx) >> 24) > 0) best_f = min(best_f, d);
            else best_e = min(best_e, d);
        }
    }
    const float src_a = ch(v, 3) / 255.0f;
    const float outside_cov = (best_f != 0x7FFFFFFF ? outline_cov(fmaxf(sqrtf((float)best_f) - 1.0f, 0.0f), radius, aa) : 0.0f) * (1.0f - src_a);
    const float inside_cov = (best_e != 0x7FFFFFFF ? outline_cov(sqrtf((float)best_e), radius, aa) : 0.0f) * src_a;
    const float under_cov = mode == 1 ? 0.0f : outside_cov, over_cov = mode == 0 ? 0.0f : inside_cov;
    const float ca = ch(color, 3) / 255.0f;
    const float a_under = ca * under_cov, a_over = ca * over_cov;
    float comp[3] = {ch(v, 0) / 255.0f, ch(v, 1) / 255.0f, ch(v, 2) / 255.0f};
    float comp_a = src_a;
    if (a_under > 0.0f) {
        const float out_a = comp_a + a_under * (1.0f - comp_a);
        if (out_a > 0.0f) {
#pragma unroll
            for (int c = 0; c < 3; c++) comp[c] = (comp[c] * comp_a + (ch(color, c) / 255.0f) * a_under * (1.0f - comp_a)) / out_a;
        }
        comp_a = out_a;
    }
    if (a_over > 0.0f) {
        const float out_a = a_over + comp_a * (1.0f - a_over);
        if (out_a > 0.0f) {
#pragma unroll
            for (int c = 0; c < 3; c++) comp[c] = ((ch(color, c) / 255.0f) * a_over + comp[c] * comp_a * (1.0f - a_over)) / out_a;
        }
        comp_a = out_a;
    }
    dst[o] = pfe_pack(pfe_as_u8(roundf(pfe_clampf(comp[0], 0.0f, 1.0f) * 255.0f)), pfe_as_u8(roundf(pfe_clampf(comp[1], 0.0f, 1.0f) * 255.0f)),
                      pfe_as_u8(roundf(pfe_clampf(comp[2], 0.0f, 1.0f) * 255.0f)), pfe_as_u8(roundf(pfe_clampf(comp_a, 0.0f, 1.0f) * 255.0f)));
}

// ---- pixel_drag_core / rgb_displace_core, glitch.rs:44-99, :142-197 ----------------------------
__global__ void __launch_bounds__(256) pixel_drag_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                         uint32_t seed, float amount, float dist, float dx_dir, float dy_dir) {
    PFE_PIXEL_XY()
    if (hash_f32((uint32_t)y, 0u, seed) > amount / 100.0f) { dst[o] = src[o]; return; }  // row not affected
    const float hd = hash_f32((uint32_t)y, 1u, seed) * dist;                              // `as i32`: >= 0 here
    const float drag = (float)__float2int_rz(fminf(hd, 2147483520.0f));
    dst[o] = px_clamped(src, w, h, round_i32((float)x - drag * dx_dir), round_i32((float)y - drag * dy_dir));
}
struct RgbOff { int v[6]; };
__global__ void __launch_bounds__(256) rgb_displace_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                           RgbOff off) {
    PFE_PIXEL_XY()
    uint32_t out[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const long long sx = min(max((long long)x + off.v[2 * c], 0ll), (long long)w - 1);
        const long long sy = min(max((long long)y + off.v[2 * c + 1], 0ll), (long long)h - 1);
        out[c] = (__ldg(src + (size_t)sy * w + sx) >> (8 * c)) & 255u;
    }
    dst[o] = pfe_pack(out[0], out[1], out[2], src[o] >> 24);
}

inline uint32_t rgba_word(const uint8_t c[4]) { return (uint32_t)c[0] | ((uint32_t)c[1] << 8) | ((uint32_t)c[2] << 16) | ((uint32_t)c[3] << 24); }
inline float to_radians(float deg) { return deg * (3.14159265358979323846f / 180.0f); }
inline int sat_i32(float v) { return (v != v) ? 0 : (v <= -2147483648.0f ? INT32_MIN : (v >= 2147483648.0f ? INT32_MAX : (int)v)); }
inline int copy_image(pfe_ctx *ctx, const uint8_t *src, uint8_t *dst, uint32_t w, uint32_t h) {
    PFE_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)w * h * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return PFE_OK;
}

}  // namespace

#define SRC32 (const uint32_t *)src
#define DST32 (uint32_t *)dst
#define WH (int)w, (int)h

extern "C" int pfe_dev_ink(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float edge_strength, float threshold,
                           const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "ink: bad args"));
    PFE_KERNEL(ctx, "ink", ink_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(SRC32, mask, DST32, WH, edge_strength, threshold));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_oil_painting(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius, uint32_t levels,
                                    const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "oil_painting: bad args"));
    const int r = (int)std::min(std::max(radius, 1u), 10u);
    const uint32_t nl = std::min(std::max(levels, 2u), 64u);
    const size_t smem = (size_t)nl * OIL_BX * OIL_BY * 8 + (size_t)(OIL_BX + 2 * r) * (OIL_BY + 2 * r) * 4;
    if (smem > 48 * 1024) PFE_CUDA(ctx, cudaFuncSetAttribute(oil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFE_KERNEL(ctx, "oil_painting", oil_kernel<<<dim3(pfe_div_up(w, OIL_BX), pfe_div_up(h, OIL_BY)), OIL_BX * OIL_BY, smem, ctx->stream>>>(
        SRC32, mask, DST32, WH, r, nl));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_color_filter(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const uint8_t color[4],
                                    float intensity, int mode, const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "color_filter: bad args"));
    if (!color || mode < 0 || mode > 3) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "color_filter: bad colour or mode");
    PFE_KERNEL(ctx, "color_filter", color_filter_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, (float)color[0] / 255.0f, (float)color[1] / 255.0f, (float)color[2] / 255.0f, intensity, mode));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_contours(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float scale, float frequency,
                                float line_width, const uint8_t color[4], uint32_t seed, uint32_t octaves, float blend,
                                const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "contours: bad args"));
    if (!color) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "contours: null colour");
    PFE_KERNEL(ctx, "contours", contours_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, 1.0f / fmaxf(scale, 0.5f), fmaxf(frequency, 0.5f), fmaxf(line_width * 0.5f, 0.3f), (float)color[0],
        (float)color[1], (float)color[2], (float)color[3] / 255.0f, seed, std::min(std::max(octaves, 1u), 8u), blend));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_crystallize(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float cell_size, uint32_t seed,
                                   const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "crystallize: bad args"));
    const float cs = fmaxf(cell_size, 2.0f);
    const int cells_x = std::max(sat_i32(ceilf((float)w / cs)), 1), cells_y = std::max(sat_i32(ceilf((float)h / cs)), 1);
    const size_t nc = (size_t)cells_x * cells_y;
    // seeds (8 B) | sums (32 B) | counts (4 B) | averages (4 B) per cell
    void *base;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, nc * 48, &base));
    unsigned long long *sums = (unsigned long long *)base;
    float2 *seeds = (float2 *)(sums + nc * 4);
    uint32_t *counts = (uint32_t *)(seeds + nc), *avg = counts + nc;
    PFE_CUDA(ctx, cudaMemsetAsync(sums, 0, nc * 32, ctx->stream));
    PFE_CUDA(ctx, cudaMemsetAsync(counts, 0, nc * 4, ctx->stream));
    PFE_KERNEL(ctx, "crystal_seeds", crystal_seeds_kernel<<<pfe_div_up(nc, 256), 256, 0, ctx->stream>>>(seeds, cells_x, cells_y, cs, seed));
    PFE_LAUNCHED(ctx);
    PFE_KERNEL(ctx, "crystal_accum", crystal_accum_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(SRC32, seeds, sums, counts, WH, cs, cells_x, cells_y));
    PFE_LAUNCHED(ctx);
    PFE_KERNEL(ctx, "crystal_avg", crystal_avg_kernel<<<pfe_div_up(nc, 256), 256, 0, ctx->stream>>>(sums, counts, avg, (int)nc));
    PFE_LAUNCHED(ctx);
    PFE_KERNEL(ctx, "crystal_assign", crystal_assign_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(SRC32, mask, DST32, WH, seeds, avg, cs, cells_x, cells_y));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_dents(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float scale, float amount, uint32_t seed,
                             uint32_t octaves, float roughness, int pinch, int wrap, const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "dents: bad args"));
    PFE_KERNEL(ctx, "dents", dents_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, scale, 1.0f / fmaxf(scale, 0.5f), amount, seed, std::min(std::max(octaves, 1u), 8u), roughness,
        pinch ? 1 : 0, wrap ? 1 : 0));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_halftone(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float dot_size, float angle_deg,
                                int shape, const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "halftone: bad args"));
    if (shape < 0 || shape > 3) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "halftone: bad shape");
    const float angle = to_radians(angle_deg);
    PFE_KERNEL(ctx, "halftone", halftone_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, fmaxf(dot_size, 2.0f), cosf(angle), sinf(angle), shape));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_bokeh_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius, const uint8_t *mask,
                                  uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "bokeh_blur: bad args"));
    if (!(radius >= 0.5f)) return copy_image(ctx, src, dst, w, h);  // blur.rs:23-25
    if (radius > 4000.0f) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "bokeh_blur: radius too large");
    const int r = (int)ceilf(radius);
    const float r2 = radius * radius;
    std::vector<int> spans;
    uint32_t sample_count = 0;
    for (int dy = -r; dy <= r; dy++) {
        const float remaining = r2 - (float)(dy * dy);
        if (remaining >= 0.0f) {
            const int span = (int)floorf(sqrtf(remaining));
            spans.push_back(dy);
            spans.push_back(span);
            sample_count += (uint32_t)(span * 2 + 1);
        }
    }
    void *sp_dev, *E;
    PFE_TRY(pfe_small_upload(ctx, spans.data(), spans.size() * sizeof(int), &sp_dev));
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_F32, ((size_t)w + 2 * r + 1) * h * sizeof(uint4), &E));
    PFE_KERNEL(ctx, "bokeh_prefix", bokeh_prefix_kernel<<<pfe_div_up(h, 8), 256, 0, ctx->stream>>>(SRC32, (uint4 *)E, WH, r));
    PFE_LAUNCHED(ctx);
    PFE_KERNEL(ctx, "bokeh_gather", bokeh_gather_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, (const uint4 *)E, (const int2 *)sp_dev, (int)(spans.size() / 2), r, 1.0f / (float)sample_count));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_zoom_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float center_x, float center_y,
                                 float strength, uint32_t samples, const float tint[4], float tint_strength, const uint8_t *mask,
                                 uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "zoom_blur: bad args"));
    if (strength < 0.001f) return copy_image(ctx, src, dst, w, h);  // blur.rs:332-334
    ZoomParams p;
    p.cx = center_x * (float)w;
    p.cy = center_y * (float)h;
    p.s = strength < 0.0f ? 0.0f : (strength > 0.99f ? 0.99f : strength);
    p.n = std::max(samples, 2u);
    p.inv_n = 1.0f / (float)p.n;
    const float ex[4] = {p.cx, (float)w - p.cx, p.cx, (float)w - p.cx}, ey[4] = {p.cy, p.cy, (float)h - p.cy, (float)h - p.cy};
    float md = 0.0f;
    for (int i = 0; i < 4; i++) md = fmaxf(md, sqrtf(ex[i] * ex[i] + ey[i] * ey[i]));
    p.max_dist = fmaxf(md, 1.0f);
    for (int c = 0; c < 4; c++) p.tint[c] = tint ? tint[c] : 0.0f;
    p.tint_strength = tint ? tint_strength : 0.0f;
    PFE_KERNEL(ctx, "zoom_blur", zoom_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(SRC32, mask, DST32, WH, p));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_grid(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t cell_w, uint32_t cell_h,
                            uint32_t line_width, const uint8_t color[4], int style, float opacity, const uint8_t *mask,
                            uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "grid: bad args"));
    if (!color || style < 0 || style > 1) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "grid: bad colour or style");
    PFE_KERNEL(ctx, "grid", grid_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, std::max(cell_w, 2u), std::max(cell_h, 2u), std::max(line_width, 1u), rgba_word(color), style, opacity));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_canvas_border(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t width,
                                     const uint8_t color[4], const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "canvas_border: bad args"));
    if (!color) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "canvas_border: null colour");
    PFE_KERNEL(ctx, "canvas_border", border_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, std::min(std::max(width, 1u), std::min(w, h)), rgba_word(color)));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_drop_shadow(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int32_t offset_x, int32_t offset_y,
                                   float blur_radius, int widen_radius, const uint8_t color[4], float opacity,
                                   const uint8_t *mask, uint8_t *dst, uint32_t flags) {
    PFE_TRY(check(ctx, src, dst, w, h, "drop_shadow: bad args"));
    if (!color) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "drop_shadow: null colour");
    const size_t n = (size_t)w * h;
    void *img, *planes;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, n * 8, &img));
    uint32_t *argba = (uint32_t *)img, *blurred = argba + n;
    const int spread = widen_radius ? sat_i32(roundf(fmaxf(blur_radius, 1.0f))) : 0;
    if (spread > 0) {
        PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, n * 2, &planes));
        uint8_t *pa = (uint8_t *)planes, *pb = pa + n;
        PFE_KERNEL(ctx, "shadow_alpha", shadow_alpha_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(SRC32, pa, nullptr, WH, offset_x, offset_y));
        PFE_LAUNCHED(ctx);
        PFE_KERNEL(ctx, "shadow_spread_h", shadow_spread_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(pa, pb, nullptr, WH, spread, 0));
        PFE_LAUNCHED(ctx);
        PFE_KERNEL(ctx, "shadow_spread_v", shadow_spread_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(pb, nullptr, argba, WH, spread, 1));
        PFE_LAUNCHED(ctx);
    } else {
        PFE_KERNEL(ctx, "shadow_alpha", shadow_alpha_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(SRC32, nullptr, argba, WH, offset_x, offset_y));
        PFE_LAUNCHED(ctx);
    }
    const uint32_t *blur = argba;
    if (blur_radius > 0.5f) {
        PFE_TRY(pfe_gauss_region(ctx, (const uint8_t *)argba, (uint8_t *)blurred, w, 0, 0, w, h, blur_radius, flags));
        blur = blurred;
    }
    PFE_KERNEL(ctx, "shadow_composite", shadow_composite_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, blur, rgba_word(color), opacity));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_outline(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t width, const uint8_t color[4],
                               int mode, int anti_alias, const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "outline: bad args"));
    if (!color || mode < 0 || mode > 2) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "outline: bad colour or mode");
    if (width > 512) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "outline: width too large");  // the search is O(width^2) per pixel
    const float radius = (float)std::max(width, 1u);
    const int sr = (int)ceilf(radius) + 1;
    void *bbd;
    const uint32_t init[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u};
    PFE_TRY(pfe_small_upload(ctx, init, sizeof(init), &bbd));
    PFE_KERNEL(ctx, "alpha_bbox", alpha_bbox_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(SRC32, w, h, (uint32_t *)bbd));
    PFE_LAUNCHED(ctx);
    PFE_KERNEL(ctx, "outline", outline_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, (const uint32_t *)bbd, radius, sr, rgba_word(color), mode, anti_alias ? 1 : 0));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_pixel_drag(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t seed, float amount,
                                  uint32_t distance, float direction, const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "pixel_drag: bad args"));
    const float dir = to_radians(direction);
    PFE_KERNEL(ctx, "pixel_drag", pixel_drag_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        SRC32, mask, DST32, WH, seed, amount, (float)std::max(distance, 1u), cosf(dir), sinf(dir)));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_rgb_displace(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const int32_t offsets[6],
                                    const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "rgb_displace: bad args"));
    if (!offsets) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "rgb_displace: null offsets");
    RgbOff off;
    for (int i = 0; i < 6; i++) off.v[i] = offsets[i];
    PFE_KERNEL(ctx, "rgb_displace", rgb_displace_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(SRC32, mask, DST32, WH, off));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}
