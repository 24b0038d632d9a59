// Box blur, motion blur, median and vignette.
// Reference: src/ops/effects/blur.rs:144-318, src/ops/effects/noise.rs:357-410,
// src/ops/effects/stylize.rs:170-191 (+ apply_per_pixel, src/ops/effects.rs:53-100).
// All four are HBM-light (8 algorithmic bytes per pixel); box and median are integer and
// bit-exact by construction, motion blur sums integers in f32 (exact), vignette is strict f32.
#include <algorithm>
#include <cmath>
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
__device__ __forceinline__ uint32_t box_out(const uint4 &s, uint32_t d) {  // (sum + d/2) / d, blur.rs:269
    uint32_t h = d / 2;
    return pfe_pack((s.x + h) / d, (s.y + h) / d, (s.z + h) / d, (s.w + h) / d);
}

// ---- box blur H pass: block = 32 rows x 128 output columns; lane = row, warp = 32-column run.
// The clamped input tile is staged with coalesced loads; each thread slides a running window
// sum along its run (sums are windows over clamped indices, identical to the reference's
// incremental add/remove at blur.rs:272-278).
constexpr int BOX_COLS = 128;
__global__ void __launch_bounds__(128) box_h_kernel(const uint32_t *src, uint32_t *dst, int w, int h, int r) {
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
        tout[lane * (BOX_COLS + 1) + warp * 32 + c] = box_out(s, d);
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
                                                    uint32_t *dst, int w, int h, int r) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const int y0 = blockIdx.y * BOX_BAND, y1 = min(y0 + BOX_BAND, h);
    const uint32_t d = 2u * (uint32_t)r + 1u;
    uint4 s = make_uint4(0, 0, 0, 0);
    for (int k = -r; k <= r; k++) add4(s, __ldg(hb + (size_t)pfe_clampi(y0 + k, 0, h - 1) * w + x));
    for (int y = y0; y < y1; y++) {
        size_t o = (size_t)y * w + x;
        dst[o] = (mask && mask[o] == 0) ? src[o] : box_out(s, d);            // blur.rs:298-306
        sub4(s, __ldg(hb + (size_t)pfe_clampi(y - r, 0, h - 1) * w + x));
        add4(s, __ldg(hb + (size_t)pfe_clampi(y + r + 1, 0, h - 1) * w + x));
    }
}

// ---- motion blur, blur.rs:144-210 ---------------------------------------------------------
__global__ void __launch_bounds__(256) motion_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst,
                                                     int w, int h, int steps, float dx, float dy, float inv_steps) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    if (mask && mask[o] == 0) { dst[o] = src[o]; return; }
    float sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f;
    const float fx = (float)x, fy = (float)y;
    for (int i = -steps; i <= steps; i++) {
        // (x as f32 + i as f32 * dx).round() as i32, then clamp (:196-199)
        float px = roundf(fx + (float)i * dx), py = roundf(fy + (float)i * dy);
        int sx = pfe_clampi(__float2int_rz(px), 0, w - 1), sy = pfe_clampi(__float2int_rz(py), 0, h - 1);
        uint32_t v = __ldg(src + (size_t)sy * w + sx);
        sr += (float)(v & 255u); sg += (float)((v >> 8) & 255u); sb += (float)((v >> 16) & 255u); sa += (float)(v >> 24);
    }
    dst[o] = pfe_pack(pfe_round_u8(sr * inv_steps), pfe_round_u8(sg * inv_steps), pfe_round_u8(sb * inv_steps),
                      pfe_round_u8(sa * inv_steps));
}

// ---- median, noise.rs:357-410 --------------------------------------------------------------
// sorted[len/2] per channel == the largest v with #(x < v) <= len/2.  Found by an 8-step bitwise
// bisection run on all four channels at once with packed byte compares; per-channel counts live in
// 16-bit lanes (window <= 255x255).
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
__global__ void __launch_bounds__(256) vignette_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w,
                                                       int h, float amount, float soft, float cx, float cy,
                                                       float max_dist) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    const uint32_t v = src[o];
    if (mask && mask[o] == 0) { dst[o] = v; return; }
    float dx = (float)x - cx, dy = (float)y - cy;
    float dist = sqrtf(dx * dx + dy * dy) / max_dist;
    float q = fminf(dist / soft, 1.0f);
    float vf = pfe_clampf(1.0f - (amount * (q * q)), 0.0f, 1.0f);  // powf(2.0) == x*x exactly
    dst[o] = pfe_pack(pfe_round_u8((float)(v & 255u) * vf), pfe_round_u8((float)((v >> 8) & 255u) * vf),
                      pfe_round_u8((float)((v >> 16) & 255u) * vf), v >> 24);
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
    void *hb;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, (size_t)w * h * 4, &hb));
    if (smem > 48 * 1024) PFE_CUDA(ctx, cudaFuncSetAttribute(box_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFE_KERNEL(ctx, "box_h", box_h_kernel<<<dim3(pfe_div_up(w, BOX_COLS), pfe_div_up(h, 32)), 128, smem, ctx->stream>>>(
        (const uint32_t *)src, (uint32_t *)hb, (int)w, (int)h, r));
    PFE_LAUNCHED(ctx);
    PFE_KERNEL(ctx, "box_v", box_v_kernel<<<dim3(pfe_div_up(w, 128), pfe_div_up(h, BOX_BAND)), 128, 0, ctx->stream>>>(
        (const uint32_t *)hb, (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, r));
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
    PFE_KERNEL(ctx, "motion", motion_kernel<<<dim3(pfe_div_up(w, 32), pfe_div_up(h, 8)), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, steps, dx, dy, inv_steps));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_median(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius,
                              const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "median: bad args"));
    if (src == dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "median: in-place not supported");
    const int r = radius < 1 ? 1 : (int)std::min<uint32_t>(radius, 1u << 20);   // noise.rs:364
    if (r > 127) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "median: radius > 127");
    const size_t smem = (size_t)(MED_BX + 2 * r) * (MED_BY + 2 * r) * 4;
    if (smem <= 160 * 1024) {
        if (smem > 48 * 1024) PFE_CUDA(ctx, cudaFuncSetAttribute(median_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PFE_KERNEL(ctx, "median", median_kernel<<<dim3(pfe_div_up(w, MED_BX), pfe_div_up(h, MED_BY)), MED_BX * MED_BY, smem, ctx->stream>>>(
            (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, r));
    } else {
        PFE_KERNEL(ctx, "median_global", median_global_kernel<<<dim3(pfe_div_up(w, 32), pfe_div_up(h, 8)), 256, 0, ctx->stream>>>(
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
    PFE_KERNEL(ctx, "vignette", vignette_kernel<<<dim3(pfe_div_up(w, 32), pfe_div_up(h, 8)), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, amount, soft, cx, cy, max_dist));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}
