// Device and host helpers shared by the effect translation units (effects2.cu, effects3.cu):
// apply_per_pixel's thread mapping and mask rule (src/ops/effects.rs:53-100), sample_clamped /
// sample_bilinear (:109-141), hash_u32 / hash_f32 (:143-161), perlin_noise_2d (effects/noise.rs:52-71),
// turbulence_2d (effects/distort.rs:229-246).
#pragma once
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

#define PFE_PIXEL_XY()                                                                                   \
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);         \
    if (x >= w || y >= h) return;                                                                        \
    const size_t o = (size_t)y * w + x;                                                                  \
    if (mask && mask[o] == 0) { dst[o] = src[o]; return; }

// sample_clamped / sample_bilinear, src/ops/effects.rs:109-141
__device__ __forceinline__ uint32_t px_clamped(const uint32_t *src, int w, int h, int x, int y) {
    return __ldg(src + (size_t)pfe_clampi(y, 0, h - 1) * w + pfe_clampi(x, 0, w - 1));
}
__device__ __forceinline__ uint32_t bilinear_round(const uint32_t *src, int w, int h, float fx, float fy) {
    const float flx = floorf(fx), fly = floorf(fy);
    // `floor() as i32` saturates; NaN -> 0
    const int x0 = (flx != flx) ? 0 : __float2int_rz(fminf(fmaxf(flx, -2147483648.0f), 2147483520.0f));
    const int y0 = (fly != fly) ? 0 : __float2int_rz(fminf(fmaxf(fly, -2147483648.0f), 2147483520.0f));
    const int x1 = x0 == 2147483647 ? x0 : x0 + 1, y1 = y0 == 2147483647 ? y0 : y0 + 1;
    const float dx = fx - (float)x0, dy = fy - (float)y0;
    const uint32_t p00 = px_clamped(src, w, h, x0, y0), p10 = px_clamped(src, w, h, x1, y0);
    const uint32_t p01 = px_clamped(src, w, h, x0, y1), p11 = px_clamped(src, w, h, x1, y1);
    uint32_t out[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const float a = (float)((p00 >> (8 * c)) & 255u), b = (float)((p10 >> (8 * c)) & 255u);
        const float cc = (float)((p01 >> (8 * c)) & 255u), d = (float)((p11 >> (8 * c)) & 255u);
        const float v = a * (1.0f - dx) * (1.0f - dy) + b * dx * (1.0f - dy) + cc * (1.0f - dx) * dy + d * dx * dy;
        out[c] = pfe_round_u8(v);
    }
    return pfe_pack(out[0], out[1], out[2], out[3]);
}

// hash_u32 / hash_f32, src/ops/effects.rs:143-161
__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
    x *= 0x9E3779B9u; x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float hash_f32(uint32_t x, uint32_t y, uint32_t seed) {
    return (float)(hash_u32(x * 374761393u + y * 668265263u + seed) & 0x00FFFFFFu) / 16777216.0f;
}
// perlin_noise_2d, noise.rs:52-71; turbulence_2d, distort.rs:229-246
__device__ __forceinline__ float perlin_noise_2d(float x, float y, uint32_t seed) {
    const int xi = __float2int_rd(x), yi = __float2int_rd(y);
    const float xf = x - (float)xi, yf = y - (float)yi;
    const float u = xf * xf * xf * (xf * (xf * 6.0f - 15.0f) + 10.0f), v = yf * yf * yf * (yf * (yf * 6.0f - 15.0f) + 10.0f);
    const float n00 = hash_f32((uint32_t)xi, (uint32_t)yi, seed), n10 = hash_f32((uint32_t)(xi + 1), (uint32_t)yi, seed);
    const float n01 = hash_f32((uint32_t)xi, (uint32_t)(yi + 1), seed), n11 = hash_f32((uint32_t)(xi + 1), (uint32_t)(yi + 1), seed);
    const float nx0 = n00 + u * (n10 - n00), nx1 = n01 + u * (n11 - n01);
    return nx0 + v * (nx1 - nx0);
}
__device__ __forceinline__ float turbulence_2d(float x, float y, uint32_t seed, uint32_t octaves, float roughness) {
    float total = 0.0f, amplitude = 1.0f, frequency = 1.0f, max_amplitude = 0.0f;
    for (uint32_t i = 0; i < octaves; i++) {
        total += perlin_noise_2d(x * frequency, y * frequency, seed + i * 1000u) * amplitude;
        max_amplitude += amplitude;
        amplitude *= roughness;
        frequency *= 2.0f;
    }
    return max_amplitude > 0.0f ? total / max_amplitude : 0.0f;
}

inline int check(pfe_ctx *ctx, const void *src, const void *dst, uint32_t w, uint32_t h, const char *what) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !dst || !w || !h || src == dst || w > 0x7FFFFFFFu / 4 || h > 0x7FFFFFFFu / 4) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, what);
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    return PFE_OK;
}
inline dim3 grid2d(uint32_t w, uint32_t h) { return dim3(pfe_div_up(w, 32), pfe_div_up(h, 8)); }
inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }


}  // namespace
