// Widened scope: the remaining Rhai Effect-API kernels.
// pixelate / bulge / twist (src/ops/effects/distort.rs:333-493), add_noise / reduce_noise
// (src/ops/effects/noise.rs:52-262), via apply_per_pixel rounding (src/ops/effects.rs:53-100).
// pixelate, bulge, uniform and Perlin noise are strict-f32 / integer and bit-exact; twist,
// Gaussian noise and the bilateral filter need sin/cos/ln/exp per pixel, evaluated in f64 and
// rounded once (the correctly rounded f32 value a <1-ulp libm returns on all but rare inputs),
// so those three are held to +-1 level.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "fx_common.cuh"

namespace {

__global__ void __launch_bounds__(256) pixelate_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                       uint32_t bs) {
    PFE_PIXEL_XY()
    const uint32_t bx = ((uint32_t)x / bs) * bs + bs / 2, by = ((uint32_t)y / bs) * bs + bs / 2;
    dst[o] = __ldg(src + (size_t)min(by, (uint32_t)h - 1) * w + min(bx, (uint32_t)w - 1));
}

__global__ void __launch_bounds__(256) bulge_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                    float amount, float strength, float cx, float cy, float max_r) {
    PFE_PIXEL_XY()
    const float dx = (float)x - cx, dy = (float)y - cy;
    const float dist = sqrtf(dx * dx + dy * dy);
    const float norm = fminf(dist / max_r, 1.0f);
    if (norm >= 1.0f) { dst[o] = src[o]; return; }                        // distort.rs:418-421
    const float falloff = 1.0f - norm;
    const float factor = amount > 0.0f ? 1.0f - falloff * strength * 0.5f : (amount < 0.0f ? 1.0f + falloff * strength * 0.5f : 1.0f);
    dst[o] = bilinear_round(src, w, h, cx + dx * factor, cy + dy * factor);
}

__global__ void __launch_bounds__(256) twist_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                    float twist_amount, float cx, float cy, float max_r) {
    PFE_PIXEL_XY()
    const float dx = (float)x - cx, dy = (float)y - cy;
    const float dist = sqrtf(dx * dx + dy * dy);
    const float norm = dist / max_r;
    const float rotation = twist_amount * (1.0f - norm);
    const float cr = (float)cos((double)rotation), sr = (float)sin((double)rotation);
    dst[o] = bilinear_round(src, w, h, cx + dx * cr - dy * sr, cy + dx * sr + dy * cr);
}

__global__ void __launch_bounds__(256) noise_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst, int w, int h,
                                                    float strength, int noise_type, int monochrome, uint32_t seed,
                                                    float inv_scale, uint32_t oct) {
    PFE_PIXEL_XY()
    const uint32_t v = src[o];
    const float r = (float)(v & 255u), g = (float)((v >> 8) & 255u), b = (float)((v >> 16) & 255u);
    const float sx = (float)x * inv_scale, sy = (float)y * inv_scale;
    const uint32_t qx = (uint32_t)__float2int_rd(sx), qy = (uint32_t)__float2int_rd(sy);  // coordinates are >= 0
    float nr, ng, nb;
    if (monochrome) {
        float nv;
        if (noise_type == 0) nv = hash_f32(qx, qy, seed) * 2.0f - 1.0f;
        else if (noise_type == 1) {
            const float u1 = fmaxf(hash_f32(qx, qy, seed), 0.0001f), u2 = hash_f32(qx, qy, seed + 7u);
            const float ln = (float)log((double)u1);
            const float cs = (float)cos((double)(2.0f * 3.14159265358979323846f * u2));
            nv = sqrtf(-2.0f * ln) * cs * 0.33f;
        } else nv = turbulence_2d(sx, sy, seed, oct, 0.5f) * 2.0f - 1.0f;
        nr = ng = nb = nv * strength;
    } else if (noise_type == 2) {
        nr = (turbulence_2d(sx, sy, seed, oct, 0.5f) * 2.0f - 1.0f) * strength;
        ng = (turbulence_2d(sx, sy, seed + 1u, oct, 0.5f) * 2.0f - 1.0f) * strength;
        nb = (turbulence_2d(sx, sy, seed + 2u, oct, 0.5f) * 2.0f - 1.0f) * strength;
    } else {  // non-monochrome Uniform and Gaussian both use the per-channel uniform hash (noise.rs:114-136)
        nr = (hash_f32(qx, qy, seed) * 2.0f - 1.0f) * strength;
        ng = (hash_f32(qx, qy, seed + 1u) * 2.0f - 1.0f) * strength;
        nb = (hash_f32(qx, qy, seed + 2u) * 2.0f - 1.0f) * strength;
    }
    dst[o] = pfe_pack(pfe_round_u8(r + nr), pfe_round_u8(g + ng), pfe_round_u8(b + nb), v >> 24);
}

// reduce_noise_core (bilateral), noise.rs:172-262: window in shared memory, taps in the reference's
// dy-outer / dx-inner order so the f32 sums associate identically.
constexpr int BIL_BX = 32, BIL_BY = 8;
__global__ void __launch_bounds__(BIL_BX *BIL_BY) bilateral_kernel(const uint32_t *src, const uint8_t *mask, uint32_t *dst,
                                                                   int w, int h, int r, float two_ss2, float two_sr2) {
    extern __shared__ uint32_t sm[];
    const int tw = BIL_BX + 2 * r, th = BIL_BY + 2 * r;
    const int x0 = blockIdx.x * BIL_BX, y0 = blockIdx.y * BIL_BY;
    for (int idx = threadIdx.x; idx < tw * th; idx += blockDim.x) {
        int ty = idx / tw, tx = idx - ty * tw;
        sm[idx] = __ldg(src + (size_t)pfe_clampi(y0 - r + ty, 0, h - 1) * w + pfe_clampi(x0 - r + tx, 0, w - 1));
    }
    __syncthreads();
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= w || y >= h) return;
    const size_t o = (size_t)y * w + x;
    const uint32_t cv = sm[(ly + r) * tw + lx + r];
    if (mask && mask[o] == 0) { dst[o] = cv; return; }
    const float cr = (float)(cv & 255u), cg = (float)((cv >> 8) & 255u), cb = (float)((cv >> 16) & 255u);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, wsum = 0.f;
    for (int dy = -r; dy <= r; dy++)
        for (int dx = -r; dx <= r; dx++) {
            const uint32_t p = sm[(ly + r + dy) * tw + lx + r + dx];
            const float pr = (float)(p & 255u), pg = (float)((p >> 8) & 255u), pb = (float)((p >> 16) & 255u), pa = (float)(p >> 24);
            const float spatial = (float)(dx * dx + dy * dy) / two_ss2;
            const float dr = cr - pr, dg = cg - pg, db = cb - pb;
            const float range = (dr * dr + dg * dg + db * db) / two_sr2;
            const float wt = (float)exp((double)(-spatial - range));
            s0 += pr * wt; s1 += pg * wt; s2 += pb * wt; s3 += pa * wt;
            wsum += wt;
        }
    if (wsum > 0.0f) {
        const float inv = 1.0f / wsum;
        dst[o] = pfe_pack(pfe_round_u8(s0 * inv), pfe_round_u8(s1 * inv), pfe_round_u8(s2 * inv), pfe_round_u8(s3 * inv));
    } else dst[o] = cv;
}

}  // namespace

extern "C" int pfe_dev_pixelate(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t block_size,
                                const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "pixelate: bad args"));
    PFE_KERNEL(ctx, "pixelate", pixelate_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, std::max(block_size, 2u)));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_bulge(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, float origin_x,
                             float origin_y, const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "bulge: bad args"));
    const float fw = (float)w, fh = (float)h;
    const float cx = clamp01(origin_x) * fmaxf(fw - 1.0f, 0.0f), cy = clamp01(origin_y) * fmaxf(fh - 1.0f, 0.0f);
    const float max_r = fmaxf(fmaxf(fmaxf(cx, fw - cx), fmaxf(cy, fh - cy)), 1.0f);
    PFE_KERNEL(ctx, "bulge", bulge_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, amount, fmaxf(fabsf(amount), 0.0001f), cx, cy, max_r));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_twist(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float angle_deg, float origin_x,
                             float origin_y, const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "twist: bad args"));
    const float fw = (float)w, fh = (float)h;
    const float cx = clamp01(origin_x) * fmaxf(fw - 1.0f, 0.0f), cy = clamp01(origin_y) * fmaxf(fh - 1.0f, 0.0f);
    const float mx = fmaxf(cx, fw - cx), my = fmaxf(cy, fh - cy);
    const float max_r = fmaxf(sqrtf(mx * mx + my * my), 1.0f);
    const float twist_amount = angle_deg * (3.14159265358979323846f / 180.0f);
    PFE_KERNEL(ctx, "twist", twist_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, twist_amount, cx, cy, max_r));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_add_noise(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, int noise_type,
                                 int monochrome, uint32_t seed, float scale, uint32_t octaves, const uint8_t *mask,
                                 uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "add_noise: bad args"));
    if (noise_type < 0 || noise_type > 2) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "add_noise: bad noise type");
    const float inv_scale = 1.0f / fmaxf(scale, 0.1f);
    const uint32_t oct = std::min(std::max(octaves, 1u), 8u);
    PFE_KERNEL(ctx, "add_noise", noise_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, amount * 255.0f / 100.0f, noise_type, monochrome ? 1 : 0,
        seed, inv_scale, oct));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_reduce_noise(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float strength, uint32_t radius,
                                    const uint8_t *mask, uint8_t *dst) {
    PFE_TRY(check(ctx, src, dst, w, h, "reduce_noise: bad args"));
    const int r = radius < 1 ? 1 : (int)std::min(radius, 1u << 20);
    const size_t smem = (size_t)(BIL_BX + 2 * r) * (BIL_BY + 2 * r) * 4;
    if (smem > 200 * 1024) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "reduce_noise: radius too large");
    const float sigma_s = (float)r, sigma_r = strength * 2.55f;
    if (smem > 48 * 1024) PFE_CUDA(ctx, cudaFuncSetAttribute(bilateral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PFE_KERNEL(ctx, "reduce_noise", bilateral_kernel<<<dim3(pfe_div_up(w, BIL_BX), pfe_div_up(h, BIL_BY)), BIL_BX * BIL_BY, smem, ctx->stream>>>(
        (const uint32_t *)src, mask, (uint32_t *)dst, (int)w, (int)h, r, 2.0f * sigma_s * sigma_s, 2.0f * sigma_r * sigma_r + 0.001f));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}
