// Per-pixel adjustments: one kernel family, op selected by a warp-uniform switch.
// Reference: src/ops/adjustments.rs (rounding family, apply_pixel_transform[_from_flat] :21-108)
// and the inline Rhai bindings in src/ops/scripting.rs:869-1075 (truncating family).
// 8 algorithmic bytes per pixel; each thread handles 4 pixels (16-byte loads/stores).
#include <cstring>

#include "common.cuh"
#include "hsl.cuh"

namespace {

struct AdjParams {
    const uint32_t *src;
    uint32_t *dst;
    const uint8_t *mask;       // w*h or null (ignored by the scripting family)
    const uint8_t *occupancy;  // chunk bitmap or null
    const uint8_t *luts;       // device, 256 or 1024 bytes, or null
    float p[12];
    uint32_t w, h, chunks_x;
    int op;
    uint64_t n;
};

__device__ __forceinline__ uint32_t round_px(float r, float g, float b, float a) {  // adjustments.rs:33-39
    return pfe_pack(pfe_round_u8(r), pfe_round_u8(g), pfe_round_u8(b), pfe_round_u8(a));
}

__device__ __forceinline__ uint32_t adjust_px(int op, const float *p, const uint8_t *lut, uint32_t v) {
    const uint32_t r8 = v & 255u, g8 = (v >> 8) & 255u, b8 = (v >> 16) & 255u, a8 = v >> 24;
    // u8 -> f32 as 0x4B0000xx - 2^23 (ALU + FMA pipes) instead of the quarter-rate I2F
    const float r = pfe_u8_to_f32(r8), g = pfe_u8_to_f32(g8), b = pfe_u8_to_f32(b8), a = pfe_u8_to_f32(a8);
    switch (op) {
    case PFE_ADJ_INVERT: return round_px(255.0f - r, 255.0f - g, 255.0f - b, a);
    case PFE_ADJ_INVERT_ALPHA: return round_px(r, g, b, 255.0f - a);
    case PFE_ADJ_SEPIA:
        return round_px(fminf(0.393f * r + 0.769f * g + 0.189f * b, 255.0f),
                        fminf(0.349f * r + 0.686f * g + 0.168f * b, 255.0f),
                        fminf(0.272f * r + 0.534f * g + 0.131f * b, 255.0f), a);
    case PFE_ADJ_DESATURATE: {
        uint32_t lum = pfe_round_u8(0.2126f * r + 0.7152f * g + 0.0722f * b);
        return pfe_pack(lum, lum, lum, a8);
    }
    case PFE_ADJ_BRIGHTNESS_CONTRAST: {
        const float factor = (259.0f * (p[1] + 255.0f)) / (255.0f * (259.0f - p[1]));
        return round_px(factor * (r + p[0] - 128.0f) + 128.0f, factor * (g + p[0] - 128.0f) + 128.0f,
                        factor * (b + p[0] - 128.0f) + 128.0f, a);
    }
    case PFE_ADJ_HSL: {
        const float sat_factor = 1.0f + p[1] / 100.0f;
        const float light_offset = p[2] * 255.0f / 100.0f;
        float hh, s, l;
        rgb_to_hsl(pfe_div255(r), pfe_div255(g), pfe_div255(b), hh, s, l);
        float t = hh + p[0] / 360.0f;
        float nh = t - truncf(t);  // f32::fract
        if (nh < 0.0f) nh = nh + 1.0f;
        float ns = pfe_clampf(s * sat_factor, 0.0f, 1.0f);
        float rr, gg, bb;
        hsl_to_rgb(nh, ns, l, 1e-6f, rr, gg, bb);
        return round_px(rr * 255.0f + light_offset, gg * 255.0f + light_offset, bb * 255.0f + light_offset, a);
    }
    case PFE_ADJ_EXPOSURE: return round_px(r * p[0], g * p[0], b * p[0], a);
    case PFE_ADJ_LUT_RGB: return pfe_pack(lut[r8], lut[g8], lut[b8], a8);
    case PFE_ADJ_LUT_RGBA: return pfe_pack(lut[r8], lut[256 + g8], lut[512 + b8], lut[768 + a8]);
    case PFE_ADJ_TEMPERATURE_TINT: {
        const float temp_shift = p[0] * 1.5f, tint_shift = p[1] * 1.0f;
        return round_px(r + temp_shift, g - tint_shift * 0.5f, b - temp_shift, a);
    }
    case PFE_ADJ_HIGHLIGHTS_SHADOWS: {
        const float shadow_amt = p[0] / 100.0f, highlight_amt = p[1] / 100.0f;
        float lum = (0.2126f * r + 0.7152f * g + 0.0722f * b) / 255.0f;
        float sw = (1.0f - lum) * (1.0f - lum), hw = lum * lum;
        float adj = sw * shadow_amt * 128.0f + hw * highlight_amt * 128.0f;
        return round_px(r + adj, g + adj, b + adj, a);
    }
    case PFE_ADJ_THRESHOLD: {
        const float v1 = (0.2126f * r + 0.7152f * g + 0.0722f * b) >= p[0] ? 255.0f : 0.0f;
        return round_px(v1, v1, v1, a);
    }
    case PFE_ADJ_POSTERIZE: {
        const float f1 = p[0] - 1.0f;
        return round_px(roundf(r / 255.0f * f1) / f1 * 255.0f, roundf(g / 255.0f * f1) / f1 * 255.0f,
                        roundf(b / 255.0f * f1) / f1 * 255.0f, a);
    }
    case PFE_ADJ_COLOR_BALANCE: {  // color_balance_pixel, adjustments.rs:1321-1338
        const float lum = (0.2126f * r + 0.7152f * g + 0.0722f * b) / 255.0f;
        const float s0 = fmaxf(1.0f - lum * 2.0f, 0.0f), h0 = fmaxf(lum * 2.0f - 1.0f, 0.0f);
        const float sw = s0 * s0, hw = h0 * h0;
        const float mw = fmaxf(1.0f - sw - hw, 0.0f);
        return round_px(r + (sw * p[0] + mw * p[3] + hw * p[6]) * 1.28f, g + (sw * p[1] + mw * p[4] + hw * p[7]) * 1.28f,
                        b + (sw * p[2] + mw * p[5] + hw * p[8]) * 1.28f, a);
    }
    case PFE_ADJ_GRADIENT_MAP: {
        const uint32_t lum = min((uint32_t)__float2int_rz(0.2126f * r + 0.7152f * g + 0.0722f * b), 255u);
        return pfe_pack(lut[lum * 4], lut[lum * 4 + 1], lut[lum * 4 + 2], a8);
    }
    case PFE_ADJ_BLACK_AND_WHITE: {
        const float v1 = pfe_clampf((r * p[0] + g * p[1] + b * p[2]) / 100.0f, 0.0f, 255.0f);
        return round_px(v1, v1, v1, a);
    }
    case PFE_ADJ_VIBRANCE: {  // vibrance_pixel, adjustments.rs:1431-1444
        float hh, sat, l;
        rgb_to_hsl(pfe_div255(r), pfe_div255(g), pfe_div255(b), hh, sat, l);
        const float boost = p[0] >= 0.0f ? p[0] * ((1.0f - sat) * (1.0f - sat)) : p[0] * (sat * sat);
        const float ns = pfe_clampf(sat + boost, 0.0f, 1.0f);
        float rr, gg, bb;
        hsl_to_rgb(hh, ns, l, 1e-6f, rr, gg, bb);
        return round_px(rr * 255.0f, gg * 255.0f, bb * 255.0f, a);
    }
    // ---- scripting.rs inline variants ----
    case PFE_ADJ_S_INVERT: return pfe_pack(255u - r8, 255u - g8, 255u - b8, a8);
    case PFE_ADJ_S_DESATURATE: {
        uint32_t gray = (r8 * 299u + g8 * 587u + b8 * 114u) / 1000u;
        return pfe_pack(gray, gray, gray, a8);
    }
    case PFE_ADJ_S_SEPIA:
        return pfe_pack(pfe_as_u8(fminf(r * 0.393f + g * 0.769f + b * 0.189f, 255.0f)),
                        pfe_as_u8(fminf(r * 0.349f + g * 0.686f + b * 0.168f, 255.0f)),
                        pfe_as_u8(fminf(r * 0.272f + g * 0.534f + b * 0.131f, 255.0f)), a8);
    case PFE_ADJ_S_SEPIA_STRENGTH: {
        const float st = p[0], inv = 1.0f - st;
        float sr = fminf(r * 0.393f + g * 0.769f + b * 0.189f, 255.0f);
        float sg = fminf(r * 0.349f + g * 0.686f + b * 0.168f, 255.0f);
        float sb = fminf(r * 0.272f + g * 0.534f + b * 0.131f, 255.0f);
        return pfe_pack(pfe_as_u8(r * inv + sr * st), pfe_as_u8(g * inv + sg * st), pfe_as_u8(b * inv + sb * st), a8);
    }
    case PFE_ADJ_S_BRIGHTNESS_CONTRAST: {
        const float factor = (259.0f * (p[1] + 255.0f)) / (255.0f * (259.0f - p[1]));
        return pfe_pack(pfe_as_u8(factor * (r + p[0] - 128.0f) + 128.0f), pfe_as_u8(factor * (g + p[0] - 128.0f) + 128.0f),
                        pfe_as_u8(factor * (b + p[0] - 128.0f) + 128.0f), a8);
    }
    case PFE_ADJ_S_HSL: {
        const float sat_factor = 1.0f + p[1] / 100.0f;
        const float light_offset = p[2] * 255.0f / 100.0f;
        float fr = pfe_div255(r), fg = pfe_div255(g), fb = pfe_div255(b);
        float cmax = fmaxf(fmaxf(fr, fg), fb), cmin = fminf(fminf(fr, fg), fb);
        float l = (cmax + cmin) / 2.0f;
        float hh = 0.0f, s = 0.0f;
        if (!(fabsf(cmax - cmin) < 1e-10f)) {  // same expressions as scripting.rs:973-990, selected instead of branched
            float d = cmax - cmin;
            s = pfe_fast_div(d, l > 0.5f ? 2.0f - cmax - cmin : cmax + cmin);
            const bool is_r = fabsf(cmax - fr) < 1e-10f, is_g = fabsf(cmax - fg) < 1e-10f;
            const float num = is_r ? fg - fb : (is_g ? fb - fr : fr - fg);
            const float off = is_r ? (fg < fb ? 6.0f : 0.0f) : (is_g ? 2.0f : 4.0f);
            hh = pfe_fast_div(pfe_fast_div(num, d) + off, 6.0f);
        }
        float t = hh + p[0] / 360.0f;
        float nh = fmodf(t, 1.0f);  // rem_euclid(1.0)
        if (nh < 0.0f) nh = nh + 1.0f;
        float ns = pfe_clampf(s * sat_factor, 0.0f, 1.0f);
        float rr, gg, bb;
        hsl_to_rgb(nh, ns, l, 1e-10f, rr, gg, bb);
        return pfe_pack(pfe_as_u8(rr * 255.0f + light_offset), pfe_as_u8(gg * 255.0f + light_offset),
                        pfe_as_u8(bb * 255.0f + light_offset), a8);
    }
    case PFE_ADJ_S_EXPOSURE: return pfe_pack(pfe_as_u8(r * p[0]), pfe_as_u8(g * p[0]), pfe_as_u8(b * p[0]), a8);
    case PFE_ADJ_S_LUT_RGB: return pfe_pack(lut[r8], lut[g8], lut[b8], a8);
    default: return v;
    }
}

// One instantiation per op: the switch in adjust_px folds away, so each op gets the registers and code
// size of its own arithmetic only (the HSL family is ~100 instructions per pixel, invert is 4).
template <int VEC, int OP>
__global__ void __launch_bounds__(256) adjust_kernel(const __grid_constant__ AdjParams P) {
    __shared__ uint8_t lut[1024];
    if (P.luts) {
        const int nl = (OP == PFE_ADJ_LUT_RGBA || OP == PFE_ADJ_GRADIENT_MAP) ? 1024 : 256;
        for (int i = threadIdx.x; i < nl; i += blockDim.x) lut[i] = P.luts[i];
        __syncthreads();
    }
    const bool use_mask = P.mask != nullptr && OP < 32;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g * VEC < P.n; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t px = g * VEC;
        uint32_t v[VEC];
        if constexpr (VEC == 4) {
            uint4 q = *reinterpret_cast<const uint4 *>(P.src + px);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
            v[0] = P.src[px];
        }
#pragma unroll
        for (int k = 0; k < VEC; k++) {
            const uint64_t i = px + k;
            bool skip = use_mask && P.mask[i] == 0;
            if (P.occupancy) {
                uint32_t y = (uint32_t)(i / P.w), x = (uint32_t)(i - (uint64_t)y * P.w);
                skip = skip || P.occupancy[(size_t)(y / PFE_CHUNK_SIZE) * P.chunks_x + x / PFE_CHUNK_SIZE] == 0;
            }
            if (!skip) v[k] = adjust_px(OP, P.p, lut, v[k]);
        }
        if constexpr (VEC == 4) *reinterpret_cast<uint4 *>(P.dst + px) = make_uint4(v[0], v[1], v[2], v[3]);
        else P.dst[px] = v[0];
    }
}

// auto_levels min/max scan, adjustments.rs:167-196: out = {min_r,max_r,min_g,max_g,min_b,max_b}
__global__ void minmax_kernel(const uint32_t *src, const uint8_t *mask, size_t n, uint32_t *out) {
    uint32_t mn = 0x00FFFFFFu, mx = 0u;  // packed per-byte r,g,b
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (mask && mask[i] == 0) continue;
        uint32_t v = src[i];
        if ((v >> 24) == 0) continue;
        mn = __vminu4(mn, v & 0x00FFFFFFu);
        mx = __vmaxu4(mx, v & 0x00FFFFFFu);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = __vminu4(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = __vmaxu4(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        for (int c = 0; c < 3; c++) {
            atomicMin(&out[c * 2], (mn >> (8 * c)) & 255u);
            atomicMax(&out[c * 2 + 1], (mx >> (8 * c)) & 255u);
        }
    }
}

}  // namespace

extern "C" int pfe_dev_adjust(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const pfe_adjust_desc *d,
                              const uint8_t *mask, const uint8_t *occupancy, uint8_t *dst) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !dst || !d || !w || !h) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "adjust: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool lut_op = d->op == PFE_ADJ_LUT_RGB || d->op == PFE_ADJ_LUT_RGBA || d->op == PFE_ADJ_S_LUT_RGB || d->op == PFE_ADJ_GRADIENT_MAP;
    const bool known = (d->op >= 0 && d->op <= PFE_ADJ_VIBRANCE) || (d->op >= 32 && d->op <= PFE_ADJ_S_LUT_RGB);
    if (!known) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "adjust: unknown op");
    if (lut_op && !d->luts) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "adjust: LUT op without luts");
    AdjParams P;
    memset(&P, 0, sizeof(P));
    P.src = (const uint32_t *)src; P.dst = (uint32_t *)dst; P.mask = mask; P.occupancy = occupancy;
    memcpy(P.p, d->params, sizeof(P.p));
    P.w = w; P.h = h; P.chunks_x = pfe_div_up(w, PFE_CHUNK_SIZE); P.op = d->op; P.n = (uint64_t)w * h;
    if (lut_op) {
        void *ld;
        PFE_TRY(pfe_small_upload(ctx, d->luts, (d->op == PFE_ADJ_LUT_RGBA || d->op == PFE_ADJ_GRADIENT_MAP) ? 1024 : 256, &ld));
        P.luts = (const uint8_t *)ld;
    }
    const bool vec = (P.n % 4 == 0) && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
    const uint64_t groups = vec ? P.n / 4 : P.n;
    unsigned blocks = (unsigned)std::min<uint64_t>((groups + 255) / 256, (uint64_t)ctx->sm_count * 16);
#define PFE_ADJ_CASE(OPV)                                                                              \
    case OPV:                                                                                          \
        if (vec) PFE_KERNEL(ctx, "adjust", adjust_kernel<4, OPV><<<blocks, 256, 0, ctx->stream>>>(P)); \
        else PFE_KERNEL(ctx, "adjust", adjust_kernel<1, OPV><<<blocks, 256, 0, ctx->stream>>>(P));     \
        break;
    switch (d->op) {
        PFE_ADJ_CASE(PFE_ADJ_INVERT) PFE_ADJ_CASE(PFE_ADJ_INVERT_ALPHA) PFE_ADJ_CASE(PFE_ADJ_SEPIA) PFE_ADJ_CASE(PFE_ADJ_DESATURATE)
        PFE_ADJ_CASE(PFE_ADJ_BRIGHTNESS_CONTRAST) PFE_ADJ_CASE(PFE_ADJ_HSL) PFE_ADJ_CASE(PFE_ADJ_EXPOSURE) PFE_ADJ_CASE(PFE_ADJ_LUT_RGB)
        PFE_ADJ_CASE(PFE_ADJ_LUT_RGBA) PFE_ADJ_CASE(PFE_ADJ_TEMPERATURE_TINT) PFE_ADJ_CASE(PFE_ADJ_HIGHLIGHTS_SHADOWS)
        PFE_ADJ_CASE(PFE_ADJ_THRESHOLD) PFE_ADJ_CASE(PFE_ADJ_POSTERIZE) PFE_ADJ_CASE(PFE_ADJ_COLOR_BALANCE) PFE_ADJ_CASE(PFE_ADJ_GRADIENT_MAP)
        PFE_ADJ_CASE(PFE_ADJ_BLACK_AND_WHITE) PFE_ADJ_CASE(PFE_ADJ_VIBRANCE) PFE_ADJ_CASE(PFE_ADJ_S_INVERT) PFE_ADJ_CASE(PFE_ADJ_S_DESATURATE)
        PFE_ADJ_CASE(PFE_ADJ_S_SEPIA) PFE_ADJ_CASE(PFE_ADJ_S_SEPIA_STRENGTH) PFE_ADJ_CASE(PFE_ADJ_S_BRIGHTNESS_CONTRAST)
        PFE_ADJ_CASE(PFE_ADJ_S_HSL) PFE_ADJ_CASE(PFE_ADJ_S_EXPOSURE) PFE_ADJ_CASE(PFE_ADJ_S_LUT_RGB)
        default: return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "adjust: unknown op");
    }
#undef PFE_ADJ_CASE
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_channel_minmax(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const uint8_t *mask,
                                      uint8_t out[6]) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !out || !w || !h) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "channel_minmax: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t init[6] = {255u, 0u, 255u, 0u, 255u, 0u};  // adjustments.rs:160-165
    void *od;
    PFE_TRY(pfe_small_upload(ctx, init, sizeof(init), &od));
    PFE_KERNEL(ctx, "minmax", minmax_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>((const uint32_t *)src, mask, (size_t)w * h, (uint32_t *)od));
    PFE_LAUNCHED(ctx);
    uint32_t res[6];
    PFE_CUDA(ctx, cudaMemcpyAsync(res, od, sizeof(res), cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 6; i++) out[i] = (uint8_t)res[i];
    return PFE_OK;
}
