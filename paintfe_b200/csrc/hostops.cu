// Host-side pieces of the ABI: LUT construction (transcendental constants are computed on the
// host with libm exactly as the reference does, then shipped to the kernels) and TiledImage
// marshalling.  No device work in this file.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

inline uint8_t round_u8(float v) {  // `.round().clamp(0.0, 255.0) as u8`
    float r = roundf(v);
    if (r != r) return 0;
    return (uint8_t)(r < 0.0f ? 0.0f : (r > 255.0f ? 255.0f : r));
}
inline uint8_t trunc_u8(float v) {  // `.clamp(0.0, 255.0) as u8`
    if (v != v) return 0;
    return (uint8_t)(v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v));
}
inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

template <class F>
void parallel_for(size_t n, F f) {
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    if (n < 64 || nt == 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([=]() { for (size_t i = t; i < n; i += nt) f(i); });
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

// build_levels_lut, src/ops/adjustments.rs:424-446
void pfe_build_levels_lut(float in_black, float in_white, float gamma, float out_black, float out_white, uint8_t lut[256]) {
    const float in_range = fmaxf(in_white - in_black, 1.0f);
    const float out_range = out_white - out_black;
    const float inv_gamma = 1.0f / fmaxf(gamma, 0.01f);
    for (int i = 0; i < 256; i++) {
        float normalized = clamp01(((float)i - in_black) / in_range);
        float corrected = powf(normalized, inv_gamma);
        lut[i] = round_u8(out_black + corrected * out_range);
    }
}

// apply_levels LUT, src/ops/scripting.rs:1054-1066
void pfe_build_levels_lut_script(float in_black, float in_white, float gamma, uint8_t lut[256]) {
    const float in_range = fmaxf(in_white - in_black, 1.0f);
    const float inv_gamma = 1.0f / fmaxf(gamma, 0.01f);
    for (int i = 0; i < 256; i++) {
        float normalized = clamp01(((float)i - in_black) / in_range);
        lut[i] = trunc_u8(powf(normalized, inv_gamma) * 255.0f);
    }
}

// build_stretch_lut, adjustments.rs:232-253
void pfe_build_stretch_lut(uint8_t mn, uint8_t mx, uint8_t lut[256]) {
    if (mx <= mn) { for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i; return; }
    const float range = (float)(mx - mn);
    for (int i = 0; i < 256; i++) {
        float v = i <= mn ? 0.0f : (i >= mx ? 255.0f : ((float)i - (float)mn) / range * 255.0f);
        lut[i] = round_u8(v);
    }
}

// build_curves_lut (Fritsch-Carlson monotone cubic), adjustments.rs:634-729
void pfe_build_curves_lut(const float *p, int n, uint8_t lut[256]) {
    if (!p || n < 2) { for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i; return; }
    auto X = [&](int i) { return p[2 * i]; };
    auto Y = [&](int i) { return p[2 * i + 1]; };
    std::vector<float> delta((size_t)n - 1), m((size_t)n, 0.0f);
    for (int i = 0; i + 1 < n; i++) {
        float dx = X(i + 1) - X(i), dy = Y(i + 1) - Y(i);
        delta[i] = fabsf(dx) < 1e-6f ? 0.0f : dy / dx;
    }
    m[0] = delta[0];
    m[n - 1] = delta[n - 2];
    for (int i = 1; i + 1 < n; i++) m[i] = (delta[i - 1] * delta[i] <= 0.0f) ? 0.0f : (delta[i - 1] + delta[i]) / 2.0f;
    for (int i = 0; i + 1 < n; i++) {
        if (fabsf(delta[i]) < 1e-6f) { m[i] = 0.0f; m[i + 1] = 0.0f; continue; }
        float alpha = m[i] / delta[i], beta = m[i + 1] / delta[i];
        float s = alpha * alpha + beta * beta;
        if (s > 9.0f) {
            float tau = 3.0f / sqrtf(s);
            m[i] = tau * alpha * delta[i];
            m[i + 1] = tau * beta * delta[i];
        }
    }
    for (int i = 0; i < 256; i++) {
        const float x = (float)i;
        int seg = 0;
        for (int j = 0; j + 1 < n; j++) if (x >= X(j)) seg = j;
        if (x <= X(0)) { lut[i] = round_u8(Y(0)); continue; }
        if (x >= X(n - 1)) { lut[i] = round_u8(Y(n - 1)); continue; }
        const float x0 = X(seg), x1 = X(seg + 1), y0 = Y(seg), y1 = Y(seg + 1), hh = x1 - x0;
        if (fabsf(hh) < 1e-6f) { lut[i] = round_u8(y0); continue; }
        const float t = (x - x0) / hh, t2 = t * t, t3 = t2 * t;
        const float h00 = 2.0f * t3 - 3.0f * t2 + 1.0f, h10 = t3 - 2.0f * t2 + t;
        const float h01 = -2.0f * t3 + 3.0f * t2, h11 = t3 - t2;
        lut[i] = round_u8(h00 * y0 + h10 * hh * m[seg] + h01 * y1 + h11 * hh * m[seg + 1]);
    }
}

// build_multi_channel_luts, adjustments.rs:576-626: in = [RGB, R, G, B, A], out = [R, G, B, A]
void pfe_compose_curve_luts(const uint8_t in[5 * 256], uint8_t out[4 * 256]) {
    for (int i = 0; i < 256; i++) {
        const uint8_t m = in[i];
        out[i] = in[256 + m];
        out[256 + i] = in[512 + m];
        out[512 + i] = in[768 + m];
        out[768 + i] = in[1024 + i];
    }
}

// TiledImage::to_rgba_image, src/canvas/tiled_image.rs:271-293
int pfe_tiles_to_flat(const uint8_t *const *table, uint32_t w, uint32_t h, uint8_t *flat) {
    if (!table || !flat || !w || !h) return PFE_ERR_INVALID_ARG;
    const uint32_t C = PFE_CHUNK_SIZE, cxn = (w + C - 1) / C, cyn = (h + C - 1) / C;
    parallel_for((size_t)cxn * cyn, [=](size_t ci) {
        const uint32_t cx = (uint32_t)(ci % cxn), cy = (uint32_t)(ci / cxn);
        const uint32_t x0 = cx * C, y0 = cy * C, cw = std::min(C, w - x0), ch = std::min(C, h - y0);
        const uint8_t *tile = table[ci];
        for (uint32_t ly = 0; ly < ch; ly++) {
            uint8_t *d = flat + ((size_t)(y0 + ly) * w + x0) * 4;
            if (tile) memcpy(d, tile + (size_t)ly * C * 4, (size_t)cw * 4);
            else memset(d, 0, (size_t)cw * 4);
        }
    });
    return PFE_OK;
}

// TiledImage::from_rgba_image, tiled_image.rs:50-104
int pfe_flat_to_tiles(const uint8_t *flat, uint32_t w, uint32_t h, uint8_t *occupancy, uint8_t *tiles) {
    if (!flat || !occupancy || !w || !h) return PFE_ERR_INVALID_ARG;
    const uint32_t C = PFE_CHUNK_SIZE, cxn = (w + C - 1) / C, cyn = (h + C - 1) / C;
    parallel_for((size_t)cxn * cyn, [=](size_t ci) {
        const uint32_t cx = (uint32_t)(ci % cxn), cy = (uint32_t)(ci / cxn);
        const uint32_t x0 = cx * C, y0 = cy * C, cw = std::min(C, w - x0), ch = std::min(C, h - y0);
        bool has = false;
        for (uint32_t ly = 0; ly < ch && !has; ly++) {
            const uint8_t *s = flat + ((size_t)(y0 + ly) * w + x0) * 4;
            for (uint32_t lx = 0; lx < cw; lx++) if (s[lx * 4 + 3] != 0) { has = true; break; }
        }
        occupancy[ci] = has ? 1 : 0;
        if (tiles && has) {
            uint8_t *t = tiles + ci * (size_t)(C * C * 4);
            memset(t, 0, (size_t)C * C * 4);
            for (uint32_t ly = 0; ly < ch; ly++)
                memcpy(t + (size_t)ly * C * 4, flat + ((size_t)(y0 + ly) * w + x0) * 4, (size_t)cw * 4);
        }
    });
    return PFE_OK;
}

}  // extern "C"
