// rgb_to_hsl / hue_to_rgb / hsl_to_rgb of src/ops/adjustments.rs:944-1012, shared by the adjustment
// kernels and the dodge / burn / sponge brush modes.
#pragma once
#include "common.cuh"

namespace {

// rgb_to_hsl / hue_to_rgb / hsl_to_rgb, adjustments.rs:944-1012
__device__ __forceinline__ void rgb_to_hsl(float r, float g, float b, float &H, float &S, float &L) {
    float mx = fmaxf(fmaxf(r, g), b), mn = fminf(fminf(r, g), b);
    float l = (mx + mn) / 2.0f;
    if (fabsf(mx - mn) < 1e-6f) { H = 0.f; S = 0.f; L = l; return; }
    float d = mx - mn;
    float s = l > 0.5f ? d / (2.0f - mx - mn) : d / (mx + mn);
    float hh;
    if (fabsf(mx - r) < 1e-6f) { hh = (g - b) / d; if (hh < 0.0f) hh += 6.0f; hh = hh / 6.0f; }
    else if (fabsf(mx - g) < 1e-6f) hh = ((b - r) / d + 2.0f) / 6.0f;
    else hh = ((r - g) / d + 4.0f) / 6.0f;
    H = hh; S = s; L = l;
}
__device__ __forceinline__ float hue_to_rgb(float p, float q, float t) {
    if (t < 0.0f) t += 1.0f;
    if (t > 1.0f) t -= 1.0f;
    if (t < 1.0f / 6.0f) return p + (q - p) * 6.0f * t;
    if (t < 1.0f / 2.0f) return q;
    if (t < 2.0f / 3.0f) return p + (q - p) * (2.0f / 3.0f - t) * 6.0f;
    return p;
}
__device__ __forceinline__ void hsl_to_rgb(float h, float s, float l, float eps, float &r, float &g, float &b) {
    if (fabsf(s) < eps) { r = g = b = l; return; }
    float q = l < 0.5f ? l * (1.0f + s) : l + s - l * s;
    float p = 2.0f * l - q;
    r = hue_to_rgb(p, q, h + 1.0f / 3.0f);
    g = hue_to_rgb(p, q, h);
    b = hue_to_rgb(p, q, h - 1.0f / 3.0f);
}

}  // namespace
