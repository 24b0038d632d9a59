// rgb_to_hsl / hue_to_rgb / hsl_to_rgb of src/ops/adjustments.rs:944-1012, shared by the adjustment
// kernels and the dodge / burn / sponge brush modes.
//
// Same operations in the same order as the reference, arranged for the GPU: every division goes through
// pfe_fast_div (the compiler's own correctly rounded sequence without its range check and slow-path call - the
// operands here are u8 / 255 derived, far from the exponent extremes), and the three-way / four-way branches are
// evaluated on both sides and selected, because neighbouring pixels take different sides and a divergent branch costs
// more than the few extra operations.  Results are identical to the branching form: each candidate is the reference's
// own expression, and the one the reference would have returned is the one selected.
#pragma once
#include "common.cuh"

namespace {

__device__ __forceinline__ void rgb_to_hsl(float r, float g, float b, float &H, float &S, float &L) {
    const float mx = fmaxf(fmaxf(r, g), b), mn = fminf(fminf(r, g), b);
    const float l = (mx + mn) / 2.0f;
    const float d = mx - mn;
    if (fabsf(d) < 1e-6f) { H = 0.f; S = 0.f; L = l; return; }
    const float s = pfe_fast_div(d, l > 0.5f ? 2.0f - mx - mn : mx + mn);
    // which channel is the maximum decides numerator and offset: (g-b)/d (+6 if negative), (b-r)/d + 2, (r-g)/d + 4
    const bool is_r = fabsf(mx - r) < 1e-6f, is_g = fabsf(mx - g) < 1e-6f;
    const float num = is_r ? g - b : (is_g ? b - r : r - g);
    const float q = pfe_fast_div(num, d);
    const float off = is_r ? (q < 0.0f ? 6.0f : 0.0f) : (is_g ? 2.0f : 4.0f);
    // the reference adds nothing in the first case when q >= 0: q + 0.0f == q bit for bit except for q == -0, which
    // cannot occur here (q < 0 takes the +6 side and g - b == 0 yields +0)
    H = pfe_fast_div(q + off, 6.0f);
    S = s;
    L = l;
}
__device__ __forceinline__ float hue_to_rgb(float p, float q, float t) {
    t = t < 0.0f ? t + 1.0f : t;
    t = t > 1.0f ? t - 1.0f : t;
    const float qp = q - p;
    const float rise = p + qp * 6.0f * t;                        // t < 1/6
    const float fall = p + qp * (2.0f / 3.0f - t) * 6.0f;        // 1/2 <= t < 2/3
    return t < 1.0f / 6.0f ? rise : (t < 1.0f / 2.0f ? q : (t < 2.0f / 3.0f ? fall : p));
}
__device__ __forceinline__ void hsl_to_rgb(float h, float s, float l, float eps, float &r, float &g, float &b) {
    if (fabsf(s) < eps) { r = g = b = l; return; }
    const float q = l < 0.5f ? l * (1.0f + s) : l + s - l * s;
    const float p = 2.0f * l - q;
    r = hue_to_rgb(p, q, h + 1.0f / 3.0f);
    g = hue_to_rgb(p, q, h);
    b = hue_to_rgb(p, q, h - 1.0f / 3.0f);
}

}  // namespace
