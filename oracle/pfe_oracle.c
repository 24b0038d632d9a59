/*
 * pfe_oracle.c — CPU restatement of PaintFE's compositor / filter / adjustment / warp /
 * brush-stamp hot path (reference: kylejckson/PaintFE v1.3.9 @ 16410bd).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it.  The product path
 * (paintfe_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below
 * against the reference's own golden PNGs (copied as data into tests/golden/ref/).
 * The reference itself (Rust, ~400 crates) cannot be built in this image, so there is
 * no oracle/_ref binary; the goldens are the anchor.
 *
 * Rules that make this bit-exact with the Rust code:
 *   - every arithmetic op is a separately rounded IEEE f32 op, evaluated left to right
 *     exactly as written in the reference (compile with -ffp-contract=off, no fast-math);
 *   - `x as u8`  == truncate toward zero, saturate to [0,255], NaN -> 0        (as_u8)
 *   - `x.round()`== roundf (half away from zero)
 *   - `x.clamp(a,b)`, `.min()`, `.max()` as Rust defines them for non-NaN input.
 *
 * All citations are file:line under /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PFE_CHUNK 64 /* src/canvas/defs.rs:7 */

/* ------------------------------------------------------------------------- */
/* Rust cast / math helpers                                                   */
/* ------------------------------------------------------------------------- */
static inline uint8_t as_u8(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
static inline int32_t as_i32(float v) {
    if (!(v == v)) return 0;
    if (v <= -2147483648.0f) return INT32_MIN;
    if (v >= 2147483648.0f) return INT32_MAX;
    return (int32_t)v;
}
static inline uint32_t as_u32(float v) {
    if (!(v == v)) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 4294967296.0f) return UINT32_MAX;
    return (uint32_t)v;
}
static inline float clampf(float v, float lo, float hi) {
    if (v < lo) return lo;
    if (v > hi) return hi;
    return v;
}
static inline float minf(float a, float b) { return fminf(a, b); }
static inline float maxf(float a, float b) { return fmaxf(a, b); }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline uint8_t round_u8(float v) { return as_u8(clampf(roundf(v), 0.0f, 255.0f)); }

/* ------------------------------------------------------------------------- */
/* Blend: src/canvas/canvas_state.rs:1246-1505                                */
/* ------------------------------------------------------------------------- */
static inline float overlay_ch(float base, float top) { /* :1425 */
    if (base < 0.5f) return 2.0f * base * top;
    return 1.0f - 2.0f * (1.0f - base) * (1.0f - top);
}
static inline float color_burn_ch(float base, float top) { /* :1433 */
    if (top == 0.0f) return 0.0f;
    return maxf(1.0f - (1.0f - base) / top, 0.0f);
}
static inline float color_dodge_ch(float base, float top) { /* :1441 */
    if (top >= 1.0f) return 1.0f;
    return minf(base / (1.0f - top), 1.0f);
}
static inline float reflect_ch(float base, float top) { /* :1449 */
    if (top >= 1.0f) return 1.0f;
    return minf(base * base / (1.0f - top), 1.0f);
}
static inline float soft_light_ch(float base, float top) { /* :1458 */
    if (top <= 0.5f) return base - (1.0f - 2.0f * top) * base * (1.0f - base);
    float d;
    if (base <= 0.25f) d = ((16.0f * base - 12.0f) * base + 4.0f) * base;
    else d = sqrtf(base);
    return base + (2.0f * top - 1.0f) * (d - base);
}
static inline float divide_ch(float base, float top) { /* :1471 */
    if (top <= 0.0f) return 1.0f;
    return minf(base / top, 1.0f);
}
static inline float vivid_light_ch(float base, float top) { /* :1479 */
    if (top <= 0.5f) {
        float t2 = 2.0f * top;
        if (t2 <= 0.0f) return 0.0f;
        return maxf(1.0f - (1.0f - base) / t2, 0.0f);
    } else {
        float t2 = 2.0f * (top - 0.5f);
        if (t2 >= 1.0f) return 1.0f;
        return minf(base / (1.0f - t2), 1.0f);
    }
}
static inline float pin_light_ch(float base, float top) { /* :1499 */
    if (top <= 0.5f) return minf(base, 2.0f * top);
    return maxf(base, 2.0f * (top - 0.5f));
}

static inline float blend_ch(int mode, float b, float t) { /* :1304-1405 */
    switch (mode) {
    case 0: return t;                                          /* Normal */
    case 1: return b * t;                                      /* Multiply */
    case 2: return 1.0f - (1.0f - b) * (1.0f - t);             /* Screen */
    case 3: return minf(b + t, 1.0f);                          /* Additive */
    case 4: return reflect_ch(b, t);                           /* Reflect */
    case 5: return reflect_ch(t, b);                           /* Glow */
    case 6: return color_burn_ch(b, t);                        /* ColorBurn */
    case 7: return color_dodge_ch(b, t);                       /* ColorDodge */
    case 8: return overlay_ch(b, t);                           /* Overlay */
    case 9: return fabsf(b - t);                               /* Difference */
    case 10: return 1.0f - fabsf(1.0f - b - t);                /* Negation */
    case 11: return maxf(b, t);                                /* Lighten */
    case 12: return minf(b, t);                                /* Darken */
    case 15: return overlay_ch(t, b);                          /* HardLight */
    case 16: return soft_light_ch(b, t);                       /* SoftLight */
    case 17: return b + t - 2.0f * b * t;                      /* Exclusion */
    case 18: return maxf(b - t, 0.0f);                         /* Subtract */
    case 19: return divide_ch(b, t);                           /* Divide */
    case 20: return maxf(b + t - 1.0f, 0.0f);                  /* LinearBurn */
    case 21: return vivid_light_ch(b, t);                      /* VividLight */
    case 22: return clampf(b + 2.0f * t - 1.0f, 0.0f, 1.0f);   /* LinearLight */
    case 23: return pin_light_ch(b, t);                        /* PinLight */
    case 24: return (b + t >= 1.0f) ? 1.0f : 0.0f;             /* HardMix */
    default: return t; /* BlendMode::from_u8 maps unknown ids to Normal, layers.rs:183 */
    }
}

/* blend_pixel_static, canvas_state.rs:1246-1422.  px are RGBA8 byte quads. */
void pfo_blend_pixel(const uint8_t base[4], const uint8_t top[4], int mode, float opacity,
                     uint8_t out[4]) {
    if (mode < 0 || mode > 24) mode = 0;
    if (top[3] == 0) { memcpy(out, base, 4); return; }                         /* :1253 */
    if (mode == 0 && opacity >= 1.0f && top[3] == 255) { memcpy(out, top, 4); return; } /* :1258 */
    opacity = clampf(opacity, 0.0f, 1.0f);
    float br = base[0] / 255.0f, bg = base[1] / 255.0f, bb = base[2] / 255.0f, ba = base[3] / 255.0f;
    float tr = top[0] / 255.0f, tg = top[1] / 255.0f, tb = top[2] / 255.0f;
    float ta = (top[3] / 255.0f) * opacity;
    if (mode == 14) { /* Overwrite :1275 */
        out[0] = as_u8(tr * 255.0f); out[1] = as_u8(tg * 255.0f);
        out[2] = as_u8(tb * 255.0f); out[3] = as_u8(ta * 255.0f);
        return;
    }
    if (mode == 13) { /* Xor :1283 */
        float xa = ba * (1.0f - ta) + ta * (1.0f - ba);
        if (xa == 0.0f) { out[0] = out[1] = out[2] = out[3] = 0; return; }
        float xr = (br * ba * (1.0f - ta) + tr * ta * (1.0f - ba)) / xa;
        float xg = (bg * ba * (1.0f - ta) + tg * ta * (1.0f - ba)) / xa;
        float xb = (bb * ba * (1.0f - ta) + tb * ta * (1.0f - ba)) / xa;
        out[0] = as_u8(clampf(xr * 255.0f, 0.0f, 255.0f));
        out[1] = as_u8(clampf(xg * 255.0f, 0.0f, 255.0f));
        out[2] = as_u8(clampf(xb * 255.0f, 0.0f, 255.0f));
        out[3] = as_u8(clampf(xa * 255.0f, 0.0f, 255.0f));
        return;
    }
    float r = blend_ch(mode, br, tr), g = blend_ch(mode, bg, tg), b = blend_ch(mode, bb, tb);
    float oa = ta + ba * (1.0f - ta);                                          /* :1407 */
    if (oa == 0.0f) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    float orr = (r * ta + br * ba * (1.0f - ta)) / oa;
    float og = (g * ta + bg * ba * (1.0f - ta)) / oa;
    float ob = (b * ta + bb * ba * (1.0f - ta)) / oa;
    out[0] = as_u8(clampf(orr * 255.0f, 0.0f, 255.0f));
    out[1] = as_u8(clampf(og * 255.0f, 0.0f, 255.0f));
    out[2] = as_u8(clampf(ob * 255.0f, 0.0f, 255.0f));
    out[3] = as_u8(clampf(oa * 255.0f, 0.0f, 255.0f));
}

/* Layer descriptor shared (by layout) with include/pfe_b200.h pfe_layer_desc. */
typedef struct pfo_layer_desc {
    const uint8_t *rgba;  /* w*h*4 straight RGBA8; ignored for adjustment layers */
    const uint8_t *mask;  /* w*h conceal plane (alpha of Layer::mask) or NULL, layers.rs:395-399 */
    float opacity;
    uint8_t blend;        /* BlendMode::to_u8, layers.rs:125-153 */
    uint8_t visible;
    uint8_t kind;         /* 0 raster, 1 Exposure, 2 BrightnessContrast, 3 Invert, 4 ChannelMixer */
    uint8_t _pad;
    float adj[16];        /* 1: [gain=2^ev]  2: [brightness, contrast]  4: red[4] green[4] blue[4] alpha[4] */
} pfo_layer_desc;

/* AdjustmentLayerData::apply_to_pixel_with_opacity, layers.rs:276-325 */
static void adj_apply(const pfo_layer_desc *L, const uint8_t p[4], uint8_t out[4]) {
    uint8_t a[4] = {p[0], p[1], p[2], p[3]};
    switch (L->kind) {
    case 1: { /* Exposure :279 — gain = 2^ev computed by the caller with powf */
        float gain = L->adj[0];
        for (int c = 0; c < 3; c++) a[c] = as_u8(clampf((float)p[c] * gain, 0.0f, 255.0f));
    } break;
    case 2: { /* BrightnessContrast :288 */
        float brightness = L->adj[0], contrast = L->adj[1];
        float factor = (259.0f * (contrast + 255.0f)) / (255.0f * (259.0f - contrast));
        for (int c = 0; c < 3; c++)
            a[c] = as_u8(clampf(factor * ((float)p[c] + brightness - 128.0f) + 128.0f, 0.0f, 255.0f));
    } break;
    case 3: /* Invert :298 */
        a[0] = 255 - p[0]; a[1] = 255 - p[1]; a[2] = 255 - p[2];
        break;
    case 4: { /* ChannelMixer :299 */
        float s[4] = {(float)p[0], (float)p[1], (float)p[2], (float)p[3]};
        for (int c = 0; c < 4; c++) {
            const float *m = &L->adj[c * 4];
            a[c] = as_u8(clampf(s[0] * m[0] + s[1] * m[1] + s[2] * m[2] + s[3] * m[3], 0.0f, 255.0f));
        }
    } break;
    default: break;
    }
    float t = clampf(L->opacity, 0.0f, 1.0f);                                   /* :316 */
    float inv = 1.0f - t;
    for (int c = 0; c < 4; c++) out[c] = as_u8(roundf((float)p[c] * inv + (float)a[c] * t));
}

/* One pixel of composite_viewport's layer loop, canvas_state.rs:575-677 (no preview layer:
 * the CLI / export path never has one). */
static inline void flatten_pixel(const pfo_layer_desc *layers, uint32_t n, size_t pi,
                                 uint8_t acc[4]) {
    acc[0] = acc[1] = acc[2] = acc[3] = 0;                                      /* :573 */
    for (uint32_t li = 0; li < n; li++) {
        const pfo_layer_desc *L = &layers[li];
        if (!L->visible) continue;                                              /* :576 */
        if (L->kind != 0) { uint8_t o[4]; adj_apply(L, acc, o); memcpy(acc, o, 4); continue; } /* :579 */
        uint8_t top[4];
        memcpy(top, L->rgba + pi * 4, 4);
        if (L->mask) {                                                          /* :660 */
            uint32_t conceal = L->mask[pi];
            if (conceal > 0) top[3] = (uint8_t)(((uint32_t)top[3] * (255u - conceal)) / 255u);
        }
        int opaque_overwrite = (L->blend == 0 && L->opacity >= 1.0f);           /* :605 */
        if (opaque_overwrite && top[3] == 255) memcpy(acc, top, 4);             /* :667 */
        else { uint8_t o[4]; pfo_blend_pixel(acc, top, L->blend, L->opacity, o); memcpy(acc, o, 4); }
    }
}

/* CanvasState::composite, canvas_state.rs:482-698.  `active` is the active-chunk bitmap
 * (ceil(w/64) x ceil(h/64) bytes, :529-550) or NULL meaning every chunk is active.
 * Inactive chunks stay transparent zero (:506). Parallel over chunks like rayon (:565). */
void pfo_flatten(const pfo_layer_desc *layers, uint32_t n, uint32_t w, uint32_t h,
                 const uint8_t *active, uint8_t *dst) {
    uint32_t cxn = (w + PFE_CHUNK - 1) / PFE_CHUNK, cyn = (h + PFE_CHUNK - 1) / PFE_CHUNK;
    long total = (long)cxn * cyn;
#pragma omp parallel for schedule(dynamic, 4)
    for (long ci = 0; ci < total; ci++) {
        uint32_t cx = (uint32_t)(ci % cxn), cy = (uint32_t)(ci / cxn);
        uint32_t x0 = cx * PFE_CHUNK, y0 = cy * PFE_CHUNK;
        uint32_t x1 = x0 + PFE_CHUNK < w ? x0 + PFE_CHUNK : w;
        uint32_t y1 = y0 + PFE_CHUNK < h ? y0 + PFE_CHUNK : h;
        int on = active ? active[ci] != 0 : 1;
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = x0; x < x1; x++) {
                size_t pi = (size_t)y * w + x;
                if (!on) { memset(dst + pi * 4, 0, 4); continue; }
                flatten_pixel(layers, n, pi, dst + pi * 4);
            }
    }
}

/* ------------------------------------------------------------------------- */
/* Gaussian: src/ops/filters.rs:141-316                                       */
/* ------------------------------------------------------------------------- */
/* build_gaussian_kernel :214-234. Returns radius; writes 2r+1 weights (caller sizes k). */
int pfo_gaussian_radius(float sigma) { return (int)as_u32(ceilf(sigma * 3.0f)); }
int pfo_build_gaussian_kernel(float sigma, float *k) {
    int radius = pfo_gaussian_radius(sigma);
    if (radius == 0) { k[0] = 1.0f; return 0; }
    int len = radius * 2 + 1;
    float s2 = 2.0f * sigma * sigma;
    float sum = 0.0f;
    for (int i = 0; i < len; i++) {
        float x = (float)i - (float)radius;
        float v = expf(-x * x / s2);
        k[i] = v;
        sum += v;
    }
    float inv = 1.0f / sum;
    for (int i = 0; i < len; i++) k[i] *= inv;
    return radius;
}

/* parallel_gaussian_blur :242-316 */
void pfo_gaussian_blur(const uint8_t *src, uint32_t w, uint32_t h, float sigma, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    int radius = pfo_gaussian_radius(sigma);
    float *k = (float *)malloc(sizeof(float) * (size_t)(2 * radius + 1));
    pfo_build_gaussian_kernel(sigma, k);
    int len = 2 * radius + 1;
    size_t n4 = (size_t)w * h * 4;
    float *buf = (float *)malloc(n4 * sizeof(float));
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++) {
        const uint8_t *row = src + (size_t)y * w * 4;
        float *out = buf + (size_t)y * w * 4;
        for (long x = 0; x < (long)w; x++) {
            float r = 0.0f, g = 0.0f, b = 0.0f, a = 0.0f;
            for (int ki = 0; ki < len; ki++) {
                long sx = x + ki - radius;
                if (sx < 0) sx = 0;
                if (sx > (long)w - 1) sx = (long)w - 1;
                const uint8_t *p = row + sx * 4;
                float kv = k[ki];
                r += (float)p[0] * kv; g += (float)p[1] * kv;
                b += (float)p[2] * kv; a += (float)p[3] * kv;
            }
            out[x * 4] = r; out[x * 4 + 1] = g; out[x * 4 + 2] = b; out[x * 4 + 3] = a;
        }
    }
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++) {
        uint8_t *out = dst + (size_t)y * w * 4;
        for (long x = 0; x < (long)w; x++) {
            float r = 0.0f, g = 0.0f, b = 0.0f, a = 0.0f;
            for (int ki = 0; ki < len; ki++) {
                long sy = y + ki - radius;
                if (sy < 0) sy = 0;
                if (sy > (long)h - 1) sy = (long)h - 1;
                const float *p = buf + ((size_t)sy * w + (size_t)x) * 4;
                float kv = k[ki];
                r += p[0] * kv; g += p[1] * kv; b += p[2] * kv; a += p[3] * kv;
            }
            out[x * 4] = round_u8(r); out[x * 4 + 1] = round_u8(g);
            out[x * 4 + 2] = round_u8(b); out[x * 4 + 3] = round_u8(a);
        }
    }
    free(buf);
    free(k);
}

/* blur_with_selection :141-207. mask is w*h (GrayImage of canvas size) or NULL. */
void pfo_blur_with_selection(const uint8_t *src, uint32_t w, uint32_t h, float sigma,
                             const uint8_t *mask, uint8_t *dst) {
    if (!mask) { pfo_gaussian_blur(src, w, h, sigma, dst); if (w == 0 || h == 0) return; return; }
    uint32_t min_x = w, min_y = h, max_x = 0, max_y = 0;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++)
            if (mask[(size_t)y * w + x] > 0) {
                if (x < min_x) min_x = x;
                if (y < min_y) min_y = y;
                if (x > max_x) max_x = x;
                if (y > max_y) max_y = y;
            }
    memcpy(dst, src, (size_t)w * h * 4);
    if (min_x > max_x || min_y > max_y) return;
    uint32_t pad = as_u32(ceilf(sigma * 3.0f));
    uint32_t cx = min_x > pad ? min_x - pad : 0, cy = min_y > pad ? min_y - pad : 0;
    uint32_t cx2 = max_x + 1 + pad < w ? max_x + 1 + pad : w;
    uint32_t cy2 = max_y + 1 + pad < h ? max_y + 1 + pad : h;
    uint32_t cw = cx2 - cx, ch = cy2 - cy;
    uint8_t *sub = (uint8_t *)malloc((size_t)cw * ch * 4), *bl = (uint8_t *)malloc((size_t)cw * ch * 4);
    for (uint32_t y = 0; y < ch; y++)
        memcpy(sub + (size_t)y * cw * 4, src + ((size_t)(cy + y) * w + cx) * 4, (size_t)cw * 4);
    pfo_gaussian_blur(sub, cw, ch, sigma, bl);
    for (uint32_t y = min_y; y <= max_y; y++)
        for (uint32_t x = min_x; x <= max_x; x++)
            if (mask[(size_t)y * w + x] > 0)
                memcpy(dst + ((size_t)y * w + x) * 4, bl + ((size_t)(y - cy) * cw + (x - cx)) * 4, 4);
    free(sub);
    free(bl);
}

/* ------------------------------------------------------------------------- */
/* Box / motion / median / sharpen / vignette                                 */
/* ------------------------------------------------------------------------- */
static inline int masked_out(const uint8_t *mask, uint32_t w, size_t x, size_t y) {
    return mask && mask[y * w + x] == 0;
}

/* box_blur_core, src/ops/effects/blur.rs:233-318 */
void pfo_box_blur(const uint8_t *src, uint32_t w, uint32_t h, float radius, const uint8_t *mask,
                  uint8_t *dst) {
    size_t n4 = (size_t)w * h * 4;
    if (radius < 0.5f || w == 0 || h == 0) { memcpy(dst, src, n4); return; }
    int r = (int)as_u32(ceilf(radius));
    int ks = r * 2 + 1;
    uint32_t d = (uint32_t)ks;
    uint8_t *hb = (uint8_t *)malloc(n4);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++) {
        const uint8_t *row = src + (size_t)y * w * 4;
        uint8_t *out = hb + (size_t)y * w * 4;
        uint32_t s[4] = {0, 0, 0, 0};
        for (int k = 0; k < ks; k++) {
            int sx = clampi(k - r, 0, (int)w - 1);
            for (int c = 0; c < 4; c++) s[c] += row[sx * 4 + c];
        }
        for (long x = 0; x < (long)w; x++) {
            for (int c = 0; c < 4; c++) out[x * 4 + c] = (uint8_t)((s[c] + d / 2) / d);
            if (x + 1 < (long)w) {
                int rx = clampi((int)x - r, 0, (int)w - 1), ax = clampi((int)x + r + 1, 0, (int)w - 1);
                for (int c = 0; c < 4; c++) s[c] = s[c] - row[rx * 4 + c] + row[ax * 4 + c];
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (long x = 0; x < (long)w; x++) {
        uint32_t s[4] = {0, 0, 0, 0};
        for (int k = 0; k < ks; k++) {
            int sy = clampi(k - r, 0, (int)h - 1);
            for (int c = 0; c < 4; c++) s[c] += hb[((size_t)sy * w + x) * 4 + c];
        }
        for (long y = 0; y < (long)h; y++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) memcpy(dst + oi, src + oi, 4);
            else for (int c = 0; c < 4; c++) dst[oi + c] = (uint8_t)((s[c] + d / 2) / d);
            if (y + 1 < (long)h) {
                int ry = clampi((int)y - r, 0, (int)h - 1), ay = clampi((int)y + r + 1, 0, (int)h - 1);
                for (int c = 0; c < 4; c++)
                    s[c] = s[c] - hb[((size_t)ry * w + x) * 4 + c] + hb[((size_t)ay * w + x) * 4 + c];
            }
        }
    }
    free(hb);
}

/* f32::to_radians: self * (PI / 180.0) with both constants f32 */
float pfo_to_radians(float deg) { return deg * (3.14159265358979323846f / 180.0f); }

/* motion_blur_core, blur.rs:144-210 */
void pfo_motion_blur(const uint8_t *src, uint32_t w, uint32_t h, float angle_deg, float distance,
                     const uint8_t *mask, uint8_t *dst) {
    size_t n4 = (size_t)w * h * 4;
    if (distance < 1.0f || w == 0 || h == 0) { memcpy(dst, src, n4); return; }
    float angle = pfo_to_radians(angle_deg);
    int steps = as_i32(ceilf(distance));
    float dx = cosf(angle), dy = sinf(angle);
    float inv_steps = 1.0f / (float)(steps * 2 + 1);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            for (int i = -steps; i <= steps; i++) {
                int sx = as_i32(roundf((float)x + (float)i * dx));
                int sy = as_i32(roundf((float)y + (float)i * dy));
                sx = clampi(sx, 0, (int)w - 1);
                sy = clampi(sy, 0, (int)h - 1);
                const uint8_t *p = src + ((size_t)sy * w + sx) * 4;
                for (int c = 0; c < 4; c++) s[c] += (float)p[c];
            }
            for (int c = 0; c < 4; c++) dst[oi + c] = round_u8(s[c] * inv_steps);
        }
}

/* median_core, src/ops/effects/noise.rs:357-410.  The reference sorts the window and takes
 * element [len/2]; a 256-bin histogram walk gives the identical element for u8 data. */
void pfo_median(const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius, const uint8_t *mask,
                uint8_t *dst) {
    if (w == 0 || h == 0) return;
    int r = radius < 1 ? 1 : (int)radius;
    int len = (2 * r + 1) * (2 * r + 1);
    int target = len / 2;
#pragma omp parallel for schedule(dynamic, 4)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            uint16_t hist[4][256];
            memset(hist, 0, sizeof(hist));
            for (int dy = -r; dy <= r; dy++) {
                int sy = clampi((int)y + dy, 0, (int)h - 1);
                for (int dxx = -r; dxx <= r; dxx++) {
                    int sx = clampi((int)x + dxx, 0, (int)w - 1);
                    const uint8_t *p = src + ((size_t)sy * w + sx) * 4;
                    for (int c = 0; c < 4; c++) hist[c][p[c]]++;
                }
            }
            for (int c = 0; c < 4; c++) {
                int acc = 0, v = 0;
                for (v = 0; v < 256; v++) { acc += hist[c][v]; if (acc > target) break; }
                dst[oi + c] = (uint8_t)v;
            }
        }
}

/* sharpen_core (unsharp mask), src/ops/effects/stylize.rs:96-141 */
void pfo_sharpen(const uint8_t *src, uint32_t w, uint32_t h, float amount, float radius,
                 const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    size_t n = (size_t)w * h;
    uint8_t *bl = (uint8_t *)malloc(n * 4);
    pfo_gaussian_blur(src, w, h, radius, bl);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        size_t si = (size_t)i * 4;
        if (mask && mask[i] == 0) { memcpy(dst + si, src + si, 4); continue; }
        for (int c = 0; c < 3; c++) {
            float s = (float)src[si + c], b = (float)bl[si + c];
            dst[si + c] = round_u8(s + amount * (s - b));
        }
        dst[si + 3] = src[si + 3];
    }
    free(bl);
}

/* vignette_core, stylize.rs:170-191 via apply_per_pixel, effects.rs:53-100.
 * powf(2.0) is folded to x*x by LLVM (and x*x is the correctly rounded square). */
void pfo_vignette(const uint8_t *src, uint32_t w, uint32_t h, float amount, float softness,
                  const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    float fw = (float)w, fh = (float)h;
    float cx = fw / 2.0f, cy = fh / 2.0f;
    float max_dist = sqrtf(cx * cx + cy * cy);
    float soft = maxf(softness, 0.01f);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float dx = (float)x - cx, dy = (float)y - cy;
            float dist = sqrtf(dx * dx + dy * dy) / max_dist;
            float q = minf(dist / soft, 1.0f);
            float vf = clampf(1.0f - (amount * (q * q)), 0.0f, 1.0f);
            dst[oi] = round_u8((float)src[oi] * vf);
            dst[oi + 1] = round_u8((float)src[oi + 1] * vf);
            dst[oi + 2] = round_u8((float)src[oi + 2] * vf);
            dst[oi + 3] = round_u8((float)src[oi + 3]);
        }
}

/* ------------------------------------------------------------------------- */
/* Per-pixel adjustments: src/ops/adjustments.rs and src/ops/scripting.rs      */
/* ------------------------------------------------------------------------- */
/* rgb_to_hsl / hsl_to_rgb / hue_to_rgb, adjustments.rs:944-1012 */
static void rgb_to_hsl(float r, float g, float b, float *H, float *S, float *L) {
    float mx = maxf(maxf(r, g), b), mn = minf(minf(r, g), b);
    float l = (mx + mn) / 2.0f;
    if (fabsf(mx - mn) < 1e-6f) { *H = 0.0f; *S = 0.0f; *L = l; return; }
    float d = mx - mn;
    float s = l > 0.5f ? d / (2.0f - mx - mn) : d / (mx + mn);
    float hh;
    if (fabsf(mx - r) < 1e-6f) { hh = (g - b) / d; if (hh < 0.0f) hh += 6.0f; hh = hh / 6.0f; }
    else if (fabsf(mx - g) < 1e-6f) hh = ((b - r) / d + 2.0f) / 6.0f;
    else hh = ((r - g) / d + 4.0f) / 6.0f;
    *H = hh; *S = s; *L = l;
}
static float hue_to_rgb(float p, float q, float t) {
    if (t < 0.0f) t += 1.0f;
    if (t > 1.0f) t -= 1.0f;
    if (t < 1.0f / 6.0f) return p + (q - p) * 6.0f * t;
    if (t < 1.0f / 2.0f) return q;
    if (t < 2.0f / 3.0f) return p + (q - p) * (2.0f / 3.0f - t) * 6.0f;
    return p;
}
static void hsl_to_rgb(float h, float s, float l, float *r, float *g, float *b) {
    if (fabsf(s) < 1e-6f) { *r = *g = *b = l; return; }
    float q = l < 0.5f ? l * (1.0f + s) : l + s - l * s;
    float p = 2.0f * l - q;
    *r = hue_to_rgb(p, q, h + 1.0f / 3.0f);
    *g = hue_to_rgb(p, q, h);
    *b = hue_to_rgb(p, q, h - 1.0f / 3.0f);
}

/* build_levels_lut, adjustments.rs:424-446 */
void pfo_build_levels_lut(float in_black, float in_white, float gamma, float out_black,
                          float out_white, uint8_t lut[256]) {
    float in_range = maxf(in_white - in_black, 1.0f);
    float out_range = out_white - out_black;
    float inv_gamma = 1.0f / maxf(gamma, 0.01f);
    for (int i = 0; i < 256; i++) {
        float v = (float)i;
        float normalized = clampf((v - in_black) / in_range, 0.0f, 1.0f);
        float gc = powf(normalized, inv_gamma);
        float output = out_black + gc * out_range;
        lut[i] = round_u8(output);
    }
}
/* scripting apply_levels LUT, scripting.rs:1054-1066 (truncating, output range fixed 0..255) */
void pfo_build_levels_lut_script(float in_black, float in_white, float gamma, uint8_t lut[256]) {
    float in_range = maxf(in_white - in_black, 1.0f);
    float inv_gamma = 1.0f / maxf(gamma, 0.01f);
    for (int i = 0; i < 256; i++) {
        float normalized = clampf(((float)i - in_black) / in_range, 0.0f, 1.0f);
        float gc = powf(normalized, inv_gamma);
        lut[i] = as_u8(clampf(gc * 255.0f, 0.0f, 255.0f));
    }
}
/* build_stretch_lut, adjustments.rs:232-253 */
void pfo_build_stretch_lut(uint8_t mn, uint8_t mx, uint8_t lut[256]) {
    if (mx <= mn) { for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i; return; }
    float range = (float)(mx - mn);
    for (int i = 0; i < 256; i++) {
        float v;
        if ((uint8_t)i <= mn) v = 0.0f;
        else if ((uint8_t)i >= mx) v = 255.0f;
        else v = ((float)i - (float)mn) / range * 255.0f;
        lut[i] = round_u8(v);
    }
}
/* build_curves_lut (Fritsch-Carlson), adjustments.rs:634-729. pts = n (x,y) pairs. */
void pfo_build_curves_lut(const float *pts, int n, uint8_t lut[256]) {
    if (n < 2) { for (int i = 0; i < 256; i++) lut[i] = (uint8_t)i; return; }
    float *delta = (float *)malloc(sizeof(float) * (size_t)(n - 1));
    float *m = (float *)calloc((size_t)n, sizeof(float));
#define PX(i) pts[(i) * 2]
#define PY(i) pts[(i) * 2 + 1]
    for (int i = 0; i < n - 1; i++) {
        float dx = PX(i + 1) - PX(i), dy = PY(i + 1) - PY(i);
        delta[i] = fabsf(dx) < 1e-6f ? 0.0f : dy / dx;
    }
    m[0] = delta[0];
    m[n - 1] = delta[n - 2];
    for (int i = 1; i < n - 1; i++) {
        if (delta[i - 1] * delta[i] <= 0.0f) m[i] = 0.0f;
        else m[i] = (delta[i - 1] + delta[i]) / 2.0f;
    }
    for (int i = 0; i < n - 1; i++) {
        if (fabsf(delta[i]) < 1e-6f) { m[i] = 0.0f; m[i + 1] = 0.0f; }
        else {
            float alpha = m[i] / delta[i], beta = m[i + 1] / delta[i];
            float s = alpha * alpha + beta * beta;
            if (s > 9.0f) {
                float tau = 3.0f / sqrtf(s);
                m[i] = tau * alpha * delta[i];
                m[i + 1] = tau * beta * delta[i];
            }
        }
    }
    for (int i = 0; i < 256; i++) {
        float x = (float)i;
        int seg = 0;
        for (int j = 0; j < n - 1; j++) if (x >= PX(j)) seg = j;
        if (x <= PX(0)) lut[i] = round_u8(PY(0));
        else if (x >= PX(n - 1)) lut[i] = round_u8(PY(n - 1));
        else {
            float x0 = PX(seg), x1 = PX(seg + 1), y0 = PY(seg), y1 = PY(seg + 1);
            float hh = x1 - x0;
            if (fabsf(hh) < 1e-6f) lut[i] = round_u8(y0);
            else {
                float t = (x - x0) / hh, t2 = t * t, t3 = t2 * t;
                float h00 = 2.0f * t3 - 3.0f * t2 + 1.0f;
                float h10 = t3 - 2.0f * t2 + t;
                float h01 = -2.0f * t3 + 3.0f * t2;
                float h11 = t3 - t2;
                float val = h00 * y0 + h10 * hh * m[seg] + h01 * y1 + h11 * hh * m[seg + 1];
                lut[i] = round_u8(val);
            }
        }
    }
#undef PX
#undef PY
    free(delta);
    free(m);
}
/* build_multi_channel_luts, adjustments.rs:576-626. in: 5 LUTs [RGB,R,G,B,A] (identity when
 * the channel is disabled); out: 4 composed LUTs [R,G,B,A]. */
void pfo_compose_curve_luts(const uint8_t in[5][256], uint8_t out[4][256]) {
    for (int i = 0; i < 256; i++) {
        out[0][i] = in[1][in[0][i]];
        out[1][i] = in[2][in[0][i]];
        out[2][i] = in[3][in[0][i]];
        out[3][i] = in[4][i];
    }
}

/* Op ids shared with include/pfe_b200.h (PFE_ADJ_*). */
enum {
    PFO_INVERT = 0,         /* adjustments.rs:115  (255-r, 255-g, 255-b, a) */
    PFO_INVERT_ALPHA = 1,   /* :122 */
    PFO_SEPIA = 2,          /* :133 */
    PFO_DESATURATE = 3,     /* filters.rs:320 (BT.709, rounds) */
    PFO_BRIGHTNESS_CONTRAST = 4, /* adjustments.rs:265  p=[brightness, contrast] */
    PFO_HSL = 5,            /* :300  p=[hue, sat, light] */
    PFO_EXPOSURE = 6,       /* :352  p=[gain=2^ev] */
    PFO_LUT_RGB = 7,        /* levels :424 — one LUT on r,g,b; alpha kept */
    PFO_LUT_RGBA = 8,       /* curves :549 — four LUTs; per-channel levels :490 */
    PFO_TEMPERATURE_TINT = 9,   /* :518  p=[temperature, tint] */
    PFO_HIGHLIGHTS_SHADOWS = 10,/* :371  p=[shadows, highlights] */
    PFO_THRESHOLD = 11,     /* :1240 p=[level] */
    PFO_POSTERIZE = 12,     /* :1267 p=[factor = max(levels, 2)] */
    PFO_COLOR_BALANCE = 13, /* :1294 p=[shadows rgb, midtones rgb, highlights rgb] */
    PFO_GRADIENT_MAP = 14,  /* :1344 luts = 256 RGBA entries */
    PFO_BLACK_AND_WHITE = 15, /* :1373 p=[r_weight, g_weight, b_weight] */
    PFO_VIBRANCE = 16,      /* :1408 p=[amount / 100] */
    /* scripting.rs inline variants: truncating casts, no mask, alpha untouched */
    PFO_S_INVERT = 32,      /* scripting.rs:869 */
    PFO_S_DESATURATE = 33,  /* :883 integer 299/587/114 */
    PFO_S_SEPIA = 34,       /* :900 */
    PFO_S_SEPIA_STRENGTH = 35,  /* :921 p=[strength] */
    PFO_S_BRIGHTNESS_CONTRAST = 36, /* :944 */
    PFO_S_HSL = 37,         /* :964 */
    PFO_S_EXPOSURE = 38,    /* :1040 p=[gain] */
    PFO_S_LUT_RGB = 39      /* :1054 apply_levels */
};

static inline void adjust_pixel(int op, const float *p, const uint8_t *luts, const uint8_t in[4],
                                uint8_t out[4]) {
    float r = (float)in[0], g = (float)in[1], b = (float)in[2], a = (float)in[3];
    float nr = r, ng = g, nb = b, na = a;
    switch (op) {
    case PFO_INVERT: nr = 255.0f - r; ng = 255.0f - g; nb = 255.0f - b; break;
    case PFO_INVERT_ALPHA: na = 255.0f - a; break;
    case PFO_SEPIA:
        nr = minf(0.393f * r + 0.769f * g + 0.189f * b, 255.0f);
        ng = minf(0.349f * r + 0.686f * g + 0.168f * b, 255.0f);
        nb = minf(0.272f * r + 0.534f * g + 0.131f * b, 255.0f);
        break;
    case PFO_DESATURATE: {
        uint8_t lum = round_u8(0.2126f * r + 0.7152f * g + 0.0722f * b);
        out[0] = out[1] = out[2] = lum; out[3] = in[3];
        return;
    }
    case PFO_BRIGHTNESS_CONTRAST: {
        float brightness = p[0], contrast = p[1];
        float factor = (259.0f * (contrast + 255.0f)) / (255.0f * (259.0f - contrast));
        nr = factor * (r + brightness - 128.0f) + 128.0f;
        ng = factor * (g + brightness - 128.0f) + 128.0f;
        nb = factor * (b + brightness - 128.0f) + 128.0f;
    } break;
    case PFO_HSL: {
        float sat_factor = 1.0f + p[1] / 100.0f;
        float light_offset = p[2] * 255.0f / 100.0f;
        float hh, s, l;
        rgb_to_hsl(r / 255.0f, g / 255.0f, b / 255.0f, &hh, &s, &l);
        float t = hh + p[0] / 360.0f;
        float nh = t - truncf(t); /* f32::fract */
        if (nh < 0.0f) nh = nh + 1.0f;
        float ns = clampf(s * sat_factor, 0.0f, 1.0f);
        float rr, gg, bb;
        hsl_to_rgb(nh, ns, l, &rr, &gg, &bb);
        nr = rr * 255.0f + light_offset; ng = gg * 255.0f + light_offset; nb = bb * 255.0f + light_offset;
    } break;
    case PFO_EXPOSURE: nr = r * p[0]; ng = g * p[0]; nb = b * p[0]; break;
    case PFO_LUT_RGB:
        nr = (float)luts[in[0]]; ng = (float)luts[in[1]]; nb = (float)luts[in[2]];
        break;
    case PFO_LUT_RGBA:
        nr = (float)luts[in[0]]; ng = (float)luts[256 + in[1]];
        nb = (float)luts[512 + in[2]]; na = (float)luts[768 + in[3]];
        break;
    case PFO_TEMPERATURE_TINT: {
        float temp_shift = p[0] * 1.5f, tint_shift = p[1] * 1.0f;
        nr = r + temp_shift; ng = g - tint_shift * 0.5f; nb = b - temp_shift;
    } break;
    case PFO_HIGHLIGHTS_SHADOWS: {
        float shadow_amt = p[0] / 100.0f, highlight_amt = p[1] / 100.0f;
        float lum = (0.2126f * r + 0.7152f * g + 0.0722f * b) / 255.0f;
        float sw = (1.0f - lum) * (1.0f - lum);
        float hw = lum * lum;
        float adj = sw * shadow_amt * 128.0f + hw * highlight_amt * 128.0f;
        nr = r + adj; ng = g + adj; nb = b + adj;
    } break;
    case PFO_THRESHOLD: {
        float lum = 0.2126f * r + 0.7152f * g + 0.0722f * b;
        nr = ng = nb = lum >= p[0] ? 255.0f : 0.0f;
    } break;
    case PFO_POSTERIZE: {
        float f1 = p[0] - 1.0f;
        nr = roundf(r / 255.0f * f1) / f1 * 255.0f;
        ng = roundf(g / 255.0f * f1) / f1 * 255.0f;
        nb = roundf(b / 255.0f * f1) / f1 * 255.0f;
    } break;
    case PFO_COLOR_BALANCE: { /* color_balance_pixel :1321-1338; powi(2) == x*x */
        float lum = (0.2126f * r + 0.7152f * g + 0.0722f * b) / 255.0f;
        float s0 = maxf(1.0f - lum * 2.0f, 0.0f), h0 = maxf(lum * 2.0f - 1.0f, 0.0f);
        float sw = s0 * s0, hw = h0 * h0;
        float mw = maxf(1.0f - sw - hw, 0.0f);
        nr = r + (sw * p[0] + mw * p[3] + hw * p[6]) * 1.28f;
        ng = g + (sw * p[1] + mw * p[4] + hw * p[7]) * 1.28f;
        nb = b + (sw * p[2] + mw * p[5] + hw * p[8]) * 1.28f;
    } break;
    case PFO_GRADIENT_MAP: {
        uint32_t lum = as_u32(0.2126f * r + 0.7152f * g + 0.0722f * b);
        if (lum > 255) lum = 255;
        nr = (float)luts[lum * 4]; ng = (float)luts[lum * 4 + 1]; nb = (float)luts[lum * 4 + 2];
    } break;
    case PFO_BLACK_AND_WHITE: {
        float v = clampf((r * p[0] + g * p[1] + b * p[2]) / 100.0f, 0.0f, 255.0f);
        nr = ng = nb = v;
    } break;
    case PFO_VIBRANCE: { /* vibrance_pixel :1431-1444 */
        float hh, sat, l;
        rgb_to_hsl(r / 255.0f, g / 255.0f, b / 255.0f, &hh, &sat, &l);
        float boost = p[0] >= 0.0f ? p[0] * ((1.0f - sat) * (1.0f - sat)) : p[0] * (sat * sat);
        float ns = clampf(sat + boost, 0.0f, 1.0f);
        float rr, gg, bb;
        hsl_to_rgb(hh, ns, l, &rr, &gg, &bb);
        nr = rr * 255.0f; ng = gg * 255.0f; nb = bb * 255.0f;
    } break;
    /* ---- scripting variants: write u8 directly ---- */
    case PFO_S_INVERT: out[0] = 255 - in[0]; out[1] = 255 - in[1]; out[2] = 255 - in[2]; out[3] = in[3]; return;
    case PFO_S_DESATURATE: {
        uint32_t gray = ((uint32_t)in[0] * 299u + (uint32_t)in[1] * 587u + (uint32_t)in[2] * 114u) / 1000u;
        out[0] = out[1] = out[2] = (uint8_t)gray; out[3] = in[3];
        return;
    }
    case PFO_S_SEPIA:
        out[0] = as_u8(minf(r * 0.393f + g * 0.769f + b * 0.189f, 255.0f));
        out[1] = as_u8(minf(r * 0.349f + g * 0.686f + b * 0.168f, 255.0f));
        out[2] = as_u8(minf(r * 0.272f + g * 0.534f + b * 0.131f, 255.0f));
        out[3] = in[3];
        return;
    case PFO_S_SEPIA_STRENGTH: {
        float strength = p[0], inv = 1.0f - strength;
        float sr = minf(r * 0.393f + g * 0.769f + b * 0.189f, 255.0f);
        float sg = minf(r * 0.349f + g * 0.686f + b * 0.168f, 255.0f);
        float sb = minf(r * 0.272f + g * 0.534f + b * 0.131f, 255.0f);
        out[0] = as_u8(r * inv + sr * strength);
        out[1] = as_u8(g * inv + sg * strength);
        out[2] = as_u8(b * inv + sb * strength);
        out[3] = in[3];
        return;
    }
    case PFO_S_BRIGHTNESS_CONTRAST: {
        float bright = p[0], contrast = p[1];
        float factor = (259.0f * (contrast + 255.0f)) / (255.0f * (259.0f - contrast));
        out[0] = as_u8(clampf(factor * (r + bright - 128.0f) + 128.0f, 0.0f, 255.0f));
        out[1] = as_u8(clampf(factor * (g + bright - 128.0f) + 128.0f, 0.0f, 255.0f));
        out[2] = as_u8(clampf(factor * (b + bright - 128.0f) + 128.0f, 0.0f, 255.0f));
        out[3] = in[3];
        return;
    }
    case PFO_S_HSL: {
        float hue_shift = p[0];
        float sat_factor = 1.0f + p[1] / 100.0f;
        float light_offset = p[2] * 255.0f / 100.0f;
        float fr = r / 255.0f, fg = g / 255.0f, fb = b / 255.0f;
        float cmax = maxf(maxf(fr, fg), fb), cmin = minf(minf(fr, fg), fb);
        float l = (cmax + cmin) / 2.0f;
        float hh = 0.0f, s = 0.0f;
        if (!(fabsf(cmax - cmin) < 1e-10f)) {
            float d = cmax - cmin;
            s = l > 0.5f ? d / (2.0f - cmax - cmin) : d / (cmax + cmin);
            float h6;
            if (fabsf(cmax - fr) < 1e-10f) h6 = (fg - fb) / d + (fg < fb ? 6.0f : 0.0f);
            else if (fabsf(cmax - fg) < 1e-10f) h6 = (fb - fr) / d + 2.0f;
            else h6 = (fr - fg) / d + 4.0f;
            hh = h6 / 6.0f;
        }
        float t = hh + hue_shift / 360.0f;
        float nh = fmodf(t, 1.0f); /* rem_euclid(1.0) */
        if (nh < 0.0f) nh = nh + 1.0f;
        float ns = clampf(s * sat_factor, 0.0f, 1.0f);
        float rr, gg, bb;
        if (fabsf(ns) < 1e-10f) { rr = gg = bb = l; }
        else {
            float q = l < 0.5f ? l * (1.0f + ns) : l + ns - l * ns;
            float pp = 2.0f * l - q;
            rr = hue_to_rgb(pp, q, nh + 1.0f / 3.0f);
            gg = hue_to_rgb(pp, q, nh);
            bb = hue_to_rgb(pp, q, nh - 1.0f / 3.0f);
        }
        out[0] = as_u8(clampf(rr * 255.0f + light_offset, 0.0f, 255.0f));
        out[1] = as_u8(clampf(gg * 255.0f + light_offset, 0.0f, 255.0f));
        out[2] = as_u8(clampf(bb * 255.0f + light_offset, 0.0f, 255.0f));
        out[3] = in[3];
        return;
    }
    case PFO_S_EXPOSURE:
        out[0] = as_u8(clampf(r * p[0], 0.0f, 255.0f));
        out[1] = as_u8(clampf(g * p[0], 0.0f, 255.0f));
        out[2] = as_u8(clampf(b * p[0], 0.0f, 255.0f));
        out[3] = in[3];
        return;
    case PFO_S_LUT_RGB:
        out[0] = luts[in[0]]; out[1] = luts[in[1]]; out[2] = luts[in[2]]; out[3] = in[3];
        return;
    default: break;
    }
    /* apply_pixel_transform rounding, adjustments.rs:33-39 */
    out[0] = round_u8(nr); out[1] = round_u8(ng); out[2] = round_u8(nb); out[3] = round_u8(na);
}

/* apply_pixel_transform_from_flat (adjustments.rs:46-108) when `occupancy` is NULL;
 * apply_pixel_transform / par_map_populated (adjustments.rs:21-42, tiled_image.rs:905-933)
 * when `occupancy` (chunk bitmap) is given: pixels in unpopulated chunks are left untouched. */
void pfo_adjust(const uint8_t *src, uint32_t w, uint32_t h, int op, const float *params,
                const uint8_t *luts, const uint8_t *mask, const uint8_t *occupancy, uint8_t *dst) {
    uint32_t cxn = (w + PFE_CHUNK - 1) / PFE_CHUNK;
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if ((occupancy && !occupancy[(size_t)(y / PFE_CHUNK) * cxn + (size_t)(x / PFE_CHUNK)]) ||
                (op < 32 && masked_out(mask, w, (size_t)x, (size_t)y))) {
                memcpy(dst + oi, src + oi, 4);
                continue;
            }
            adjust_pixel(op, params, luts, src + oi, dst + oi);
        }
}

/* ------------------------------------------------------------------------- */
/* Warps: src/ops/transform.rs:1015-1345, 1558-1761                           */
/* ------------------------------------------------------------------------- */
static inline void cr_weights(float t, float w[4]) { /* :1558 */
    float t2 = t * t, t3 = t2 * t;
    w[0] = -0.5f * t3 + t2 - 0.5f * t;
    w[1] = 1.5f * t3 - 2.5f * t2 + 1.0f;
    w[2] = -1.5f * t3 + 2.0f * t2 + 0.5f * t;
    w[3] = 0.5f * t3 - 0.5f * t2;
}
void pfo_catmull_rom_weights(float t, float w[4]) { cr_weights(t, w); }

/* catmull_rom_surface :1589-1646. points row-major (rows+1)x(cols+1) of [x,y]. */
void pfo_catmull_rom_surface(const float *pts, int cols, int rows, float ug, float vg, float out[2]) {
    int ppr = cols + 1, nrows = rows + 1;
    float col_f = clampf(ug, 0.0f, (float)cols - 0.0001f);
    float row_f = clampf(vg, 0.0f, (float)rows - 0.0001f);
    int ci = (int)as_u32(col_f); if (ci > cols - 1) ci = cols - 1;
    int ri = (int)as_u32(row_f); if (ri > rows - 1) ri = rows - 1;
    float ul = col_f - (float)ci, vl = row_f - (float)ri;
    float wv[4], wu[4];
    cr_weights(vl, wv);
    int rv[4] = {ri == 0 ? 0 : ri - 1, ri, ri + 1 < nrows - 1 ? ri + 1 : nrows - 1,
                 ri + 2 < nrows - 1 ? ri + 2 : nrows - 1};
    cr_weights(ul, wu);
    int cu[4] = {ci == 0 ? 0 : ci - 1, ci, ci + 1 < ppr - 1 ? ci + 1 : ppr - 1,
                 ci + 2 < ppr - 1 ? ci + 2 : ppr - 1};
    float rvx[4], rvy[4];
    for (int j = 0; j < 4; j++) {
        const float *b = pts + (size_t)rv[j] * ppr * 2;
        const float *p0 = b + cu[0] * 2, *p1 = b + cu[1] * 2, *p2 = b + cu[2] * 2, *p3 = b + cu[3] * 2;
        rvx[j] = wu[0] * p0[0] + wu[1] * p1[0] + wu[2] * p2[0] + wu[3] * p3[0];
        rvy[j] = wu[0] * p0[1] + wu[1] * p1[1] + wu[2] * p2[1] + wu[3] * p3[1];
    }
    out[0] = wv[0] * rvx[0] + wv[1] * rvx[1] + wv[2] * rvx[2] + wv[3] * rvx[3];
    out[1] = wv[0] * rvy[0] + wv[1] * rvy[1] + wv[2] * rvy[2] + wv[3] * rvy[3];
}

/* generate_displacement_from_mesh :1670-1706 (orig != NULL) and _fast :1712-1739 (orig NULL) */
void pfo_mesh_displacement(const float *orig, const float *def, int cols, int rows, uint32_t w,
                           uint32_t h, float *out) {
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            float ug = ((float)x + 0.5f) / (float)w * (float)cols;
            float vg = ((float)y + 0.5f) / (float)h * (float)rows;
            float d[2];
            pfo_catmull_rom_surface(def, cols, rows, ug, vg, d);
            float *o = out + ((size_t)y * w + x) * 2;
            if (orig) {
                float og[2];
                pfo_catmull_rom_surface(orig, cols, rows, ug, vg, og);
                o[0] = d[0] - og[0]; o[1] = d[1] - og[1];
            } else {
                o[0] = d[0] - ((float)x + 0.5f); o[1] = d[1] - ((float)y + 0.5f);
            }
        }
}

/* warp_displacement_full :1288-1345. src is sw x sh; disp / dst are w x h. */
void pfo_warp_displacement(const uint8_t *src, uint32_t sw, uint32_t sh, const float *disp,
                           uint32_t w, uint32_t h, uint8_t *dst) {
    int src_w = (int)sw, src_h = (int)sh;
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t i = (size_t)y * w + x;
            uint8_t *o = dst + i * 4;
            float sx = (float)x - disp[i * 2], sy = (float)y - disp[i * 2 + 1];
            int x0 = as_i32(floorf(sx)), y0 = as_i32(floorf(sy));
            if (x0 < -1 || y0 < -1 || x0 >= src_w || y0 >= src_h) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
            float fx = sx - (float)x0, fy = sy - (float)y0;
            float tl[4], tr[4], bl[4], br[4];
            int xs[2] = {x0, x0 + 1}, ys[2] = {y0, y0 + 1};
            float *q[4] = {tl, tr, bl, br};
            for (int k = 0; k < 4; k++) {
                int px = xs[k & 1], py = ys[k >> 1];
                if (px < 0 || py < 0 || px >= src_w || py >= src_h) { q[k][0] = q[k][1] = q[k][2] = q[k][3] = 0.0f; }
                else { const uint8_t *s = src + ((size_t)py * sw + px) * 4; for (int c = 0; c < 4; c++) q[k][c] = (float)s[c]; }
            }
            for (int c = 0; c < 4; c++) {
                float top = tl[c] + (tr[c] - tl[c]) * fx;
                float bot = bl[c] + (br[c] - bl[c]) * fx;
                o[c] = round_u8(top + (bot - top) * fy);
            }
        }
}

/* warp_displacement_region :1206-1285: prev everywhere, the warp inside dirty_rect = (x0, y0, x1, y1).  The rect is
 * clamped the way the reference does it: x0.max(0), (x1 as u32).min(out_w) - a negative x1 wraps to a huge u32 and
 * clamps to the full width.  An inverted rect (the reference would panic on the row length) is treated as empty. */
void pfo_warp_displacement_region(const uint8_t *src, uint32_t sw, uint32_t sh, const float *disp, const uint8_t *prev,
                                  const int rect[4], uint32_t w, uint32_t h, uint8_t *dst) {
    uint32_t dx0 = rect[0] > 0 ? (uint32_t)rect[0] : 0u, dy0 = rect[1] > 0 ? (uint32_t)rect[1] : 0u;
    uint32_t dx1 = (uint32_t)rect[2] < w ? (uint32_t)rect[2] : w, dy1 = (uint32_t)rect[3] < h ? (uint32_t)rect[3] : h;
    memcpy(dst, prev, (size_t)w * h * 4);
    if (dx1 <= dx0 || dy1 <= dy0) return;
    uint8_t *full = (uint8_t *)malloc((size_t)w * h * 4);
    pfo_warp_displacement(src, sw, sh, disp, w, h, full);  /* per-pixel formula is the full warp's (:1288-1345) */
    for (uint32_t y = dy0; y < dy1; y++)
        memcpy(dst + ((size_t)y * w + dx0) * 4, full + ((size_t)y * w + dx0) * 4, (size_t)(dx1 - dx0) * 4);
    free(full);
}

/* warp_mesh_catmull_rom :1743-1761 */
void pfo_mesh_warp(const uint8_t *src, uint32_t sw, uint32_t sh, const float *orig, const float *def,
                   int cols, int rows, uint32_t w, uint32_t h, uint8_t *dst) {
    float *d = (float *)malloc((size_t)w * h * 2 * sizeof(float));
    pfo_mesh_displacement(orig, def, cols, rows, w, h, d);
    pfo_warp_displacement(src, sw, sh, d, w, h, dst);
    free(d);
}

/* DisplacementField::apply_push/expand/contract/twirl :1051-1200.
 * kind: 0 push (a0=delta_x, a1=delta_y), 1 expand, 2 contract, 3 twirl (a0 = clockwise?1:0). */
void pfo_liquify(float *field, uint32_t w, uint32_t h, int kind, float cx, float cy, float radius,
                 float strength, float a0, float a1, int bbox[4]) {
    float r = maxf(radius, 1.0f);
    float sigma = r / 3.0f;
    float s2 = 2.0f * sigma * sigma;
    int x0 = as_i32(floorf(cx - r)); if (x0 < 0) x0 = 0;
    int y0 = as_i32(floorf(cy - r)); if (y0 < 0) y0 = 0;
    int x1 = as_i32(ceilf(cx + r)); if (x1 > (int)w) x1 = (int)w;
    int y1 = as_i32(ceilf(cy + r)); if (y1 > (int)h) y1 = (int)h;
    float dir = a0 != 0.0f ? 1.0f : -1.0f;
    for (int py = y0; py < y1; py++)
        for (int px = x0; px < x1; px++) {
            float dx = (float)px - cx, dy = (float)py - cy;
            float dsq = dx * dx + dy * dy;
            if (dsq > r * r) continue;
            float *f = field + ((size_t)py * w + px) * 2;
            if (kind == 0) {
                float wgt = expf(-dsq / s2) * strength;
                f[0] += a0 * wgt; f[1] += a1 * wgt;
            } else if (kind == 1) {
                float dist = maxf(sqrtf(dsq), 0.001f);
                float t = dist / r;
                float wgt = (1.0f - t) * (1.0f - t) * strength * 3.0f;
                f[0] += dx / dist * wgt; f[1] += dy / dist * wgt;
            } else if (kind == 2) {
                float dist = maxf(sqrtf(dsq), 0.001f);
                float wgt = expf(-dsq / s2) * strength;
                f[0] += -dx / dist * wgt * 2.0f; f[1] += -dy / dist * wgt * 2.0f;
            } else {
                float wgt = expf(-dsq / s2) * strength * dir;
                f[0] += -dy * wgt * 0.1f; f[1] += dx * wgt * 0.1f;
            }
        }
    if (bbox) { bbox[0] = x0; bbox[1] = y0; bbox[2] = x1; bbox[3] = y1; }
}

/* ------------------------------------------------------------------------- */
/* Brush stamp: src/ui/panels/tools/behavior/raster/brush_render.rs           */
/* ------------------------------------------------------------------------- */
typedef struct pfo_brush {
    float size, hardness, flow;   /* ToolProperties: state.rs:136-150 (pressure off) */
    int anti_aliased;
    float color[4];               /* brush colour, straight RGBA in 0..1 */
    int is_eraser;
    int mode;                     /* BrushMode: 0 Normal, 1 Dodge, 2 Burn, 3 Sponge (:362-395) */
} pfo_brush;

/* compute_brush_alpha :54-82 */
float pfo_brush_alpha(const pfo_brush *b, float dist, float radius) {
    if (radius <= 0.0f) return 0.0f;
    float sh = clampf(b->hardness, 0.0f, 1.0f);
    float t = clampf(dist / radius, 0.0f, 1.0f);
    float falloff = t * t * (3.0f - 2.0f * t);
    float material = 1.0f + (sh - 1.0f) * falloff;
    float coverage;
    if (b->anti_aliased) {
        float e0 = radius + 0.5f, e1 = radius - 0.5f;
        if (dist <= e1) coverage = 1.0f;
        else if (dist >= e0) coverage = 0.0f;
        else { float x = clampf((dist - e0) / (e1 - e0), 0.0f, 1.0f); coverage = x * x * (3.0f - 2.0f * x); }
    } else coverage = dist <= radius ? 1.0f : 0.0f;
    return material * coverage;
}
/* rebuild_brush_lut :27-50 */
void pfo_brush_lut(const pfo_brush *b, uint8_t lut[256]) {
    float radius = b->size / 2.0f;
    if (radius < 0.001f) { memset(lut, 0, 256); return; }
    for (int i = 0; i < 256; i++) {
        float t_sq = (float)i / 255.0f;
        float dist = sqrtf(t_sq) * radius;
        float alpha = pfo_brush_alpha(b, dist, radius);
        lut[i] = as_u8(minf(roundf(alpha * 255.0f), 255.0f));
    }
}
/* draw_circle_no_dirty :135-400, circle tip, every BrushMode and the eraser, no scatter/jitter,
 * on a flat w*h RGBA8 target (the chunk walk only decides which tiles get allocated). */
void pfo_brush_stamp(uint8_t *img, uint32_t w, uint32_t h, const pfo_brush *b, float cx, float cy,
                     const uint8_t *sel_mask) {
    uint8_t lut[256];
    pfo_brush_lut(b, lut);
    float radius = b->size / 2.0f, radius_sq = radius * radius;
    if (radius_sq < 0.001f) return;
    float draw_radius = b->anti_aliased ? radius + 0.5f : radius;
    float draw_radius_sq = draw_radius * draw_radius;
    int direct = draw_radius > radius;
    float inv_radius_sq = 1.0f / radius_sq;
    uint32_t min_x = as_u32(maxf(floorf(cx - draw_radius), 0.0f));
    uint32_t max_x = as_u32(ceilf(cx + draw_radius)); if (max_x > (w ? w - 1 : 0)) max_x = w ? w - 1 : 0;
    uint32_t min_y = as_u32(maxf(floorf(cy - draw_radius), 0.0f));
    uint32_t max_y = as_u32(ceilf(cy + draw_radius)); if (max_y > (h ? h - 1 : 0)) max_y = h ? h - 1 : 0;
    if (min_x > max_x || min_y > max_y) return;
    uint8_t r8 = as_u8(b->color[0] * 255.0f), g8 = as_u8(b->color[1] * 255.0f), b8 = as_u8(b->color[2] * 255.0f);
    float src_a = b->color[3];
    for (uint32_t gy = min_y; gy <= max_y; gy++) {
        float dy = (float)gy - cy, dy_sq = dy * dy;
        for (uint32_t gx = min_x; gx <= max_x; gx++) {
            if (sel_mask && sel_mask[(size_t)gy * w + gx] == 0) continue;
            float dx = (float)gx - cx;
            float dist_sq = dx * dx + dy_sq;
            if (dist_sq > draw_radius_sq) continue;
            uint8_t ga8;
            if (direct) ga8 = as_u8(minf(roundf(pfo_brush_alpha(b, sqrtf(dist_sq), radius) * 255.0f), 255.0f));
            else ga8 = lut[as_u32(minf(dist_sq * inv_radius_sq * 255.0f, 255.0f))];
            if (ga8 == 0) continue;
            float ga = (float)ga8 / 255.0f;
            uint8_t *p = img + ((size_t)gy * w + gx) * 4;
            float strength = ga * src_a * b->flow;
            if (strength < 0.01f) continue;
            if (b->is_eraser) {
                float old_mask = (float)p[3] / 255.0f;
                if (strength > old_mask) { p[0] = p[1] = p[2] = 0; p[3] = as_u8(strength * 255.0f); }
            } else {
                if (b->mode == 0) {
                    uint8_t a8 = as_u8(strength * 255.0f);
                    if (a8 >= p[3]) { p[0] = r8; p[1] = g8; p[2] = b8; p[3] = a8; }
                } else { /* Dodge / Burn / Sponge :374-394: HSL edit of the existing pixel, alpha kept */
                    float hh, sat, l, nr, ng, nb, st = strength * 0.5f;
                    rgb_to_hsl((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, &hh, &sat, &l);
                    if (b->mode == 1) l = clampf(l + st, 0.0f, 1.0f);
                    else if (b->mode == 2) l = clampf(l - st, 0.0f, 1.0f);
                    else if (b->mode == 3) sat = clampf(sat - st, 0.0f, 1.0f);
                    hsl_to_rgb(hh, sat, l, &nr, &ng, &nb);
                    p[0] = as_u8(nr * 255.0f); p[1] = as_u8(ng * 255.0f); p[2] = as_u8(nb * 255.0f);
                }
            }
        }
    }
}
/* draw_line_no_dirty :762-838 (circle tip: 1-px stepping). Writes stamp centres into
 * `centres` (x,y pairs; capacity cap) and returns their count, so the same list can be fed to
 * the device kernel. */
int pfo_brush_line_centres(uint32_t w, uint32_t h, float x0, float y0, float x1, float y1,
                           float *centres, int cap) {
    float dx = x1 - x0, dy = y1 - y0;
    float distance = sqrtf(dx * dx + dy * dy);
    int n = 0;
    if (distance < 0.1f) {
        if (x0 >= 0.0f && as_u32(x0) < w && y0 >= 0.0f && as_u32(y0) < h && n < cap) {
            centres[0] = x0; centres[1] = y0; n = 1;
        }
        return n;
    }
    float step = 1.0f;
    uint32_t steps = as_u32(ceilf(distance / step));
    for (uint32_t i = 0; i <= steps; i++) {
        float t = (float)i / (float)steps;
        float x = x0 + dx * t, y = y0 + dy * t;
        if (x >= 0.0f && as_u32(x) < w && y >= 0.0f && as_u32(y) < h && n < cap) {
            centres[n * 2] = x; centres[n * 2 + 1] = y; n++;
        }
    }
    return n;
}
void pfo_brush_line(uint8_t *img, uint32_t w, uint32_t h, const pfo_brush *b, float x0, float y0,
                    float x1, float y1, const uint8_t *sel_mask) {
    float dx = x1 - x0, dy = y1 - y0;
    int cap = (int)ceilf(sqrtf(dx * dx + dy * dy)) + 4;
    float *c = (float *)malloc(sizeof(float) * 2 * (size_t)cap);
    int n = pfo_brush_line_centres(w, h, x0, y0, x1, y1, c, cap);
    for (int i = 0; i < n; i++) pfo_brush_stamp(img, w, h, b, c[i * 2], c[i * 2 + 1], sel_mask);
    free(c);
}

/* ------------------------------------------------------------------------- */
/* TiledImage chunk semantics: src/canvas/tiled_image.rs:50-104, 271-293       */
/* ------------------------------------------------------------------------- */
/* from_rgba_image keeps a chunk only if some pixel in it has alpha != 0; to_rgba_image of
 * the result therefore zeroes RGB in chunks that were fully transparent. occupancy is
 * ceil(w/64)*ceil(h/64) bytes; dst may be NULL to compute occupancy only. */
void pfo_tiled_roundtrip(const uint8_t *src, uint32_t w, uint32_t h, uint8_t *occupancy, uint8_t *dst) {
    uint32_t cxn = (w + PFE_CHUNK - 1) / PFE_CHUNK, cyn = (h + PFE_CHUNK - 1) / PFE_CHUNK;
    for (uint32_t cy = 0; cy < cyn; cy++)
        for (uint32_t cx = 0; cx < cxn; cx++) {
            uint32_t x0 = cx * PFE_CHUNK, y0 = cy * PFE_CHUNK;
            uint32_t x1 = x0 + PFE_CHUNK < w ? x0 + PFE_CHUNK : w, y1 = y0 + PFE_CHUNK < h ? y0 + PFE_CHUNK : h;
            int has = 0;
            for (uint32_t y = y0; y < y1 && !has; y++)
                for (uint32_t x = x0; x < x1; x++)
                    if (src[((size_t)y * w + x) * 4 + 3] != 0) { has = 1; break; }
            if (occupancy) occupancy[(size_t)cy * cxn + cx] = (uint8_t)has;
            if (dst)
                for (uint32_t y = y0; y < y1; y++) {
                    size_t o = ((size_t)y * w + x0) * 4;
                    if (has) memcpy(dst + o, src + o, (size_t)(x1 - x0) * 4);
                    else memset(dst + o, 0, (size_t)(x1 - x0) * 4);
                }
        }
}

/* bench.py's CPU arm sets the thread count explicitly: a launcher (torchrun) exports OMP_NUM_THREADS=1, which
 * would otherwise turn the "all host cores" baseline into a single-thread run. */
void pfo_set_num_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int pfo_num_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ========================================================================= */
/* Widened scope: remaining Rhai Effect-API kernels        */
/* ========================================================================= */

/* glow_core, src/ops/effects/stylize.rs:26-76 */
void pfo_glow(const uint8_t *src, uint32_t w, uint32_t h, float radius, float intensity,
              const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    size_t n = (size_t)w * h;
    uint8_t *bl = (uint8_t *)malloc(n * 4);
    pfo_gaussian_blur(src, w, h, radius, bl);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        size_t si = (size_t)i * 4;
        if (mask && mask[i] == 0) { memcpy(dst + si, src + si, 4); continue; }
        for (int c = 0; c < 3; c++) {
            float s = (float)src[si + c] / 255.0f, b = (float)bl[si + c] / 255.0f;
            float result = 1.0f - (1.0f - s) * (1.0f - b * intensity);
            dst[si + c] = round_u8(result * 255.0f);
        }
        dst[si + 3] = src[si + 3];
    }
    free(bl);
}

/* pixelate_core, src/ops/effects/distort.rs:333-373 */
void pfo_pixelate(const uint8_t *src, uint32_t w, uint32_t h, uint32_t block_size, const uint8_t *mask,
                  uint8_t *dst) {
    if (w == 0 || h == 0) return;
    uint32_t bs = block_size < 2 ? 2 : block_size;
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            uint32_t bx = ((uint32_t)x / bs) * bs + bs / 2, by = ((uint32_t)y / bs) * bs + bs / 2;
            uint32_t sx = bx < w - 1 ? bx : w - 1, sy = by < h - 1 ? by : h - 1;
            memcpy(dst + oi, src + ((size_t)sy * w + sx) * 4, 4);
        }
}

/* sample_clamped / sample_bilinear, src/ops/effects.rs:109-141 */
static inline void sample_clamped(const uint8_t *src, int w, int h, int x, int y, float p[4]) {
    int cx = clampi(x, 0, w - 1), cy = clampi(y, 0, h - 1);
    const uint8_t *s = src + ((size_t)cy * w + cx) * 4;
    for (int c = 0; c < 4; c++) p[c] = (float)s[c];
}
static inline void sample_bilinear(const uint8_t *src, int w, int h, float fx, float fy, float out[4]) {
    int x0 = as_i32(floorf(fx)), y0 = as_i32(floorf(fy));
    int x1 = x0 + 1, y1 = y0 + 1;
    float dx = fx - (float)x0, dy = fy - (float)y0;
    float p00[4], p10[4], p01[4], p11[4];
    sample_clamped(src, w, h, x0, y0, p00); sample_clamped(src, w, h, x1, y0, p10);
    sample_clamped(src, w, h, x0, y1, p01); sample_clamped(src, w, h, x1, y1, p11);
    for (int c = 0; c < 4; c++)
        out[c] = p00[c] * (1.0f - dx) * (1.0f - dy) + p10[c] * dx * (1.0f - dy) + p01[c] * (1.0f - dx) * dy + p11[c] * dx * dy;
}

/* bulge_core_at, distort.rs:400-437 */
void pfo_bulge(const uint8_t *src, uint32_t w, uint32_t h, float amount, float ox, float oy,
               const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    float fw = (float)w, fh = (float)h;
    float cx = clampf(ox, 0.0f, 1.0f) * maxf(fw - 1.0f, 0.0f), cy = clampf(oy, 0.0f, 1.0f) * maxf(fh - 1.0f, 0.0f);
    float max_r = maxf(maxf(maxf(cx, fw - cx), maxf(cy, fh - cy)), 1.0f);
    float strength = maxf(fabsf(amount), 0.0001f);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float dx = (float)x - cx, dy = (float)y - cy;
            float dist = sqrtf(dx * dx + dy * dy);
            float norm = minf(dist / max_r, 1.0f);
            float p[4];
            if (norm >= 1.0f) sample_clamped(src, (int)w, (int)h, (int)x, (int)y, p);
            else {
                float falloff = 1.0f - norm;
                float factor = amount > 0.0f ? 1.0f - falloff * strength * 0.5f : (amount < 0.0f ? 1.0f + falloff * strength * 0.5f : 1.0f);
                sample_bilinear(src, (int)w, (int)h, cx + dx * factor, cy + dy * factor, p);
            }
            for (int c = 0; c < 4; c++) dst[oi + c] = round_u8(p[c]);
        }
}

/* twist_core_at, distort.rs:464-493 */
void pfo_twist(const uint8_t *src, uint32_t w, uint32_t h, float angle_deg, float ox, float oy,
               const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    float fw = (float)w, fh = (float)h;
    float cx = clampf(ox, 0.0f, 1.0f) * maxf(fw - 1.0f, 0.0f), cy = clampf(oy, 0.0f, 1.0f) * maxf(fh - 1.0f, 0.0f);
    float mx = maxf(cx, fw - cx), my = maxf(cy, fh - cy);
    float max_r = maxf(sqrtf(mx * mx + my * my), 1.0f);
    float twist_amount = pfo_to_radians(angle_deg);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float dx = (float)x - cx, dy = (float)y - cy;
            float dist = sqrtf(dx * dx + dy * dy);
            float norm = dist / max_r;
            float rotation = twist_amount * (1.0f - norm);
            float cr = cosf(rotation), sr = sinf(rotation);
            float p[4];
            sample_bilinear(src, (int)w, (int)h, cx + dx * cr - dy * sr, cy + dx * sr + dy * cr, p);
            for (int c = 0; c < 4; c++) dst[oi + c] = round_u8(p[c]);
        }
}

/* hash_u32 / hash_f32, src/ops/effects.rs:143-161 */
static inline uint32_t hash_u32(uint32_t x) {
    x *= 0x9E3779B9u; x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}
static inline float hash_f32(uint32_t x, uint32_t y, uint32_t seed) {
    uint32_t hh = hash_u32(x * 374761393u + y * 668265263u + seed);
    return (float)(hh & 0x00FFFFFFu) / 16777216.0f;
}
/* perlin_noise_2d, noise.rs:52-71; turbulence_2d, distort.rs:229-246 */
static float perlin_noise_2d(float x, float y, uint32_t seed) {
    int xi = as_i32(floorf(x)), yi = as_i32(floorf(y));
    float xf = x - (float)xi, yf = y - (float)yi;
    float u = xf * xf * xf * (xf * (xf * 6.0f - 15.0f) + 10.0f), v = yf * yf * yf * (yf * (yf * 6.0f - 15.0f) + 10.0f);
    float n00 = hash_f32((uint32_t)xi, (uint32_t)yi, seed), n10 = hash_f32((uint32_t)(xi + 1), (uint32_t)yi, seed);
    float n01 = hash_f32((uint32_t)xi, (uint32_t)(yi + 1), seed), n11 = hash_f32((uint32_t)(xi + 1), (uint32_t)(yi + 1), seed);
    float nx0 = n00 + u * (n10 - n00), nx1 = n01 + u * (n11 - n01);
    return nx0 + v * (nx1 - nx0);
}
static float turbulence_2d(float x, float y, uint32_t seed, uint32_t octaves, float roughness) {
    float total = 0.0f, amplitude = 1.0f, frequency = 1.0f, max_amplitude = 0.0f;
    for (uint32_t i = 0; i < octaves; i++) {
        total += perlin_noise_2d(x * frequency, y * frequency, seed + i * 1000u) * amplitude;
        max_amplitude += amplitude;
        amplitude *= roughness;
        frequency *= 2.0f;
    }
    return max_amplitude > 0.0f ? total / max_amplitude : 0.0f;
}

/* add_noise_core, noise.rs:73-143. noise_type: 0 Uniform, 1 Gaussian, 2 Perlin */
void pfo_add_noise(const uint8_t *src, uint32_t w, uint32_t h, float amount, int noise_type, int monochrome,
                   uint32_t seed, float scale, uint32_t octaves, const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    float inv_scale = 1.0f / maxf(scale, 0.1f);
    uint32_t oct = octaves < 1 ? 1 : (octaves > 8 ? 8 : octaves);
    float strength = amount * 255.0f / 100.0f;
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float r = (float)src[oi], g = (float)src[oi + 1], b = (float)src[oi + 2], a = (float)src[oi + 3];
            float sx = (float)x * inv_scale, sy = (float)y * inv_scale;
            uint32_t qx = as_u32(floorf((float)x * inv_scale)), qy = as_u32(floorf((float)y * inv_scale));
            float nr, ng, nb;
            if (monochrome) {
                float nv;
                if (noise_type == 0) nv = hash_f32(qx, qy, seed) * 2.0f - 1.0f;
                else if (noise_type == 1) {
                    float u1 = maxf(hash_f32(qx, qy, seed), 0.0001f), u2 = hash_f32(qx, qy, seed + 7u);
                    nv = sqrtf(-2.0f * logf(u1)) * cosf(2.0f * 3.14159265358979323846f * u2) * 0.33f;
                } else nv = turbulence_2d(sx, sy, seed, oct, 0.5f) * 2.0f - 1.0f;
                nr = ng = nb = nv * strength;
            } else if (noise_type == 2) {
                nr = (turbulence_2d(sx, sy, seed, oct, 0.5f) * 2.0f - 1.0f) * strength;
                ng = (turbulence_2d(sx, sy, seed + 1u, oct, 0.5f) * 2.0f - 1.0f) * strength;
                nb = (turbulence_2d(sx, sy, seed + 2u, oct, 0.5f) * 2.0f - 1.0f) * strength;
            } else {  /* non-monochrome Uniform AND Gaussian both use the uniform hash per channel (:114-136) */
                nr = (hash_f32(qx, qy, seed) * 2.0f - 1.0f) * strength;
                ng = (hash_f32(qx, qy, seed + 1u) * 2.0f - 1.0f) * strength;
                nb = (hash_f32(qx, qy, seed + 2u) * 2.0f - 1.0f) * strength;
            }
            dst[oi] = round_u8(r + nr); dst[oi + 1] = round_u8(g + ng); dst[oi + 2] = round_u8(b + nb); dst[oi + 3] = round_u8(a);
        }
}

/* reduce_noise_core (bilateral), noise.rs:172-262 */
void pfo_reduce_noise(const uint8_t *src, uint32_t w, uint32_t h, float strength, uint32_t radius,
                      const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    int r = radius < 1 ? 1 : (int)radius;
    float sigma_s = (float)r, sigma_r = strength * 2.55f;
#pragma omp parallel for schedule(dynamic, 4)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float cr = (float)src[oi], cg = (float)src[oi + 1], cb = (float)src[oi + 2];
            float s[4] = {0, 0, 0, 0}, wsum = 0.0f;
            for (int dy = -r; dy <= r; dy++) {
                int sy = clampi((int)y + dy, 0, (int)h - 1);
                for (int dx = -r; dx <= r; dx++) {
                    int sx = clampi((int)x + dx, 0, (int)w - 1);
                    const uint8_t *p = src + ((size_t)sy * w + sx) * 4;
                    float pr = (float)p[0], pg = (float)p[1], pb = (float)p[2], pa = (float)p[3];
                    float spatial = (float)(dx * dx + dy * dy) / (2.0f * sigma_s * sigma_s);
                    float dr = cr - pr, dg = cg - pg, db = cb - pb;
                    float range = (dr * dr + dg * dg + db * db) / (2.0f * sigma_r * sigma_r + 0.001f);
                    float wt = expf(-spatial - range);
                    s[0] += pr * wt; s[1] += pg * wt; s[2] += pb * wt; s[3] += pa * wt;
                    wsum += wt;
                }
            }
            if (wsum > 0.0f) {
                float inv = 1.0f / wsum;
                for (int c = 0; c < 4; c++) dst[oi + c] = round_u8(s[c] * inv);
            } else memcpy(dst + oi, src + oi, 4);
        }
}

/* ========================================================================= */
/* Widened scope, part 2: the rest of src/ops/effects/ (every effect the      */
/* reference pins with a golden in tests/visual_filters.rs)                    */
/* ========================================================================= */

/* ink_core, src/ops/effects/artistic.rs:31-99 (Sobel on Rec.709 luminance) */
static inline float ink_lum(const uint8_t *src, int w, int h, int px, int py) {
    const uint8_t *s = src + ((size_t)clampi(py, 0, h - 1) * w + clampi(px, 0, w - 1)) * 4;
    return 0.2126f * (float)s[0] + 0.7152f * (float)s[1] + 0.0722f * (float)s[2];
}
void pfo_ink(const uint8_t *src, uint32_t w, uint32_t h, float edge_strength, float threshold,
             const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    int iw = (int)w, ih = (int)h;
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            int ix = (int)x, iy = (int)y;
#define L(a, b) ink_lum(src, iw, ih, (a), (b))
            float gx = -L(ix - 1, iy - 1) - 2.0f * L(ix - 1, iy) - L(ix - 1, iy + 1) + L(ix + 1, iy - 1) +
                       2.0f * L(ix + 1, iy) + L(ix + 1, iy + 1);
            float gy = -L(ix - 1, iy - 1) - 2.0f * L(ix, iy - 1) - L(ix + 1, iy - 1) + L(ix - 1, iy + 1) +
                       2.0f * L(ix, iy + 1) + L(ix + 1, iy + 1);
#undef L
            float edge = sqrtf(gx * gx + gy * gy) * edge_strength / 100.0f;
            uint8_t val = edge > threshold ? 0 : 255;
            dst[oi] = dst[oi + 1] = dst[oi + 2] = val;
            dst[oi + 3] = src[oi + 3];
        }
}

/* oil_painting_core, artistic.rs:123-217 (integer histogram of intensity bins) */
void pfo_oil_painting(const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius, uint32_t levels,
                      const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    int r = (int)(radius < 1 ? 1 : (radius > 10 ? 10 : radius));
    uint32_t nl = levels < 2 ? 2 : (levels > 64 ? 64 : levels);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++) {
        uint32_t cnt[64], sr[64], sg[64], sb[64];
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            for (uint32_t i = 0; i < nl; i++) cnt[i] = sr[i] = sg[i] = sb[i] = 0;
            for (int dy = -r; dy <= r; dy++) {
                int sy = clampi((int)y + dy, 0, (int)h - 1);
                for (int dx = -r; dx <= r; dx++) {
                    int sx = clampi((int)x + dx, 0, (int)w - 1);
                    const uint8_t *p = src + ((size_t)sy * w + sx) * 4;
                    uint32_t pr = p[0], pg = p[1], pb = p[2];
                    uint32_t bin = (pr + pg + pb) / 3 * nl / 256;
                    if (bin > nl - 1) bin = nl - 1;
                    cnt[bin]++; sr[bin] += pr; sg[bin] += pg; sb[bin] += pb;
                }
            }
            uint32_t max_count = 0, max_idx = 0;
            for (uint32_t i = 0; i < nl; i++)
                if (cnt[i] > max_count) { max_count = cnt[i]; max_idx = i; }
            dst[oi] = (uint8_t)(sr[max_idx] / max_count);
            dst[oi + 1] = (uint8_t)(sg[max_idx] / max_count);
            dst[oi + 2] = (uint8_t)(sb[max_idx] / max_count);
            dst[oi + 3] = src[oi + 3];
        }
    }
}

/* color_filter_core, artistic.rs:266-307. mode: 0 Multiply, 1 Screen, 2 Overlay, 3 SoftLight */
static inline float color_filter_blend(int mode, float s, float f) {
    switch (mode) {
    case 0: return s * f;
    case 1: return 1.0f - (1.0f - s) * (1.0f - f);
    case 2: return s < 0.5f ? 2.0f * s * f : 1.0f - 2.0f * (1.0f - s) * (1.0f - f);
    default: return f < 0.5f ? s - (1.0f - 2.0f * f) * s * (1.0f - s) : s + (2.0f * f - 1.0f) * (sqrtf(s) - s);
    }
}
void pfo_color_filter(const uint8_t *src, uint32_t w, uint32_t h, const uint8_t color[4], float intensity, int mode,
                      const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    float fc[3] = {(float)color[0] / 255.0f, (float)color[1] / 255.0f, (float)color[2] / 255.0f};
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)((size_t)w * h); i++) {
        size_t oi = (size_t)i * 4;
        if (mask && mask[i] == 0) { memcpy(dst + oi, src + oi, 4); continue; }
        for (int c = 0; c < 3; c++) {
            float s = (float)src[oi + c] / 255.0f;
            dst[oi + c] = round_u8((s * (1.0f - intensity) + color_filter_blend(mode, s, fc[c]) * intensity) * 255.0f);
        }
        dst[oi + 3] = src[oi + 3];
    }
}

/* contours_core, src/ops/effects/contours.rs:56-112 */
void pfo_contours(const uint8_t *src, uint32_t w, uint32_t h, float scale, float frequency, float line_width,
                  const uint8_t color[4], uint32_t seed, uint32_t octaves, float blend, const uint8_t *mask,
                  uint8_t *dst) {
    if (w == 0 || h == 0) return;
    float inv_scale = 1.0f / maxf(scale, 0.5f);
    uint32_t oct = octaves < 1 ? 1 : (octaves > 8 ? 8 : octaves);
    float half_lw = maxf(line_width * 0.5f, 0.3f);
    float lr = (float)color[0], lg = (float)color[1], lb = (float)color[2], la = (float)color[3] / 255.0f;
    float freq = maxf(frequency, 0.5f);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float r = (float)src[oi], g = (float)src[oi + 1], b = (float)src[oi + 2];
            float noise_val = turbulence_2d((float)x * inv_scale, (float)y * inv_scale, seed, oct, 0.5f);
            float level = noise_val * freq;
            float dist = fabsf(level - roundf(level)) / freq;
            float edge = half_lw * inv_scale * 0.5f;
            float line_alpha = dist < edge ? 1.0f : (dist < edge * 2.0f ? 1.0f - (dist - edge) / edge : 0.0f);
            float alpha = line_alpha * la * blend;
            dst[oi] = round_u8(r * (1.0f - alpha) + lr * alpha);
            dst[oi + 1] = round_u8(g * (1.0f - alpha) + lg * alpha);
            dst[oi + 2] = round_u8(b * (1.0f - alpha) + lb * alpha);
            dst[oi + 3] = src[oi + 3];
        }
}

/* crystallize_core, src/ops/effects/distort.rs:26-169 (jittered-grid Voronoi, f64 cell means) */
static inline size_t crystal_cell(const float *seeds, int cells_x, int cells_y, float cs, long x, long y) {
    int gcx = as_i32((float)x / cs), gcy = as_i32((float)y / cs);
    float px = (float)x + 0.5f, py = (float)y + 0.5f;
    float best = 3.40282347e+38f;
    size_t best_idx = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            int nx = gcx + dx, ny = gcy + dy;
            if (nx < 0 || ny < 0 || nx >= cells_x || ny >= cells_y) continue;
            size_t idx = (size_t)ny * cells_x + nx;
            float sx = seeds[idx * 2], sy = seeds[idx * 2 + 1];
            float d = (px - sx) * (px - sx) + (py - sy) * (py - sy);
            if (d < best) { best = d; best_idx = idx; }
        }
    return best_idx;
}
void pfo_crystallize(const uint8_t *src, uint32_t w, uint32_t h, float cell_size, uint32_t seed,
                     const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    float cs = maxf(cell_size, 2.0f);
    int cells_x = as_i32(ceilf((float)w / cs)), cells_y = as_i32(ceilf((float)h / cs));
    if (cells_x < 1) cells_x = 1;
    if (cells_y < 1) cells_y = 1;
    size_t nc = (size_t)cells_x * cells_y;
    float *seeds = (float *)malloc(nc * 2 * sizeof(float));
    double *sums = (double *)calloc(nc * 4, sizeof(double));
    uint32_t *counts = (uint32_t *)calloc(nc, sizeof(uint32_t));
    uint8_t *avg = (uint8_t *)calloc(nc, 4);
    for (int cy = 0; cy < cells_y; cy++)
        for (int cx = 0; cx < cells_x; cx++) {
            float jx = hash_f32((uint32_t)cx, (uint32_t)cy, seed), jy = hash_f32((uint32_t)cx, (uint32_t)cy, seed + 77u);
            seeds[((size_t)cy * cells_x + cx) * 2] = (float)cx * cs + jx * cs;
            seeds[((size_t)cy * cells_x + cx) * 2 + 1] = (float)cy * cs + jy * cs;
        }
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t b = crystal_cell(seeds, cells_x, cells_y, cs, x, y), si = ((size_t)y * w + x) * 4;
            for (int c = 0; c < 4; c++) sums[b * 4 + c] += (double)src[si + c];
            counts[b]++;
        }
    for (size_t i = 0; i < nc; i++)
        if (counts[i] > 0) {
            double inv = 1.0 / (double)counts[i];
            for (int c = 0; c < 4; c++) {
                double v = round(sums[i * 4 + c] * inv);
                avg[i * 4 + c] = (uint8_t)(v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v));
            }
        }
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            memcpy(dst + oi, avg + crystal_cell(seeds, cells_x, cells_y, cs, x, y) * 4, 4);
        }
    free(seeds); free(sums); free(counts); free(avg);
}

/* f32::rem_euclid */
static inline float rem_euclid_f(float a, float b) {
    float r = fmodf(a, b);
    return r < 0.0f ? r + fabsf(b) : r;
}
/* dents_core, distort.rs:248-310 */
void pfo_dents(const uint8_t *src, uint32_t w, uint32_t h, float scale, float amount, uint32_t seed, uint32_t octaves,
               float roughness, int pinch, int wrap, const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    uint32_t oct = octaves < 1 ? 1 : (octaves > 8 ? 8 : octaves);
    float inv_scale = 1.0f / maxf(scale, 0.5f);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float nx = turbulence_2d((float)x * inv_scale, (float)y * inv_scale, seed, oct, roughness) * 2.0f - 1.0f;
            float ny = turbulence_2d((float)x * inv_scale, (float)y * inv_scale, seed + 9999u, oct, roughness) * 2.0f - 1.0f;
            if (pinch) {
                float cx = (float)w * 0.5f, cy = (float)h * 0.5f;
                float dx = (float)x - cx, dy = (float)y - cy;
                float dist = maxf(sqrtf(dx * dx + dy * dy), 1.0f);
                float factor = (1.0f - dist / maxf(cx, cy)) * 0.5f;
                nx = nx + dx / dist * factor;
                ny = ny + dy / dist * factor;
            }
            float sx = (float)x + nx * amount * scale, sy = (float)y + ny * amount * scale;
            if (wrap) { sx = rem_euclid_f(sx, (float)w); sy = rem_euclid_f(sy, (float)h); }
            float p[4];
            sample_bilinear(src, (int)w, (int)h, sx, sy, p);
            for (int c = 0; c < 4; c++) dst[oi + c] = round_u8(p[c]);
        }
}

/* halftone_core, src/ops/effects/stylize.rs:242-277. shape: 0 Circle, 1 Square, 2 Diamond, 3 Line */
void pfo_halftone(const uint8_t *src, uint32_t w, uint32_t h, float dot_size, float angle_deg, int shape,
                  const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    float ds = maxf(dot_size, 2.0f);
    float angle = pfo_to_radians(angle_deg);
    float cos_a = cosf(angle), sin_a = sinf(angle);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float r = (float)src[oi], g = (float)src[oi + 1], b = (float)src[oi + 2];
            float lum = (0.2126f * r + 0.7152f * g + 0.0722f * b) / 255.0f;
            float fx = (float)x * cos_a + (float)y * sin_a;
            float fy = -((float)x) * sin_a + (float)y * cos_a;
            float qx = fx / ds, qy = fy / ds;
            float cx = fabsf(qx - truncf(qx)) - 0.5f, cy = fabsf(qy - truncf(qy)) - 0.5f;
            float thr;
            switch (shape) {
            case 0: thr = sqrtf(cx * cx + cy * cy) * 2.0f; break;
            case 1: thr = maxf(fabsf(cx), fabsf(cy)) * 2.0f; break;
            case 2: thr = fabsf(cx) + fabsf(cy); break;
            default: thr = fabsf(cy) * 2.0f; break;
            }
            uint8_t val = thr < lum ? 255 : 0;
            dst[oi] = dst[oi + 1] = dst[oi + 2] = val;
            dst[oi + 3] = src[oi + 3];
        }
}

/* bokeh_blur_core, src/ops/effects/blur.rs:22-115: equal-weight disc, integer sums (the reference's
 * sliding row sums equal the direct clamped-coordinate sum). Fills spans[] (<= 2r+1) for the caller. */
int pfo_bokeh_spans(float radius, int *spans, int cap, uint32_t *sample_count) {
    int r = as_i32(ceilf(radius)), n = 0;
    float r2 = radius * radius;
    uint32_t cnt = 0;
    for (int dy = -r; dy <= r; dy++) {
        float remaining = r2 - (float)(dy * dy);
        if (remaining >= 0.0f) {
            int span = as_i32(floorf(sqrtf(remaining)));
            if (n < cap) { spans[2 * n] = dy; spans[2 * n + 1] = span; }
            n++;
            cnt += (uint32_t)(span * 2 + 1);
        }
    }
    *sample_count = cnt;
    return n;
}
void pfo_bokeh_blur(const uint8_t *src, uint32_t w, uint32_t h, float radius, const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    if (radius < 0.5f) { memcpy(dst, src, (size_t)w * h * 4); return; }
    int r = as_i32(ceilf(radius));
    int *spans = (int *)malloc(sizeof(int) * 2 * (size_t)(2 * r + 1));
    uint32_t sample_count;
    int ns = pfo_bokeh_spans(radius, spans, 2 * r + 1, &sample_count);
    float inv_count = 1.0f / (float)sample_count;
#pragma omp parallel for schedule(dynamic, 4)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            uint64_t tot[4] = {0, 0, 0, 0};
            for (int s = 0; s < ns; s++) {
                int sy = clampi((int)y + spans[2 * s], 0, (int)h - 1), span = spans[2 * s + 1];
                for (int dx = -span; dx <= span; dx++) {
                    const uint8_t *p = src + ((size_t)sy * w + clampi((int)x + dx, 0, (int)w - 1)) * 4;
                    for (int c = 0; c < 4; c++) tot[c] += p[c];
                }
            }
            for (int c = 0; c < 4; c++) dst[oi + c] = round_u8((float)tot[c] * inv_count);
        }
    free(spans);
}

/* zoom_blur_core, blur.rs:322-427 */
void pfo_zoom_blur(const uint8_t *src, uint32_t w, uint32_t h, float center_x, float center_y, float strength,
                   uint32_t samples, const float tint[4], float tint_strength, const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    if (strength < 0.001f) { memcpy(dst, src, (size_t)w * h * 4); return; }
    float cx = center_x * (float)w, cy = center_y * (float)h;
    float s = clampf(strength, 0.0f, 0.99f);
    uint32_t n = samples < 2 ? 2 : samples;
    float inv_n = 1.0f / (float)n;
    float cdx[4] = {cx, (float)w - cx, cx, (float)w - cx}, cdy[4] = {cy, cy, (float)h - cy, (float)h - cy};
    float max_dist = 0.0f;
    for (int i = 0; i < 4; i++) max_dist = maxf(max_dist, sqrtf(cdx[i] * cdx[i] + cdy[i] * cdy[i]));
    max_dist = maxf(max_dist, 1.0f);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            float dx = (float)x - cx, dy = (float)y - cy;
            float sum[4] = {0, 0, 0, 0};
            for (uint32_t i = 0; i < n; i++) {
                float t = 1.0f - s * ((float)i / (float)(n - 1));
                int sx = clampi(as_i32(roundf(cx + dx * t)), 0, (int)w - 1);
                int sy = clampi(as_i32(roundf(cy + dy * t)), 0, (int)h - 1);
                const uint8_t *p = src + ((size_t)sy * w + sx) * 4;
                for (int c = 0; c < 4; c++) sum[c] += (float)p[c];
            }
            float v[4];
            for (int c = 0; c < 4; c++) v[c] = sum[c] * inv_n;
            if (tint_strength > 0.001f) {
                float dist = sqrtf(dx * dx + dy * dy);
                float t = maxf(1.0f - dist / max_dist, 0.0f) * tint_strength;
                for (int c = 0; c < 4; c++) v[c] = v[c] + (tint[c] * 255.0f - v[c]) * t;
            }
            for (int c = 0; c < 4; c++) dst[oi + c] = round_u8(v[c]);
        }
}

/* grid_core, src/ops/effects/render.rs:52-92. style: 0 Lines, 1 Checkerboard */
void pfo_grid(const uint8_t *src, uint32_t w, uint32_t h, uint32_t cell_w, uint32_t cell_h, uint32_t line_width,
              const uint8_t color[4], int style, float opacity, const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    uint32_t cw = cell_w < 2 ? 2 : cell_w, ch = cell_h < 2 ? 2 : cell_h, lw = line_width < 1 ? 1 : line_width;
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            int draw = style == 0 ? (((uint32_t)x % cw) < lw || ((uint32_t)y % ch) < lw)
                                  : ((((uint32_t)x / cw) + ((uint32_t)y / ch)) % 2 == 0);
            for (int c = 0; c < 4; c++) {
                float v = (float)src[oi + c];
                dst[oi + c] = round_u8(draw ? v * (1.0f - opacity) + (float)color[c] * opacity : v);
            }
        }
}

/* canvas_border_core, render.rs:114-165 */
void pfo_canvas_border(const uint8_t *src, uint32_t w, uint32_t h, uint32_t width, const uint8_t color[4],
                       const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    uint32_t bw = width < 1 ? 1 : width, m = w < h ? w : h;
    if (bw > m) bw = m;
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            int border = x < bw || y < bw || x >= w - bw || y >= h - bw;
            if (masked_out(mask, w, x, y) || !border) memcpy(dst + oi, src + oi, 4);
            else memcpy(dst + oi, color, 4);
        }
}

/* shadow_core (drop shadow), render.rs:220-352 */
void pfo_drop_shadow(const uint8_t *src, uint32_t w, uint32_t h, int32_t offset_x, int32_t offset_y, float blur_radius,
                     int widen_radius, const uint8_t color[4], float opacity, const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    size_t n = (size_t)w * h;
    uint8_t *sa = (uint8_t *)calloc(n, 1);
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            long sx = x - offset_x, sy = y - offset_y;
            if (sx >= 0 && sx < (long)w && sy >= 0 && sy < (long)h) sa[(size_t)y * w + x] = src[((size_t)sy * w + sx) * 4 + 3];
        }
    if (widen_radius) {
        int spread = as_i32(roundf(maxf(blur_radius, 1.0f)));
        if (spread > 0) {
            uint8_t *hz = (uint8_t *)calloc(n, 1);
            for (long y = 0; y < (long)h; y++)
                for (long x = 0; x < (long)w; x++) {
                    long x0 = x - spread < 0 ? 0 : x - spread, x1 = x + spread > (long)w - 1 ? (long)w - 1 : x + spread;
                    uint8_t m = 0;
                    for (long s = x0; s <= x1; s++) if (sa[(size_t)y * w + s] > m) m = sa[(size_t)y * w + s];
                    hz[(size_t)y * w + x] = m;
                }
            for (long y = 0; y < (long)h; y++) {
                long y0 = y - spread < 0 ? 0 : y - spread, y1 = y + spread > (long)h - 1 ? (long)h - 1 : y + spread;
                for (long x = 0; x < (long)w; x++) {
                    uint8_t m = 0;
                    for (long s = y0; s <= y1; s++) if (hz[(size_t)s * w + x] > m) m = hz[(size_t)s * w + x];
                    sa[(size_t)y * w + x] = m;
                }
            }
            free(hz);
        }
    }
    uint8_t *argba = (uint8_t *)malloc(n * 4), *blur = argba;
    for (size_t i = 0; i < n; i++) argba[i * 4] = argba[i * 4 + 1] = argba[i * 4 + 2] = argba[i * 4 + 3] = sa[i];
    if (blur_radius > 0.5f) {
        blur = (uint8_t *)malloc(n * 4);
        pfo_gaussian_blur(argba, w, h, blur_radius, blur);
    }
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        size_t si = (size_t)i * 4;
        if (mask && mask[i] == 0) { memcpy(dst + si, src + si, 4); continue; }
        float shadow_a = ((float)blur[si] / 255.0f) * opacity * ((float)color[3] / 255.0f);
        float src_a = (float)src[si + 3] / 255.0f;
        float out_a = src_a + shadow_a * (1.0f - src_a);
        for (int c = 0; c < 3; c++) {
            float shadow_c = (float)color[c] / 255.0f, src_c = (float)src[si + c] / 255.0f;
            float out_c = out_a > 0.0f ? (src_c * src_a + shadow_c * shadow_a * (1.0f - src_a)) / out_a : 0.0f;
            dst[si + c] = round_u8(out_c * 255.0f);
        }
        dst[si + 3] = round_u8(out_a * 255.0f);
    }
    if (blur != argba) free(blur);
    free(argba); free(sa);
}

/* outline_core, render.rs:403-572. mode: 0 Outside, 1 Inside, 2 Center */
static inline float outline_cov(float distance, float radius, int aa) {
    if (aa) {
        float t = clampf((radius + 0.5f - distance) / 1.0f, 0.0f, 1.0f);
        return t * t * (3.0f - 2.0f * t);
    }
    return distance <= radius ? 1.0f : 0.0f;
}
static inline int outline_nearest_sq(const uint8_t *src, int w, int h, int x, int y, int sr, int want_filled) {
    int best = -1;
    for (int dy = -sr; dy <= sr; dy++)
        for (int dx = -sr; dx <= sr; dx++) {
            int d = dx * dx + dy * dy;
            if (best >= 0 && d > best) continue;
            int sx = x + dx, sy = y + dy;
            if (sx < 0 || sy < 0 || sx >= w || sy >= h) continue;
            uint8_t a = src[((size_t)sy * w + sx) * 4 + 3];
            if (want_filled ? a > 0 : a == 0) best = d;
        }
    return best;
}
void pfo_outline(const uint8_t *src, uint32_t w, uint32_t h, uint32_t width, const uint8_t color[4], int mode,
                 int anti_alias, const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    memcpy(dst, src, (size_t)w * h * 4);
    float radius = (float)(width < 1 ? 1 : width);
    int sr = as_i32(ceilf(radius)) + 1;
    long minx = w, miny = h, maxx = 0, maxy = 0;
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++)
            if (src[((size_t)y * w + x) * 4 + 3] > 0) {
                if (x < minx) minx = x;
                if (y < miny) miny = y;
                if (x > maxx) maxx = x;
                if (y > maxy) maxy = y;
            }
    if (minx == (long)w) return;
    long pminx = minx - (sr + 1) < 0 ? 0 : minx - (sr + 1), pminy = miny - (sr + 1) < 0 ? 0 : miny - (sr + 1);
    long pmaxx = maxx + sr + 1 > (long)w - 1 ? (long)w - 1 : maxx + sr + 1;
    long pmaxy = maxy + sr + 1 > (long)h - 1 ? (long)h - 1 : maxy + sr + 1;
    float ca = (float)color[3] / 255.0f;
#pragma omp parallel for schedule(dynamic, 4)
    for (long y = pminy; y <= pmaxy; y++)
        for (long x = pminx; x <= pmaxx; x++) {
            if (masked_out(mask, w, (size_t)x, (size_t)y)) continue;
            size_t oi = ((size_t)y * w + x) * 4;
            float src_a = (float)src[oi + 3] / 255.0f;
            int dfo = outline_nearest_sq(src, (int)w, (int)h, (int)x, (int)y, sr, 1);
            int dfi = outline_nearest_sq(src, (int)w, (int)h, (int)x, (int)y, sr, 0);
            float outside_cov = (dfo >= 0 ? outline_cov(maxf(sqrtf((float)dfo) - 1.0f, 0.0f), radius, anti_alias) : 0.0f) * (1.0f - src_a);
            float inside_cov = (dfi >= 0 ? outline_cov(sqrtf((float)dfi), radius, anti_alias) : 0.0f) * src_a;
            float under_cov = mode == 1 ? 0.0f : outside_cov, over_cov = mode == 0 ? 0.0f : inside_cov;
            float a_under = ca * under_cov, a_over = ca * over_cov;
            float comp[3] = {(float)src[oi] / 255.0f, (float)src[oi + 1] / 255.0f, (float)src[oi + 2] / 255.0f};
            float comp_a = (float)src[oi + 3] / 255.0f;
            if (a_under > 0.0f) {
                float out_a = comp_a + a_under * (1.0f - comp_a);
                if (out_a > 0.0f)
                    for (int c = 0; c < 3; c++)
                        comp[c] = (comp[c] * comp_a + ((float)color[c] / 255.0f) * a_under * (1.0f - comp_a)) / out_a;
                comp_a = out_a;
            }
            if (a_over > 0.0f) {
                float out_a = a_over + comp_a * (1.0f - a_over);
                if (out_a > 0.0f)
                    for (int c = 0; c < 3; c++)
                        comp[c] = (((float)color[c] / 255.0f) * a_over + comp[c] * comp_a * (1.0f - a_over)) / out_a;
                comp_a = out_a;
            }
            for (int c = 0; c < 3; c++) dst[oi + c] = as_u8(roundf(clampf(comp[c], 0.0f, 1.0f) * 255.0f));
            dst[oi + 3] = as_u8(roundf(clampf(comp_a, 0.0f, 1.0f) * 255.0f));
        }
}

/* pixel_drag_core, src/ops/effects/glitch.rs:44-99 */
void pfo_pixel_drag(const uint8_t *src, uint32_t w, uint32_t h, uint32_t seed, float amount, uint32_t distance,
                    float direction, const uint8_t *mask, uint8_t *dst) {
    if (w == 0 || h == 0) return;
    memcpy(dst, src, (size_t)w * h * 4);
    float dir = pfo_to_radians(direction);
    float dx_dir = cosf(dir), dy_dir = sinf(dir);
    float dist = (float)(distance < 1 ? 1 : distance);
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++) {
        if (hash_f32((uint32_t)y, 0, seed) > amount / 100.0f) continue;
        int drag = as_i32(hash_f32((uint32_t)y, 1, seed) * dist);
        for (long x = 0; x < (long)w; x++) {
            if (masked_out(mask, w, (size_t)x, (size_t)y)) continue;
            int sx = clampi(as_i32(roundf((float)x - (float)drag * dx_dir)), 0, (int)w - 1);
            int sy = clampi(as_i32(roundf((float)y - (float)drag * dy_dir)), 0, (int)h - 1);
            memcpy(dst + ((size_t)y * w + x) * 4, src + ((size_t)sy * w + sx) * 4, 4);
        }
    }
}

/* rgb_displace_core, glitch.rs:142-197. off = {rx, ry, gx, gy, bx, by} */
void pfo_rgb_displace(const uint8_t *src, uint32_t w, uint32_t h, const int32_t off[6], const uint8_t *mask,
                      uint8_t *dst) {
    if (w == 0 || h == 0) return;
#pragma omp parallel for schedule(static)
    for (long y = 0; y < (long)h; y++)
        for (long x = 0; x < (long)w; x++) {
            size_t oi = ((size_t)y * w + x) * 4;
            if (masked_out(mask, w, (size_t)x, (size_t)y)) { memcpy(dst + oi, src + oi, 4); continue; }
            for (int c = 0; c < 3; c++) {
                long sx = x + off[2 * c], sy = y + off[2 * c + 1];
                sx = sx < 0 ? 0 : (sx > (long)w - 1 ? (long)w - 1 : sx);
                sy = sy < 0 ? 0 : (sy > (long)h - 1 ? (long)h - 1 : sy);
                dst[oi + c] = src[((size_t)sy * w + sx) * 4 + c];
            }
            dst[oi + 3] = src[oi + 3];
        }
}

/* ========================================================================= */
/* Geometry: flips / quarter turns, resize_canvas,         */
/* apply_affine, and imageops::resize                                          */
/* ========================================================================= */

/* imageops::flip_horizontal / flip_vertical / rotate90 / rotate270 / rotate180 as called from
 * src/ops/transform.rs:326-334 and src/ops/scripting.rs:645-740; identical to
 * TiledImage::{flip_*,rotate_*}_chunked (transform.rs:62-131) on the flattened layer.
 * op: 0 flip H, 1 flip V, 2 rotate 90 cw, 3 rotate 90 ccw, 4 rotate 180. dst is w*h or h*w. */
void pfo_orient(const uint8_t *src, uint32_t w, uint32_t h, int op, uint8_t *dst) {
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            size_t si = ((size_t)y * w + x) * 4, di;
            switch (op) {
            case 0: di = ((size_t)y * w + (w - 1 - x)) * 4; break;
            case 1: di = ((size_t)(h - 1 - y) * w + x) * 4; break;
            case 2: di = ((size_t)x * h + (h - 1 - y)) * 4; break;       /* out(h-1-y, x), width h */
            case 3: di = ((size_t)(w - 1 - x) * h + y) * 4; break;       /* out(y, w-1-x), width h */
            default: di = ((size_t)(h - 1 - y) * w + (w - 1 - x)) * 4; break;
            }
            memcpy(dst + di, src + si, 4);
        }
}

/* resize_canvas / resize_canvas_layers, transform.rs:382-463. anchor in {0,1,2}^2. */
void pfo_resize_canvas(const uint8_t *src, uint32_t ow, uint32_t oh, uint32_t nw, uint32_t nh, uint32_t ax, uint32_t ay,
                       const uint8_t fill[4], uint8_t *dst) {
    int32_t offx = ax == 0 ? 0 : (ax == 1 ? ((int32_t)nw - (int32_t)ow) / 2 : (int32_t)nw - (int32_t)ow);
    int32_t offy = ay == 0 ? 0 : (ay == 1 ? ((int32_t)nh - (int32_t)oh) / 2 : (int32_t)nh - (int32_t)oh);
    for (size_t i = 0; i < (size_t)nw * nh; i++) memcpy(dst + i * 4, fill, 4);
    for (uint32_t y = 0; y < oh; y++)
        for (uint32_t x = 0; x < ow; x++) {
            int32_t nx = (int32_t)x + offx, ny = (int32_t)y + offy;
            if (nx >= 0 && ny >= 0 && (uint32_t)nx < nw && (uint32_t)ny < nh)
                memcpy(dst + ((size_t)ny * nw + nx) * 4, src + ((size_t)y * ow + x) * 4, 4);
        }
}

/* invert_3x3, transform.rs:949-976 */
static void invert_3x3(const float m[3][3], float o[3][3]) {
    float a = m[0][0], b = m[0][1], c = m[0][2], d = m[1][0], e = m[1][1], f = m[1][2], g = m[2][0], h = m[2][1], i = m[2][2];
    float det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (fabsf(det) < 1e-12f) {
        float id[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        memcpy(o, id, sizeof(id));
        return;
    }
    float inv = 1.0f / det;
    o[0][0] = (e * i - f * h) * inv; o[0][1] = (c * h - b * i) * inv; o[0][2] = (b * f - c * e) * inv;
    o[1][0] = (f * g - d * i) * inv; o[1][1] = (a * i - c * g) * inv; o[1][2] = (c * d - a * f) * inv;
    o[2][0] = (d * h - e * g) * inv; o[2][1] = (b * g - a * h) * inv; o[2][2] = (a * e - b * d) * inv;
}
/* The inverse homography of apply_affine (transform.rs:838-876): out[0..8] = hi row-major,
 * out[9] = inv_scale. Host arithmetic (libm sinf/cosf), shared by the oracle and its callers. */
void pfo_affine_matrix(uint32_t canvas_w, uint32_t canvas_h, float rotation_z, float rotation_x, float rotation_y,
                       float scale, float out[10]) {
    float focal = (float)(canvas_w > canvas_h ? canvas_w : canvas_h) * 1.5f;
    float rz = pfo_to_radians(rotation_z), rx = pfo_to_radians(rotation_x), ry = pfo_to_radians(rotation_y);
    float sz = sinf(rz), cz = cosf(rz), sxr = sinf(rx), cxr = cosf(rx), syr = sinf(ry), cyr = cosf(ry);
    float r00 = cz * cyr, r01 = cz * syr * sxr - sz * cxr, r10 = sz * cyr, r11 = sz * syr * sxr + cz * cxr;
    float r20 = -syr, r21 = cyr * sxr;
    float hm[3][3] = {{focal * r00, focal * r01, 0.0f}, {focal * r10, focal * r11, 0.0f}, {r20, r21, focal}}, hi[3][3];
    invert_3x3(hm, hi);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) out[r * 3 + c] = hi[r][c];
    out[9] = fabsf(scale) > 1e-6f ? 1.0f / scale : 1.0f;
}
/* apply_affine, transform.rs:826-946. nearest != 0 selects Interpolation::Nearest; every other
 * interpolation is the bilinear branch. dst = canvas_w x canvas_h, transparent where unmapped. */
void pfo_affine(const uint8_t *src, uint32_t sw, uint32_t sh, uint32_t canvas_w, uint32_t canvas_h, float rotation_z,
                float rotation_x, float rotation_y, float scale, float off_x, float off_y, int nearest, uint8_t *dst) {
    memset(dst, 0, (size_t)canvas_w * canvas_h * 4);
    if (canvas_w == 0 || canvas_h == 0) return;
    float m[10];
    pfo_affine_matrix(canvas_w, canvas_h, rotation_z, rotation_x, rotation_y, scale, m);
    float cx = (float)canvas_w * 0.5f, cy = (float)canvas_h * 0.5f, inv_scale = m[9];
    int src_w = (int)sw, src_h = (int)sh;
#pragma omp parallel for schedule(static)
    for (long dy = 0; dy < (long)canvas_h; dy++) {
        float v = ((float)dy - cy - off_y) * inv_scale;
        float base_sx = m[1] * v + m[2], base_sy = m[4] * v + m[5], base_sw = m[7] * v + m[8];
        for (long dx = 0; dx < (long)canvas_w; dx++) {
            float u = ((float)dx - cx - off_x) * inv_scale;
            float wq = m[6] * u + base_sw;
            if (fabsf(wq) < 1e-8f) continue;
            float inv_w = 1.0f / wq;
            float src_x = (m[0] * u + base_sx) * inv_w + cx, src_y = (m[3] * u + base_sy) * inv_w + cy;
            uint8_t *o = dst + ((size_t)dy * canvas_w + dx) * 4;
            if (nearest) {
                int nx = as_i32(roundf(src_x)), ny = as_i32(roundf(src_y));
                if (nx >= 0 && ny >= 0 && nx < src_w && ny < src_h) memcpy(o, src + ((size_t)ny * sw + nx) * 4, 4);
                continue;
            }
            int x0 = as_i32(floorf(src_x)), y0 = as_i32(floorf(src_y));
            if (x0 < -1 || y0 < -1 || x0 >= src_w || y0 >= src_h) continue;
            float fx = src_x - (float)x0, fy = src_y - (float)y0;
            float t[4][4];
            for (int k = 0; k < 4; k++) {
                int sx = x0 + (k & 1), sy = y0 + (k >> 1);
                for (int c = 0; c < 4; c++)
                    t[k][c] = (sx < 0 || sy < 0 || sx >= src_w || sy >= src_h) ? 0.0f : (float)src[((size_t)sy * sw + sx) * 4 + c];
            }
            for (int c = 0; c < 4; c++) {
                float top = t[0][c] + (t[1][c] - t[0][c]) * fx, bot = t[2][c] + (t[3][c] - t[2][c]) * fx;
                o[c] = round_u8(top + (bot - top) * fy);
            }
        }
    }
}

/* imageops::resize of the `image` crate 0.25.9 (Cargo.lock; the crate is NOT vendored under
 * /root/reference, so this restates its published algorithm: imageops/sample.rs `resize` =
 * vertical_sample into an f32 image, then horizontal_sample with FloatNearest rounding and a clamp to
 * [0, 255]).  Called by resize_image / resize_layers (transform.rs:347-378).  Pinned by the reference's
 * goldens transforms/resize_{2x_nearest,half_bilinear,half_lanczos}.png; CatmullRom has no golden.
 * filter: 0 Nearest (box, support 0), 1 Triangle, 2 CatmullRom, 3 Lanczos3. */
static float rs_sinc(float t) {
    float a = t * 3.14159265358979323846f;
    return t == 0.0f ? 1.0f : sinf(a) / a;
}
static float rs_kernel(int filter, float x) {
    switch (filter) {
    case 0: return 1.0f; /* box_kernel over a window of one sample */
    case 1: return fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f;
    case 2: { /* bc_cubic_spline(x, b = 0, c = 0.5) */
        float a = fabsf(x), b = 0.0f, c = 0.5f, k;
        if (a < 1.0f) k = (12.0f - 9.0f * b - 6.0f * c) * a * a * a + (-18.0f + 12.0f * b + 6.0f * c) * a * a + (6.0f - 2.0f * b);
        else if (a < 2.0f) k = (-b - 6.0f * c) * a * a * a + (6.0f * b + 30.0f * c) * a * a + (-12.0f * b - 48.0f * c) * a + (8.0f * b + 24.0f * c);
        else k = 0.0f;
        return k / 6.0f;
    }
    default: return fabsf(x) < 3.0f ? rs_sinc(x) * rs_sinc(x / 3.0f) : 0.0f;
    }
}
static float rs_support(int filter) { return filter == 0 ? 0.0f : (filter == 1 ? 1.0f : (filter == 2 ? 2.0f : 3.0f)); }
/* One axis of sample weights: for every output index, `left[o]`, `count[o]` and count normalised
 * weights at weights[offset[o] ...]. Returns the total number of weights (call with weights == NULL
 * to size the buffer). */
size_t pfo_resize_weights(uint32_t n_in, uint32_t n_out, int filter, uint32_t *left_out, uint32_t *count_out,
                          uint32_t *offset_out, float *weights) {
    float ratio = (float)n_in / (float)n_out;
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float src_support = rs_support(filter) * sratio;
    size_t total = 0;
    for (uint32_t o = 0; o < n_out; o++) {
        float inp = ((float)o + 0.5f) * ratio;
        int64_t left = (int64_t)floorf(inp - src_support);
        if (left < 0) left = 0;
        if (left > (int64_t)n_in - 1) left = (int64_t)n_in - 1;
        int64_t right = (int64_t)ceilf(inp + src_support);
        if (right < left + 1) right = left + 1;
        if (right > (int64_t)n_in) right = (int64_t)n_in;
        inp = inp - 0.5f;
        uint32_t cnt = (uint32_t)(right - left);
        if (left_out) { left_out[o] = (uint32_t)left; count_out[o] = cnt; offset_out[o] = (uint32_t)total; }
        if (weights) {
            float sum = 0.0f;
            for (uint32_t i = 0; i < cnt; i++) {
                float wv = rs_kernel(filter, ((float)(left + i) - inp) / sratio);
                weights[total + i] = wv;
                sum += wv;
            }
            for (uint32_t i = 0; i < cnt; i++) weights[total + i] /= sum;
        }
        total += cnt;
    }
    return total;
}
void pfo_resize(const uint8_t *src, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, int filter, uint8_t *dst) {
    if (nw == 0 || nh == 0) return;
    if (w == 0 || h == 0) { memset(dst, 0, (size_t)nw * nh * 4); return; }
    if (nw == w && nh == h) { memcpy(dst, src, (size_t)w * h * 4); return; }
    uint32_t *vl = (uint32_t *)malloc(sizeof(uint32_t) * 3 * nh), *hl = (uint32_t *)malloc(sizeof(uint32_t) * 3 * nw);
    size_t nvw = pfo_resize_weights(h, nh, filter, NULL, NULL, NULL, NULL), nhw = pfo_resize_weights(w, nw, filter, NULL, NULL, NULL, NULL);
    float *vw = (float *)malloc(sizeof(float) * nvw), *hw = (float *)malloc(sizeof(float) * nhw);
    pfo_resize_weights(h, nh, filter, vl, vl + nh, vl + 2 * nh, vw);
    pfo_resize_weights(w, nw, filter, hl, hl + nw, hl + 2 * nw, hw);
    float *tmp = (float *)malloc(sizeof(float) * 4 * (size_t)w * nh);
#pragma omp parallel for schedule(static)
    for (long oy = 0; oy < (long)nh; oy++)
        for (uint32_t x = 0; x < w; x++) {
            float t[4] = {0, 0, 0, 0};
            for (uint32_t i = 0; i < vl[nh + oy]; i++) {
                const uint8_t *p = src + ((size_t)(vl[oy] + i) * w + x) * 4;
                float wt = vw[vl[2 * nh + oy] + i];
                for (int c = 0; c < 4; c++) t[c] += (float)p[c] * wt;
            }
            memcpy(tmp + ((size_t)oy * w + x) * 4, t, sizeof(t));
        }
#pragma omp parallel for schedule(static)
    for (long oy = 0; oy < (long)nh; oy++)
        for (uint32_t ox = 0; ox < nw; ox++) {
            float t[4] = {0, 0, 0, 0};
            for (uint32_t i = 0; i < hl[nw + ox]; i++) {
                const float *p = tmp + ((size_t)oy * w + hl[ox] + i) * 4;
                float wt = hw[hl[2 * nw + ox] + i];
                for (int c = 0; c < 4; c++) t[c] += p[c] * wt;
            }
            for (int c = 0; c < 4; c++) dst[((size_t)oy * nw + ox) * 4 + c] = as_u8(roundf(clampf(t[c], 0.0f, 255.0f)));
        }
    free(tmp); free(vw); free(hw); free(vl); free(hl);
}
