// Brush-stamp inner loop: a whole stroke's stamps in one launch.
// Reference: ToolsPanel::draw_circle_no_dirty (src/ui/panels/tools/behavior/raster/brush_render.rs:
// 135-400) for the circle tip in every BrushMode and the eraser; rebuild_brush_lut (:27-50);
// compute_brush_alpha (:54-82); draw_line_no_dirty's stamp placement (:762-838).
//
// The reference stamps sequentially, one circle per pixel of stroke length, each a read-modify-
// write of the target tiles. Here one thread owns one pixel of the stroke's bounding box and walks
// the stamp list in order, so the per-pixel sequence of compares and writes is the reference's,
// the pixel is read once and written once, and overdraw costs registers instead of memory traffic.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "hsl.cuh"

namespace {

struct BrushParams {
    float radius, radius_sq, draw_radius_sq, inv_radius_sq;
    float hardness;  // clamped to [0,1]
    float src_a, flow;
    int anti_aliased, direct, is_eraser, mode;
    uint32_t rgb;  // packed r8 | g8<<8 | b8<<16
    uint8_t lut[256];
};

// compute_brush_alpha, brush_render.rs:54-82
__host__ __device__ inline float brush_alpha(float dist, float radius, float hardness, int aa) {
    if (radius <= 0.0f) return 0.0f;
    float t = dist / radius;
    t = t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t);
    float falloff = t * t * (3.0f - 2.0f * t);
    float material = 1.0f + (hardness - 1.0f) * falloff;
    float coverage;
    if (aa) {
        float e0 = radius + 0.5f, e1 = radius - 0.5f;
        if (dist <= e1) coverage = 1.0f;
        else if (dist >= e0) coverage = 0.0f;
        else {
            float x = (dist - e0) / (e1 - e0);
            x = x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x);
            coverage = x * x * (3.0f - 2.0f * x);
        }
    } else {
        coverage = dist <= radius ? 1.0f : 0.0f;
    }
    return material * coverage;
}

// One stamp on one pixel: the body of the reference's inner loop (:319-396).
__device__ __forceinline__ uint32_t stamp_pixel(const BrushParams &B, const uint8_t *lut, uint32_t px, float fgx, float fgy,
                                                float2 c, float draw_radius) {
    // the stamp's own bounding box (:205-211); pixels outside it are never visited
    if (fgx < fmaxf(floorf(c.x - draw_radius), 0.0f) || fgx > ceilf(c.x + draw_radius) ||
        fgy < fmaxf(floorf(c.y - draw_radius), 0.0f) || fgy > ceilf(c.y + draw_radius))
        return px;
    float dy = fgy - c.y, dx = fgx - c.x;
    float dist_sq = dx * dx + dy * dy;
    if (dist_sq > B.draw_radius_sq) return px;                                // :321
    uint32_t ga8;
    if (B.direct) {                                                           // :325-332
        float a = brush_alpha(sqrtf(dist_sq), B.radius, B.hardness, B.anti_aliased);
        ga8 = (uint32_t)__float2int_rz(fminf(fmaxf(fminf(roundf(a * 255.0f), 255.0f), 0.0f), 255.0f));
    } else {                                                                  // :334-336
        float fi = fminf(dist_sq * B.inv_radius_sq * 255.0f, 255.0f);
        ga8 = lut[(uint32_t)__float2int_rz(fmaxf(fi, 0.0f))];
    }
    if (ga8 == 0) return px;
    float strength = (float)ga8 / 255.0f * B.src_a * B.flow;                  // :340, :347, :361
    if (strength < 0.01f) return px;
    if (B.is_eraser) {
        float old_mask = (float)(px >> 24) / 255.0f;
        if (strength > old_mask) px = pfe_as_u8(strength * 255.0f) << 24;     // :352-357
    } else if (B.mode == 0) {
        uint32_t a8 = pfe_as_u8(strength * 255.0f);
        if (a8 >= (px >> 24)) px = B.rgb | (a8 << 24);                        // :366-372
    } else {                                                                  // Dodge / Burn / Sponge :374-394
        float hh, sat, l, nr, ng, nb;
        const float st = strength * 0.5f;
        rgb_to_hsl((float)(px & 255u) / 255.0f, (float)((px >> 8) & 255u) / 255.0f, (float)((px >> 16) & 255u) / 255.0f, hh, sat, l);
        if (B.mode == 1) l = pfe_clampf(l + st, 0.0f, 1.0f);
        else if (B.mode == 2) l = pfe_clampf(l - st, 0.0f, 1.0f);
        else if (B.mode == 3) sat = pfe_clampf(sat - st, 0.0f, 1.0f);
        hsl_to_rgb(hh, sat, l, 1e-6f, nr, ng, nb);
        px = pfe_pack(pfe_as_u8(nr * 255.0f), pfe_as_u8(ng * 255.0f), pfe_as_u8(nb * 255.0f), px >> 24);
    }
    return px;
}

// A CTA owns a 32 x 8 pixel tile of the stroke's bounding box.  The stamp list is read 256 stamps at a time: every
// thread tests one stamp's bounding box against the TILE, the survivors are compacted IN ORDER into shared memory
// (warp ballots + an 8-entry prefix), and each pixel then walks just those - for a long stroke a tile meets a few dozen
// of the thousands of stamps, and most tiles of the bounding box meet none and cost eight box tests per thread.  The
// per-pixel sequence of stamps is unchanged, so the result is the one-thread-per-pixel-over-every-stamp result.
__global__ void __launch_bounds__(256) brush_kernel(const __grid_constant__ BrushParams B, uint32_t *img, uint32_t w,
                                                    const float2 *centres, uint32_t n, const uint8_t *sel, int bx0,
                                                    int by0, int bw, int bh, float draw_radius) {
    __shared__ uint8_t lut[256];
    __shared__ float2 near_c[256];
    __shared__ uint32_t warp_count[8];
    lut[threadIdx.x] = B.lut[threadIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lx = blockIdx.x * 32 + lane, ly = blockIdx.y * 8 + warp;
    const int gx = bx0 + lx, gy = by0 + ly;
    const size_t o = (size_t)gy * w + gx;
    bool active = lx < bw && ly < bh;
    if (active && sel && sel[o] == 0) active = false;                         // :309-317
    uint32_t px = active ? img[o] : 0u;
    const uint32_t before = px;
    const float fgx = (float)gx, fgy = (float)gy;
    // tile rectangle in image coordinates (as floats: the per-stamp test below is the per-pixel test's own arithmetic)
    const float tx0 = (float)(bx0 + (int)blockIdx.x * 32), tx1 = tx0 + 31.0f;
    const float ty0 = (float)(by0 + (int)blockIdx.y * 8), ty1 = ty0 + 7.0f;
    for (uint32_t base = 0; base < n; base += 256) {
        const uint32_t s = base + threadIdx.x;
        float2 c = make_float2(0.f, 0.f);
        bool keep = false;
        if (s < n) {
            c = __ldg(centres + s);
            // some pixel of the tile lies inside the stamp's bounding box (:205-211)
            keep = !(tx1 < fmaxf(floorf(c.x - draw_radius), 0.0f) || tx0 > ceilf(c.x + draw_radius) ||
                     ty1 < fmaxf(floorf(c.y - draw_radius), 0.0f) || ty0 > ceilf(c.y + draw_radius));
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        __syncthreads();  // the previous chunk's list has been consumed (and, first time round, lut is complete)
        if (lane == 0) warp_count[warp] = __popc(ballot);
        __syncthreads();
        uint32_t offset = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const uint32_t cnt = warp_count[k];
            offset += k < warp ? cnt : 0u;
            total += cnt;
        }
        if (total == 0) continue;  // CTA-uniform
        if (keep) near_c[offset + __popc(ballot & ((1u << lane) - 1u))] = c;
        __syncthreads();
        if (active)
            for (uint32_t k = 0; k < total; k++) px = stamp_pixel(B, lut, px, fgx, fgy, near_c[k], draw_radius);
    }
    if (active && px != before) img[o] = px;
}

void fill_params(const pfe_brush_desc *b, BrushParams *P) {
    memset(P, 0, sizeof(*P));
    const float radius = b->size / 2.0f;
    P->radius = radius;
    P->radius_sq = radius * radius;
    const float draw_radius = b->anti_aliased ? radius + 0.5f : radius;
    P->draw_radius_sq = draw_radius * draw_radius;
    P->direct = draw_radius > radius;
    P->inv_radius_sq = 1.0f / P->radius_sq;
    P->hardness = std::min(std::max(b->hardness, 0.0f), 1.0f);
    P->src_a = b->color[3];
    P->flow = b->flow;
    P->anti_aliased = b->anti_aliased ? 1 : 0;
    P->is_eraser = b->is_eraser ? 1 : 0;
    P->mode = (b->mode >= 0 && b->mode <= 3) ? b->mode : 0;  // unknown modes leave the pixel as Normal would
    auto as_u8 = [](float v) -> uint32_t { return v != v || v <= 0.0f ? 0u : (v >= 255.0f ? 255u : (uint32_t)v); };
    P->rgb = as_u8(b->color[0] * 255.0f) | (as_u8(b->color[1] * 255.0f) << 8) | (as_u8(b->color[2] * 255.0f) << 16);
    pfe_brush_lut(b, P->lut);
}

}  // namespace

// rebuild_brush_lut, brush_render.rs:27-50 (host arithmetic)
extern "C" void pfe_brush_lut(const pfe_brush_desc *b, uint8_t lut[256]) {
    const float radius = b->size / 2.0f;
    if (radius < 0.001f) { memset(lut, 0, 256); return; }
    const float hardness = std::min(std::max(b->hardness, 0.0f), 1.0f);
    for (int i = 0; i < 256; i++) {
        float t_sq = (float)i / 255.0f;
        float dist = sqrtf(t_sq) * radius;
        float alpha = brush_alpha(dist, radius, hardness, b->anti_aliased ? 1 : 0);
        float v = fminf(roundf(alpha * 255.0f), 255.0f);
        lut[i] = (uint8_t)(v != v || v <= 0.0f ? 0 : (int)v);
    }
}

// draw_line_no_dirty, brush_render.rs:762-838 (circle tip: step = 1 px)
extern "C" int pfe_brush_line_centres(uint32_t w, uint32_t h, float x0, float y0, float x1, float y1, float *c, int cap) {
    if (!c || cap <= 0) return 0;
    auto as_u32 = [](float v) -> uint32_t { return v != v || v <= 0.0f ? 0u : (v >= 4294967296.0f ? 0xFFFFFFFFu : (uint32_t)v); };
    const float dx = x1 - x0, dy = y1 - y0;
    const float distance = sqrtf(dx * dx + dy * dy);
    int n = 0;
    if (distance < 0.1f) {
        if (x0 >= 0.0f && as_u32(x0) < w && y0 >= 0.0f && as_u32(y0) < h) { c[0] = x0; c[1] = y0; n = 1; }
        return n;
    }
    const uint32_t steps = as_u32(ceilf(distance / 1.0f));
    for (uint32_t i = 0; i <= steps && n < cap; i++) {
        float t = (float)i / (float)steps;
        float x = x0 + dx * t, y = y0 + dy * t;
        if (x >= 0.0f && as_u32(x) < w && y >= 0.0f && as_u32(y) < h) { c[n * 2] = x; c[n * 2 + 1] = y; n++; }
    }
    return n;
}

extern "C" int pfe_dev_brush_stamps(pfe_ctx *ctx, uint8_t *image, uint32_t w, uint32_t h, const pfe_brush_desc *brush,
                                    const float *centres, uint32_t n, const uint8_t *sel) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!image || !brush || (n && !centres) || !w || !h) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "brush: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return PFE_OK;
    BrushParams P;
    fill_params(brush, &P);
    if (P.radius_sq < 0.001f) return PFE_OK;                                  // :196-198
    const float draw_radius = brush->anti_aliased ? P.radius + 0.5f : P.radius;
    // union of the stamps' bounding boxes (:205-211)
    float fx0 = 3.0e38f, fy0 = 3.0e38f, fx1 = -3.0e38f, fy1 = -3.0e38f;
    for (uint32_t i = 0; i < n; i++) {
        fx0 = std::min(fx0, centres[i * 2]); fx1 = std::max(fx1, centres[i * 2]);
        fy0 = std::min(fy0, centres[i * 2 + 1]); fy1 = std::max(fy1, centres[i * 2 + 1]);
    }
    auto clampi = [](double v, double lo, double hi) { return (int)std::min(std::max(v, lo), hi); };
    const int bx0 = clampi(floor((double)fx0 - draw_radius), 0, (double)w - 1), by0 = clampi(floor((double)fy0 - draw_radius), 0, (double)h - 1);
    const int bx1 = clampi(ceil((double)fx1 + draw_radius), 0, (double)w - 1), by1 = clampi(ceil((double)fy1 + draw_radius), 0, (double)h - 1);
    if (bx1 < bx0 || by1 < by0) return PFE_OK;
    const size_t cbytes = (size_t)n * 8;
    void *cdev;
    if (cbytes <= PFE_SMALL_BYTES / 2) {
        PFE_TRY(pfe_small_upload(ctx, centres, cbytes, &cdev));
    } else {
        PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, cbytes, &cdev));
        PFE_CUDA(ctx, cudaMemcpyAsync(cdev, centres, cbytes, cudaMemcpyHostToDevice, ctx->stream));
        PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // centres is caller memory
    }
    const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
    PFE_KERNEL(ctx, "brush", brush_kernel<<<dim3(pfe_div_up(bw, 32), pfe_div_up(bh, 8)), 256, 0, ctx->stream>>>(
        P, (uint32_t *)image, w, (const float2 *)cdev, n, sel, bx0, by0, bw, bh, draw_radius));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}
