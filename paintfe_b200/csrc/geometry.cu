// Geometry: flips and quarter turns (src/ops/transform.rs:62-131, :326-334;
// scripting.rs:645-740), resize_canvas (transform.rs:382-463), apply_affine (transform.rs:826-946) and
// imageops::resize as resize_image / resize_layers call it (transform.rs:347-378).
//
// imageops::resize is the `image` crate 0.25.9 (Cargo.lock), not vendored under the reference: its
// published algorithm (imageops/sample.rs: vertical_sample to an f32 image, then horizontal_sample with a
// clamp and FloatNearest rounding) is restated here; the sample weights are computed on the host with
// libm exactly as the crate does, the two weighted sums run on the device in strict f32, in tap order.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

// ---- flips / quarter turns -----------------------------------------------------------------------
// op: 0 flip H, 1 flip V, 4 rotate 180 — same shape, one coalesced read and write per pixel
__global__ void __launch_bounds__(256) mirror_kernel(const uint32_t *src, uint32_t *dst, int w, int h, int op) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    const int sx = (op == 1) ? x : w - 1 - x, sy = (op == 0) ? y : h - 1 - y;
    dst[(size_t)y * w + x] = src[(size_t)sy * w + sx];
}
// op 2: rotate90 (out(h-1-y, x) = in(x, y)); op 3: rotate270 (out(y, w-1-x) = in(x, y)); out is h wide, w tall.
// A 32x32 tile goes through shared memory so both the read and the write are coalesced rows.
__global__ void __launch_bounds__(256) quarter_turn_kernel(const uint32_t *src, uint32_t *dst, int w, int h, int op) {
    __shared__ uint32_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    for (int j = ty; j < 32; j += 8) {
        const int x = x0 + tx, y = y0 + j;
        if (x < w && y < h) tile[j][tx] = src[(size_t)y * w + x];
    }
    __syncthreads();
    // output rows of this tile: one per source column; output columns: one per source row
    for (int j = ty; j < 32; j += 8) {
        // thread (tx, j) writes output pixel whose source is (x0 + j, y0 + k) with k chosen so that
        // consecutive tx are consecutive output columns
        const int sx = x0 + j;
        const int k = (op == 2) ? 31 - tx : tx;
        const int sy = y0 + k;
        if (sx < w && sy < h) {
            const int ox = (op == 2) ? h - 1 - sy : sy, oy = (op == 2) ? sx : w - 1 - sx;
            dst[(size_t)oy * h + ox] = tile[k][j];
        }
    }
}

// ---- resize_canvas -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_canvas_kernel(const uint32_t *src, uint32_t *dst, int ow, int oh, int nw, int nh,
                                                            int offx, int offy, uint32_t fill) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= nw || y >= nh) return;
    const long long sx = (long long)x - offx, sy = (long long)y - offy;
    dst[(size_t)y * nw + x] = (sx >= 0 && sy >= 0 && sx < ow && sy < oh) ? src[(size_t)sy * ow + sx] : fill;
}

// ---- apply_affine --------------------------------------------------------------------------------
struct AffineParams { float m[9], inv_scale, cx, cy, off_x, off_y; int src_w, src_h, nearest; };
__device__ __forceinline__ int sat_i32(float v) {
    return (v != v) ? 0 : __float2int_rz(fminf(fmaxf(v, -2147483648.0f), 2147483520.0f));
}
__global__ void __launch_bounds__(256) affine_kernel(const uint32_t *src, uint32_t *dst, int cw, int chh, AffineParams P) {
    const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (dx >= cw || dy >= chh) return;
    uint32_t out = 0u;  // RgbaImage::new: transparent where nothing maps
    const float v = ((float)dy - P.cy - P.off_y) * P.inv_scale;
    const float base_sx = P.m[1] * v + P.m[2], base_sy = P.m[4] * v + P.m[5], base_sw = P.m[7] * v + P.m[8];
    const float u = ((float)dx - P.cx - P.off_x) * P.inv_scale;
    const float wq = P.m[6] * u + base_sw;
    if (!(fabsf(wq) < 1e-8f)) {
        const float inv_w = 1.0f / wq;
        const float src_x = (P.m[0] * u + base_sx) * inv_w + P.cx, src_y = (P.m[3] * u + base_sy) * inv_w + P.cy;
        if (P.nearest) {
            const int nx = sat_i32(roundf(src_x)), ny = sat_i32(roundf(src_y));
            if (nx >= 0 && ny >= 0 && nx < P.src_w && ny < P.src_h) out = __ldg(src + (size_t)ny * P.src_w + nx);
        } else {
            const int x0 = sat_i32(floorf(src_x)), y0 = sat_i32(floorf(src_y));
            if (!(x0 < -1 || y0 < -1 || x0 >= P.src_w || y0 >= P.src_h)) {
                const float fx = src_x - (float)x0, fy = src_y - (float)y0;
                uint32_t t[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int sx = x0 + (k & 1), sy = y0 + (k >> 1);
                    t[k] = (sx < 0 || sy < 0 || sx >= P.src_w || sy >= P.src_h) ? 0u : __ldg(src + (size_t)sy * P.src_w + sx);
                }
                uint32_t o[4];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float tl = (float)((t[0] >> (8 * c)) & 255u), tr = (float)((t[1] >> (8 * c)) & 255u);
                    const float bl = (float)((t[2] >> (8 * c)) & 255u), br = (float)((t[3] >> (8 * c)) & 255u);
                    const float top = tl + (tr - tl) * fx, bot = bl + (br - bl) * fx;
                    o[c] = pfe_round_u8(top + (bot - top) * fy);
                }
                out = pfe_pack(o[0], o[1], o[2], o[3]);
            }
        }
    }
    dst[(size_t)dy * cw + dx] = out;
}

// ---- imageops::resize ----------------------------------------------------------------------------
struct AxisTable { const uint32_t *left, *count, *offset; const float *weights; };
__global__ void __launch_bounds__(256) resize_v_kernel(const uint32_t *src, float4 *tmp, int w, int nh, AxisTable T) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || oy >= nh) return;
    const uint32_t left = __ldg(T.left + oy), cnt = __ldg(T.count + oy);
    const float *wt = T.weights + __ldg(T.offset + oy);
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (uint32_t i = 0; i < cnt; i++) {
        const uint32_t p = __ldg(src + (size_t)(left + i) * w + x);
        const float k = __ldg(wt + i);
        t0 += (float)(p & 255u) * k; t1 += (float)((p >> 8) & 255u) * k;
        t2 += (float)((p >> 16) & 255u) * k; t3 += (float)(p >> 24) * k;
    }
    tmp[(size_t)oy * w + x] = make_float4(t0, t1, t2, t3);
}
__global__ void __launch_bounds__(256) resize_h_kernel(const float4 *tmp, uint32_t *dst, int w, int nw, int nh, AxisTable T) {
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ox >= nw || oy >= nh) return;
    const uint32_t left = __ldg(T.left + ox), cnt = __ldg(T.count + ox);
    const float *wt = T.weights + __ldg(T.offset + ox);
    const float4 *row = tmp + (size_t)oy * w + left;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    for (uint32_t i = 0; i < cnt; i++) {
        const float4 p = __ldg(row + i);
        const float k = __ldg(wt + i);
        t0 += p.x * k; t1 += p.y * k; t2 += p.z * k; t3 += p.w * k;
    }
    // clamp(t, 0, max) then FloatNearest (round half away from zero)
    dst[(size_t)oy * nw + ox] = pfe_pack(pfe_as_u8(roundf(pfe_clampf(t0, 0.0f, 255.0f))), pfe_as_u8(roundf(pfe_clampf(t1, 0.0f, 255.0f))),
                                         pfe_as_u8(roundf(pfe_clampf(t2, 0.0f, 255.0f))), pfe_as_u8(roundf(pfe_clampf(t3, 0.0f, 255.0f))));
}

// sample kernels of imageops/sample.rs
float rs_sinc(float t) {
    const float a = t * 3.14159265358979323846f;
    return t == 0.0f ? 1.0f : sinf(a) / a;
}
float rs_kernel(int filter, float x) {
    switch (filter) {
    case PFE_RESIZE_NEAREST: return 1.0f;
    case PFE_RESIZE_TRIANGLE: return fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f;
    case PFE_RESIZE_CATMULL_ROM: {
        const float a = fabsf(x), b = 0.0f, c = 0.5f;
        float k;
        if (a < 1.0f) k = (12.0f - 9.0f * b - 6.0f * c) * a * a * a + (-18.0f + 12.0f * b + 6.0f * c) * a * a + (6.0f - 2.0f * b);
        else if (a < 2.0f) k = (-b - 6.0f * c) * a * a * a + (6.0f * b + 30.0f * c) * a * a + (-12.0f * b - 48.0f * c) * a + (8.0f * b + 24.0f * c);
        else k = 0.0f;
        return k / 6.0f;
    }
    default: return fabsf(x) < 3.0f ? rs_sinc(x) * rs_sinc(x / 3.0f) : 0.0f;
    }
}
// One axis: table = [left | count | offset] (3 * n_out u32) followed by the weights.
void axis_weights(uint32_t n_in, uint32_t n_out, int filter, std::vector<uint32_t> &table, std::vector<float> &weights) {
    const float support = filter == PFE_RESIZE_NEAREST ? 0.0f : (filter == PFE_RESIZE_TRIANGLE ? 1.0f : (filter == PFE_RESIZE_CATMULL_ROM ? 2.0f : 3.0f));
    const float ratio = (float)n_in / (float)n_out;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float src_support = support * sratio;
    table.assign((size_t)n_out * 3, 0u);
    weights.clear();
    for (uint32_t o = 0; o < n_out; o++) {
        float inp = ((float)o + 0.5f) * ratio;
        int64_t left = (int64_t)floorf(inp - src_support);
        left = std::min<int64_t>(std::max<int64_t>(left, 0), (int64_t)n_in - 1);
        int64_t right = (int64_t)ceilf(inp + src_support);
        right = std::min<int64_t>(std::max<int64_t>(right, left + 1), (int64_t)n_in);
        inp = inp - 0.5f;
        const uint32_t cnt = (uint32_t)(right - left);
        table[o] = (uint32_t)left; table[n_out + o] = cnt; table[2 * (size_t)n_out + o] = (uint32_t)weights.size();
        const size_t base = weights.size();
        float sum = 0.0f;
        for (uint32_t i = 0; i < cnt; i++) {
            const float wv = rs_kernel(filter, ((float)(left + i) - inp) / sratio);
            weights.push_back(wv);
            sum += wv;
        }
        for (uint32_t i = 0; i < cnt; i++) weights[base + i] /= sum;
    }
}

inline dim3 grid2d(uint32_t w, uint32_t h) { return dim3(pfe_div_up(w, 32), pfe_div_up(h, 8)); }
int check2(pfe_ctx *ctx, const void *src, const void *dst, uint32_t w, uint32_t h, uint32_t nw, uint32_t nh, const char *what) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    const uint32_t lim = 0x7FFFFFFFu / 4;
    if (!src || !dst || !w || !h || !nw || !nh || src == dst || w > lim || h > lim || nw > lim || nh > lim)
        return pfe_fail(ctx, PFE_ERR_INVALID_ARG, what);
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    return PFE_OK;
}
inline float to_radians(float deg) { return deg * (3.14159265358979323846f / 180.0f); }
// invert_3x3, transform.rs:949-976
void invert_3x3(const float m[3][3], float o[3][3]) {
    const float a = m[0][0], b = m[0][1], c = m[0][2], d = m[1][0], e = m[1][1], f = m[1][2], g = m[2][0], h = m[2][1], i = m[2][2];
    const float det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (fabsf(det) < 1e-12f) {
        const float id[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        memcpy(o, id, sizeof(id));
        return;
    }
    const float inv = 1.0f / det;
    o[0][0] = (e * i - f * h) * inv; o[0][1] = (c * h - b * i) * inv; o[0][2] = (b * f - c * e) * inv;
    o[1][0] = (f * g - d * i) * inv; o[1][1] = (a * i - c * g) * inv; o[1][2] = (c * d - a * f) * inv;
    o[2][0] = (d * h - e * g) * inv; o[2][1] = (b * g - a * h) * inv; o[2][2] = (a * e - b * d) * inv;
}

}  // namespace

extern "C" int pfe_dev_orient(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int op, uint8_t *dst) {
    PFE_TRY(check2(ctx, src, dst, w, h, w, h, "orient: bad args"));
    if (op < PFE_ORIENT_FLIP_H || op > PFE_ORIENT_ROTATE_180) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "orient: unknown op");
    if (op == PFE_ORIENT_ROTATE_90CW || op == PFE_ORIENT_ROTATE_90CCW)
        PFE_KERNEL(ctx, "quarter_turn", quarter_turn_kernel<<<dim3(pfe_div_up(w, 32), pfe_div_up(h, 32)), 256, 0, ctx->stream>>>(
            (const uint32_t *)src, (uint32_t *)dst, (int)w, (int)h, op));
    else
        PFE_KERNEL(ctx, "mirror", mirror_kernel<<<grid2d(w, h), 256, 0, ctx->stream>>>((const uint32_t *)src, (uint32_t *)dst, (int)w, (int)h, op));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_resize_canvas(pfe_ctx *ctx, const uint8_t *src, uint32_t old_w, uint32_t old_h, uint32_t new_w,
                                     uint32_t new_h, uint32_t anchor_x, uint32_t anchor_y, const uint8_t fill[4], uint8_t *dst) {
    PFE_TRY(check2(ctx, src, dst, old_w, old_h, new_w, new_h, "resize_canvas: bad args"));
    if (!fill) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "resize_canvas: null fill");
    const int32_t dw = (int32_t)new_w - (int32_t)old_w, dh = (int32_t)new_h - (int32_t)old_h;
    const int32_t offx = anchor_x == 0 ? 0 : (anchor_x == 1 ? dw / 2 : dw), offy = anchor_y == 0 ? 0 : (anchor_y == 1 ? dh / 2 : dh);
    const uint32_t f = (uint32_t)fill[0] | ((uint32_t)fill[1] << 8) | ((uint32_t)fill[2] << 16) | ((uint32_t)fill[3] << 24);
    PFE_KERNEL(ctx, "resize_canvas", resize_canvas_kernel<<<grid2d(new_w, new_h), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, (uint32_t *)dst, (int)old_w, (int)old_h, (int)new_w, (int)new_h, offx, offy, f));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_affine(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h, uint32_t canvas_w,
                              uint32_t canvas_h, float rotation_z, float rotation_x, float rotation_y, float scale,
                              float offset_x, float offset_y, int nearest, uint8_t *dst) {
    PFE_TRY(check2(ctx, src, dst, src_w, src_h, canvas_w, canvas_h, "affine: bad args"));
    // transform.rs:838-876: R = Rz * Ry * Rx, perspective with focal = 1.5 * max(w, h); host libm like the reference
    const float focal = (float)std::max(canvas_w, canvas_h) * 1.5f;
    const float rz = to_radians(rotation_z), rx = to_radians(rotation_x), ry = to_radians(rotation_y);
    const float sz = sinf(rz), cz = cosf(rz), sxr = sinf(rx), cxr = cosf(rx), syr = sinf(ry), cyr = cosf(ry);
    const float r00 = cz * cyr, r01 = cz * syr * sxr - sz * cxr, r10 = sz * cyr, r11 = sz * syr * sxr + cz * cxr;
    const float r20 = -syr, r21 = cyr * sxr;
    const float hm[3][3] = {{focal * r00, focal * r01, 0.0f}, {focal * r10, focal * r11, 0.0f}, {r20, r21, focal}};
    float hi[3][3];
    invert_3x3(hm, hi);
    AffineParams P;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) P.m[r * 3 + c] = hi[r][c];
    P.inv_scale = fabsf(scale) > 1e-6f ? 1.0f / scale : 1.0f;
    P.cx = (float)canvas_w * 0.5f; P.cy = (float)canvas_h * 0.5f;
    P.off_x = offset_x; P.off_y = offset_y;
    P.src_w = (int)src_w; P.src_h = (int)src_h; P.nearest = nearest ? 1 : 0;
    PFE_KERNEL(ctx, "affine", affine_kernel<<<grid2d(canvas_w, canvas_h), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, (uint32_t *)dst, (int)canvas_w, (int)canvas_h, P));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_resize(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t new_w, uint32_t new_h,
                              int filter, uint8_t *dst) {
    PFE_TRY(check2(ctx, src, dst, w, h, new_w, new_h, "resize: bad args"));
    if (filter < PFE_RESIZE_NEAREST || filter > PFE_RESIZE_LANCZOS3) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "resize: unknown filter");
    if (new_w == w && new_h == h) {  // sample.rs: same dimensions -> plain copy
        PFE_CUDA(ctx, cudaMemcpyAsync(dst, src, (size_t)w * h * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return PFE_OK;
    }
    std::vector<uint32_t> vt, ht;
    std::vector<float> vw, hw;
    axis_weights(h, new_h, filter, vt, vw);
    axis_weights(w, new_w, filter, ht, hw);
    // device layout in scratch C: [v table | h table | v weights | h weights]; the f32 image in the F32 slot
    const size_t bytes = (vt.size() + ht.size()) * 4 + (vw.size() + hw.size()) * 4;
    void *tab, *tmp;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, bytes, &tab));
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_F32, (size_t)w * new_h * 16, &tmp));
    uint32_t *d_vt = (uint32_t *)tab, *d_ht = d_vt + vt.size();
    float *d_vw = (float *)(d_ht + ht.size()), *d_hw = d_vw + vw.size();
    // pageable sources: these copies complete with respect to the host buffers before returning
    PFE_CUDA(ctx, cudaMemcpyAsync(d_vt, vt.data(), vt.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaMemcpyAsync(d_ht, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaMemcpyAsync(d_vw, vw.data(), vw.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaMemcpyAsync(d_hw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    const AxisTable TV{d_vt, d_vt + new_h, d_vt + 2 * (size_t)new_h, d_vw}, TH{d_ht, d_ht + new_w, d_ht + 2 * (size_t)new_w, d_hw};
    PFE_KERNEL(ctx, "resize_v", resize_v_kernel<<<grid2d(w, new_h), 256, 0, ctx->stream>>>((const uint32_t *)src, (float4 *)tmp, (int)w, (int)new_h, TV));
    PFE_LAUNCHED(ctx);
    PFE_KERNEL(ctx, "resize_h", resize_h_kernel<<<grid2d(new_w, new_h), 256, 0, ctx->stream>>>((const float4 *)tmp, (uint32_t *)dst, (int)w, (int)new_w, (int)new_h, TH));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

// ---- host tier: shapes differ between source and result, so staging is explicit --------------------
namespace {
template <class F>
int host_reshape(pfe_ctx *ctx, const uint8_t *src, size_t src_px, uint8_t *dst, size_t dst_px, F call) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !dst || !src_px || !dst_px) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "null image or zero size");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sa = (src_px * 4 + 255) & ~size_t(255);
    void *a;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_A, sa + dst_px * 4, &a));
    uint8_t *ds = (uint8_t *)a, *dd = ds + sa;
    PFE_CUDA(ctx, cudaMemcpyAsync(ds, src, src_px * 4, cudaMemcpyHostToDevice, ctx->stream));
    PFE_TRY(call(ds, dd));
    PFE_CUDA(ctx, cudaMemcpyAsync(dst, dd, dst_px * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}
}  // namespace

extern "C" int pfe_orient(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int op, uint8_t *dst) {
    return host_reshape(ctx, src, (size_t)w * h, dst, (size_t)w * h, [&](uint8_t *s, uint8_t *d) { return pfe_dev_orient(ctx, s, w, h, op, d); });
}
extern "C" int pfe_resize_canvas(pfe_ctx *ctx, const uint8_t *src, uint32_t old_w, uint32_t old_h, uint32_t new_w, uint32_t new_h,
                                 uint32_t anchor_x, uint32_t anchor_y, const uint8_t fill[4], uint8_t *dst) {
    return host_reshape(ctx, src, (size_t)old_w * old_h, dst, (size_t)new_w * new_h, [&](uint8_t *s, uint8_t *d) {
        return pfe_dev_resize_canvas(ctx, s, old_w, old_h, new_w, new_h, anchor_x, anchor_y, fill, d);
    });
}
extern "C" int pfe_affine(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h, uint32_t canvas_w, uint32_t canvas_h,
                          float rotation_z, float rotation_x, float rotation_y, float scale, float offset_x, float offset_y,
                          int nearest, uint8_t *dst) {
    return host_reshape(ctx, src, (size_t)src_w * src_h, dst, (size_t)canvas_w * canvas_h, [&](uint8_t *s, uint8_t *d) {
        return pfe_dev_affine(ctx, s, src_w, src_h, canvas_w, canvas_h, rotation_z, rotation_x, rotation_y, scale, offset_x, offset_y, nearest, d);
    });
}
extern "C" int pfe_resize(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t new_w, uint32_t new_h, int filter,
                          uint8_t *dst) {
    return host_reshape(ctx, src, (size_t)w * h, dst, (size_t)new_w * new_h, [&](uint8_t *s, uint8_t *d) {
        return pfe_dev_resize(ctx, s, w, h, new_w, new_h, filter, d);
    });
}
