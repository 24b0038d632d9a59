// Displacement warp, Catmull-Rom mesh displacement, the fused mesh warp, and the Liquify brushes.
// Reference: src/ops/transform.rs:1015-1345 and :1558-1761; wgpu twins
// src/gpu/compute/liquify.rs:176 and src/gpu/compute/mesh_warp.rs:131.
// The "Catmull-Rom" part is the control-point surface; the pixel resample is bilinear with a
// transparent border (transform.rs:1317-1341) and that is what is built here.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

struct MeshParams {
    float orig[PFE_MESH_MAX_POINTS * 2];
    float def[PFE_MESH_MAX_POINTS * 2];
    int cols, rows;
    int has_orig;
};

__device__ __forceinline__ void cr_weights(float t, float w[4]) {  // transform.rs:1558
    float t2 = t * t, t3 = t2 * t;
    w[0] = -0.5f * t3 + t2 - 0.5f * t;
    w[1] = 1.5f * t3 - 2.5f * t2 + 1.0f;
    w[2] = -1.5f * t3 + 2.0f * t2 + 0.5f * t;
    w[3] = 0.5f * t3 - 0.5f * t2;
}

// catmull_rom_surface (transform.rs:1589-1646) and generate_displacement_from_mesh[_fast] (:1670-1739),
// factored for a thread that walks DOWN a column: everything that depends on x alone
// (column index, the four column weights and clamped column indices) is computed once, and the four
// row-wise interpolants rx[j], ry[j] - 8 of the 10 dot products of catmull_rom_surface - depend only on x
// and on the mesh row `ri`, so they are recomputed only when the walk crosses into the next mesh row.
// Every product and sum is the one catmull_rom_surface evaluates, in the same order: results are identical.
constexpr int kMeshRows = 8;  // consecutive output rows per thread
struct MeshColumn {
    float ug, px;  // px = x + 0.5
    float wu[4];
    int cu[4];
};
__device__ __forceinline__ MeshColumn mesh_column(const MeshParams &M, int x, uint32_t w) {
    MeshColumn C;
    C.px = (float)x + 0.5f;
    C.ug = C.px / (float)w * (float)M.cols;
    const int ppr = M.cols + 1;
    const float col_f = pfe_clampf(C.ug, 0.0f, (float)M.cols - 0.0001f);
    int ci = max(min(__float2int_rz(col_f), M.cols - 1), 0);
    cr_weights(col_f - (float)ci, C.wu);
    C.cu[0] = ci == 0 ? 0 : ci - 1; C.cu[1] = ci; C.cu[2] = min(ci + 1, ppr - 1); C.cu[3] = min(ci + 2, ppr - 1);
    return C;
}
struct MeshRowCache {
    int ri;
    float ax[4], ay[4], bx[4], by[4];  // row interpolants of the deformed (a) and original (b) grids
};
__device__ __forceinline__ void mesh_rows(const float *pts, int ppr, int nrows, int ri, const MeshColumn &C, float rx[4], float ry[4]) {
    const int rv[4] = {ri == 0 ? 0 : ri - 1, ri, min(ri + 1, nrows - 1), min(ri + 2, nrows - 1)};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float *b = pts + (size_t)rv[j] * ppr * 2;
        const float *p0 = b + C.cu[0] * 2, *p1 = b + C.cu[1] * 2, *p2 = b + C.cu[2] * 2, *p3 = b + C.cu[3] * 2;
        rx[j] = C.wu[0] * p0[0] + C.wu[1] * p1[0] + C.wu[2] * p2[0] + C.wu[3] * p3[0];
        ry[j] = C.wu[0] * p0[1] + C.wu[1] * p1[1] + C.wu[2] * p2[1] + C.wu[3] * p3[1];
    }
}
// generate_displacement_from_mesh[_fast] for pixel (C's x, y), reusing `rc` while the mesh row is unchanged
__device__ __forceinline__ void mesh_disp_col(const MeshParams &M, const float *sdef, const float *sorig, const MeshColumn &C,
                                              MeshRowCache &rc, int y, uint32_t h, float &dx, float &dy) {
    const float vg = ((float)y + 0.5f) / (float)h * (float)M.rows;
    const float row_f = pfe_clampf(vg, 0.0f, (float)M.rows - 0.0001f);
    const int ri = max(min(__float2int_rz(row_f), M.rows - 1), 0);
    float wv[4];
    cr_weights(row_f - (float)ri, wv);
    if (ri != rc.ri) {
        rc.ri = ri;
        mesh_rows(sdef, M.cols + 1, M.rows + 1, ri, C, rc.ax, rc.ay);
        if (M.has_orig) mesh_rows(sorig, M.cols + 1, M.rows + 1, ri, C, rc.bx, rc.by);
    }
    const float ax = wv[0] * rc.ax[0] + wv[1] * rc.ax[1] + wv[2] * rc.ax[2] + wv[3] * rc.ax[3];
    const float ay = wv[0] * rc.ay[0] + wv[1] * rc.ay[1] + wv[2] * rc.ay[2] + wv[3] * rc.ay[3];
    if (M.has_orig) {
        const float bx = wv[0] * rc.bx[0] + wv[1] * rc.bx[1] + wv[2] * rc.bx[2] + wv[3] * rc.bx[3];
        const float by = wv[0] * rc.by[0] + wv[1] * rc.by[1] + wv[2] * rc.by[2] + wv[3] * rc.by[3];
        dx = ax - bx; dy = ay - by;
    } else {
        dx = ax - C.px; dy = ay - ((float)y + 0.5f);
    }
}

// warp_displacement_full inner body, transform.rs:1303-1342
// `src` holds source rows [sy0, sy0 + snr) of an sw x sh image (sy0 = 0, snr = sh for a whole image).
// A tap inside the image but outside the provided rows raises *missing (band callers size their halo
// so this never happens; the flag turns a wrong halo into an error instead of a wrong pixel).
__device__ __forceinline__ uint32_t warp_sample(const uint32_t *src, int sw, int sh, int x, int y, float ddx, float ddy,
                                                int sy0 = 0, int snr = 0x7fffffff, int *missing = nullptr) {
    float sx = (float)x - ddx, sy = (float)y - ddy;
    float flx = floorf(sx), fly = floorf(sy);
    // `as i32` saturates; anything outside [-1, size) leaves the pixel transparent
    if (!(flx >= -1.0f) || !(fly >= -1.0f) || !(flx < (float)sw) || !(fly < (float)sh)) return 0u;
    int x0 = (int)flx, y0 = (int)fly;
    float fx = sx - flx, fy = sy - fly;
    uint32_t q[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int px = x0 + (k & 1), py = y0 + (k >> 1);
        if (px < 0 || py < 0 || px >= sw || py >= sh) q[k] = 0u;
        else if (py < sy0 || py - sy0 >= snr) { q[k] = 0u; if (missing) atomicOr(missing, PFE_ASYNC_WARP_WINDOW); }
        else q[k] = __ldg(src + (size_t)(py - sy0) * sw + px);
    }
    uint32_t o[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        float tl = (float)((q[0] >> (8 * c)) & 255u), tr = (float)((q[1] >> (8 * c)) & 255u);
        float bl = (float)((q[2] >> (8 * c)) & 255u), br = (float)((q[3] >> (8 * c)) & 255u);
        float top = tl + (tr - tl) * fx;
        float bot = bl + (br - bl) * fx;
        o[c] = pfe_round_u8(top + (bot - top) * fy);
    }
    return pfe_pack(o[0], o[1], o[2], o[3]);
}

// Four rows per thread, all loads of the four pixels (4 field entries, 16 taps) issued before any store: the
// kernel is a gather, bound by how many loads an SM keeps in flight, not by arithmetic or DRAM.
constexpr int kWarpRows = 4;
__global__ void __launch_bounds__(256) warp_kernel(const uint32_t *__restrict__ src, int sw, int sh, const float2 *__restrict__ disp,
                                                   uint32_t *__restrict__ dst, int w, int h) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), ya = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kWarpRows;
    if (x >= w || ya >= h) return;
    float2 d[kWarpRows];
    uint32_t o[kWarpRows];
#pragma unroll
    for (int r = 0; r < kWarpRows; r++) d[r] = (ya + r < h) ? __ldg(disp + (size_t)(ya + r) * w + x) : make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kWarpRows; r++) o[r] = (ya + r < h) ? warp_sample(src, sw, sh, x, ya + r, d[r].x, d[r].y) : 0u;
#pragma unroll
    for (int r = 0; r < kWarpRows; r++)
        if (ya + r < h) dst[(size_t)(ya + r) * w + x] = o[r];
}

// warp_displacement_region: the same sample, only for the pixels of the dirty rectangle
__global__ void __launch_bounds__(256) warp_region_kernel(const uint32_t *__restrict__ src, int sw, int sh, const float2 *__restrict__ disp,
                                                          uint32_t *__restrict__ dst, int w, int rx0, int ry0, int rx1, int ry1) {
    const int x = rx0 + blockIdx.x * 32 + (threadIdx.x & 31), y = ry0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= rx1 || y >= ry1) return;
    const float2 d = __ldg(disp + (size_t)y * w + x);
    dst[(size_t)y * w + x] = warp_sample(src, sw, sh, x, y, d.x, d.y);
}

__global__ void __launch_bounds__(256) mesh_disp_kernel(const __grid_constant__ MeshParams M, float2 *out, uint32_t w,
                                                        uint32_t h) {
    __shared__ float sdef[PFE_MESH_MAX_POINTS * 2], sorig[PFE_MESH_MAX_POINTS * 2];
    const int np = (M.cols + 1) * (M.rows + 1) * 2;
    for (int i = threadIdx.x; i < np; i += blockDim.x) { sdef[i] = M.def[i]; sorig[i] = M.orig[i]; }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x >= (int)w) return;
    const MeshColumn C = mesh_column(M, x, w);
    MeshRowCache rc;
    rc.ri = -1;
    const int ya = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kMeshRows;
    for (int y = ya; y < ya + kMeshRows && y < (int)h; y++) {
        float dx, dy;
        mesh_disp_col(M, sdef, sorig, C, rc, y, h, dx, dy);
        out[(size_t)y * w + x] = make_float2(dx, dy);
    }
}

// warp_mesh_catmull_rom fused: displacement stays in registers. Rows [y0, y0+rows_out).
__global__ void __launch_bounds__(256) mesh_warp_kernel(const __grid_constant__ MeshParams M, const uint32_t *src,
                                                        int sw, int sh, uint32_t *dst, uint32_t w, uint32_t h,
                                                        uint32_t y0, uint32_t rows_out) {
    __shared__ float sdef[PFE_MESH_MAX_POINTS * 2], sorig[PFE_MESH_MAX_POINTS * 2];
    const int np = (M.cols + 1) * (M.rows + 1) * 2;
    for (int i = threadIdx.x; i < np; i += blockDim.x) { sdef[i] = M.def[i]; sorig[i] = M.orig[i]; }
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x >= (int)w) return;
    const MeshColumn C = mesh_column(M, x, w);
    MeshRowCache rc;
    rc.ri = -1;
    const uint32_t ra = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kMeshRows;
    for (uint32_t ry = ra; ry < ra + kMeshRows && ry < rows_out; ry++) {
        const int y = (int)(y0 + ry);
        float dx, dy;
        mesh_disp_col(M, sdef, sorig, C, rc, y, h, dx, dy);
        dst[(size_t)ry * w + x] = warp_sample(src, sw, sh, x, y, dx, dy);
    }
}

// Band form shared by the displacement warp and the fused mesh warp: output rows [y0, y0+rows_out)
// of a w x h result, source given as a row window. disp == nullptr selects the mesh path.
template <bool MESH>
__global__ void __launch_bounds__(256) warp_band_kernel(const __grid_constant__ MeshParams M, const uint32_t *__restrict__ src, int sw,
                                                        int sh, int sy0, int snr, const float2 *__restrict__ disp,
                                                        uint32_t *__restrict__ dst, uint32_t w, uint32_t h, uint32_t y0,
                                                        uint32_t rows_out, int *missing) {
    __shared__ float sdef[MESH ? PFE_MESH_MAX_POINTS * 2 : 1], sorig[MESH ? PFE_MESH_MAX_POINTS * 2 : 1];
    if (MESH) {
        const int np = (M.cols + 1) * (M.rows + 1) * 2;
        for (int i = threadIdx.x; i < np; i += blockDim.x) { sdef[i] = M.def[i]; sorig[i] = M.orig[i]; }
        __syncthreads();
    }
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    if (x >= (int)w) return;
    const uint32_t ra = (blockIdx.y * 8 + (threadIdx.x >> 5)) * kMeshRows;
    if (MESH) {
        const MeshColumn C = mesh_column(M, x, w);
        MeshRowCache rc;
        rc.ri = -1;
        for (uint32_t ry = ra; ry < ra + kMeshRows && ry < rows_out; ry++) {
            const int y = (int)(y0 + ry);
            float dx, dy;
            mesh_disp_col(M, sdef, sorig, C, rc, y, h, dx, dy);
            dst[(size_t)ry * w + x] = warp_sample(src, sw, sh, x, y, dx, dy, sy0, snr, missing);
        }
    } else {  // field path: loads of kWarpRows pixels batched like warp_kernel
#pragma unroll
        for (int half = 0; half < kMeshRows / kWarpRows; half++) {
            const uint32_t rb = ra + half * kWarpRows;
            float2 d[kWarpRows];
            uint32_t o[kWarpRows];
#pragma unroll
            for (int r = 0; r < kWarpRows; r++) d[r] = (rb + r < rows_out) ? __ldg(disp + (size_t)(rb + r) * w + x) : make_float2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < kWarpRows; r++)
                o[r] = (rb + r < rows_out) ? warp_sample(src, sw, sh, x, (int)(y0 + rb + r), d[r].x, d[r].y, sy0, snr, missing) : 0u;
#pragma unroll
            for (int r = 0; r < kWarpRows; r++)
                if (rb + r < rows_out) dst[(size_t)(rb + r) * w + x] = o[r];
        }
    }
}

// exp() as the reference's libm expf sees it: correctly rounded f32. Evaluated in f64 and rounded
// once, which agrees with a <1-ulp libm except when the f64 result sits within ~1e-16 of a
// rounding boundary.
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }

// DisplacementField::apply_*, transform.rs:1051-1200, one thread per bbox pixel
__global__ void __launch_bounds__(256) liquify_kernel(float2 *field, uint32_t w, int kind, float cx, float cy, float r,
                                                      float s2, float strength, float a0, float a1, int x0, int y0,
                                                      int bw, int bh) {
    const int lx = blockIdx.x * 32 + (threadIdx.x & 31), ly = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (lx >= bw || ly >= bh) return;
    const int px = x0 + lx, py = y0 + ly;
    float dx = (float)px - cx, dy = (float)py - cy;
    float dsq = dx * dx + dy * dy;
    if (dsq > r * r) return;
    float2 *f = field + (size_t)py * w + px;
    float2 v = *f;
    if (kind == PFE_LIQ_PUSH) {
        float wgt = exp_cr(-dsq / s2) * strength;
        v.x += a0 * wgt; v.y += a1 * wgt;
    } else if (kind == PFE_LIQ_EXPAND) {
        float dist = fmaxf(sqrtf(dsq), 0.001f);
        float t = dist / r;
        float wgt = (1.0f - t) * (1.0f - t) * strength * 3.0f;
        v.x += dx / dist * wgt; v.y += dy / dist * wgt;
    } else if (kind == PFE_LIQ_CONTRACT) {
        float dist = fmaxf(sqrtf(dsq), 0.001f);
        float wgt = exp_cr(-dsq / s2) * strength;
        v.x += -dx / dist * wgt * 2.0f; v.y += -dy / dist * wgt * 2.0f;
    } else {
        float dir = a0 != 0.0f ? 1.0f : -1.0f;
        float wgt = exp_cr(-dsq / s2) * strength * dir;
        v.x += -dy * wgt * 0.1f; v.y += dx * wgt * 0.1f;
    }
    *f = v;
}

int fill_mesh(pfe_ctx *ctx, MeshParams *M, const float *orig, const float *def, uint32_t cols, uint32_t rows) {
    if (!def || cols == 0 || rows == 0) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "mesh: bad grid");
    const size_t np = (size_t)(cols + 1) * (rows + 1);
    if (np > PFE_MESH_MAX_POINTS) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "mesh: more than PFE_MESH_MAX_POINTS control points");
    memset(M, 0, sizeof(*M));
    memcpy(M->def, def, np * 2 * sizeof(float));
    if (orig) memcpy(M->orig, orig, np * 2 * sizeof(float));
    M->cols = (int)cols; M->rows = (int)rows; M->has_orig = orig ? 1 : 0;
    return PFE_OK;
}

int f2i_sat(float v) {  // Rust `as i32`
    if (v != v) return 0;
    if (v <= -2147483648.0f) return INT32_MIN;
    if (v >= 2147483648.0f) return INT32_MAX;
    return (int)v;
}

}  // namespace

extern "C" int pfe_dev_warp_displacement(pfe_ctx *ctx, const uint8_t *src, uint32_t sw, uint32_t sh, const float *disp,
                                         uint32_t w, uint32_t h, uint8_t *dst) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !disp || !dst || !sw || !sh || !w || !h || src == dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "warp: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    PFE_KERNEL(ctx, "warp", warp_kernel<<<dim3(pfe_div_up(w, 32), pfe_div_up(h, 8 * kWarpRows)), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, (int)sw, (int)sh, (const float2 *)disp, (uint32_t *)dst, (int)w, (int)h));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_warp_displacement_region(pfe_ctx *ctx, const uint8_t *src, uint32_t sw, uint32_t sh, const float *disp,
                                                const uint8_t *prev, const int32_t rect[4], uint32_t w, uint32_t h, uint8_t *dst) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !disp || !prev || !rect || !dst || !sw || !sh || !w || !h || src == dst)
        return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "warp_region: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    if (dst != prev) PFE_CUDA(ctx, cudaMemcpyAsync(dst, prev, (size_t)w * h * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    // transform.rs:1214-1218
    const uint32_t x0 = rect[0] > 0 ? (uint32_t)rect[0] : 0u, y0 = rect[1] > 0 ? (uint32_t)rect[1] : 0u;
    const uint32_t x1 = std::min((uint32_t)rect[2], w), y1 = std::min((uint32_t)rect[3], h);
    if (x1 <= x0 || y1 <= y0) return PFE_OK;
    PFE_KERNEL(ctx, "warp_region", warp_region_kernel<<<dim3(pfe_div_up(x1 - x0, 32), pfe_div_up(y1 - y0, 8)), 256, 0, ctx->stream>>>(
        (const uint32_t *)src, (int)sw, (int)sh, (const float2 *)disp, (uint32_t *)dst, (int)w, (int)x0, (int)y0, (int)x1, (int)y1));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_mesh_displacement(pfe_ctx *ctx, const float *orig, const float *def, uint32_t cols, uint32_t rows,
                                         uint32_t w, uint32_t h, float *out) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!out || !w || !h) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "mesh_displacement: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    MeshParams M;
    PFE_TRY(fill_mesh(ctx, &M, orig, def, cols, rows));
    PFE_KERNEL(ctx, "mesh_disp", mesh_disp_kernel<<<dim3(pfe_div_up(w, 32), pfe_div_up(h, 8 * kMeshRows)), 256, 0, ctx->stream>>>(M, (float2 *)out, w, h));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_mesh_warp(pfe_ctx *ctx, const uint8_t *src, uint32_t sw, uint32_t sh, const float *orig,
                                 const float *def, uint32_t cols, uint32_t rows, uint32_t w, uint32_t h, uint32_t y0,
                                 uint32_t rows_out, uint8_t *dst) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src || !dst || !sw || !sh || !w || !h || !rows_out || (uint64_t)y0 + rows_out > h)
        return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "mesh_warp: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    MeshParams M;
    PFE_TRY(fill_mesh(ctx, &M, orig, def, cols, rows));
    PFE_KERNEL(ctx, "mesh_warp", mesh_warp_kernel<<<dim3(pfe_div_up(w, 32), pfe_div_up(rows_out, 8 * kMeshRows)), 256, 0, ctx->stream>>>(
        M, (const uint32_t *)src, (int)sw, (int)sh, (uint32_t *)dst, w, h, y0, rows_out));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_warp_band(pfe_ctx *ctx, const uint8_t *src_rows, uint32_t sw, uint32_t sh, uint32_t src_y0,
                                 uint32_t src_nrows, const float *disp_band, const float *orig, const float *def,
                                 uint32_t cols, uint32_t rows, uint32_t w, uint32_t h, uint32_t y0, uint32_t rows_out,
                                 uint8_t *dst_band) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!src_rows || !dst_band || !sw || !sh || !w || !h || !rows_out || (uint64_t)y0 + rows_out > h ||
        (uint64_t)src_y0 + src_nrows > sh || !src_nrows || (!disp_band && !def))
        return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "warp_band: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    MeshParams M;
    memset(&M, 0, sizeof(M));
    if (!disp_band) PFE_TRY(fill_mesh(ctx, &M, orig, def, cols, rows));
    int *flag = ctx->async_err;  // sticky: read (and cleared) by pfe_ctx_check_async, so this call stays asynchronous
    const dim3 grid(pfe_div_up(w, 32), pfe_div_up(rows_out, 8 * kMeshRows));
    if (disp_band)
        PFE_KERNEL(ctx, "warp_band", warp_band_kernel<false><<<grid, 256, 0, ctx->stream>>>(
            M, (const uint32_t *)src_rows, (int)sw, (int)sh, (int)src_y0, (int)src_nrows, (const float2 *)disp_band,
            (uint32_t *)dst_band, w, h, y0, rows_out, flag));
    else
        PFE_KERNEL(ctx, "warp_band", warp_band_kernel<true><<<grid, 256, 0, ctx->stream>>>(
            M, (const uint32_t *)src_rows, (int)sw, (int)sh, (int)src_y0, (int)src_nrows, nullptr,
            (uint32_t *)dst_band, w, h, y0, rows_out, flag));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

// How far a band's displacement field reaches outside the band: min / max over the band of
// floor(clamp(y - dy, -1, h)), non-finite dy counted as 0 (such taps read nothing, transform.rs:1303-1312).
// out = {min, max} stays on the device so the caller can fold it into its collective before reading it.
namespace {
__global__ void __launch_bounds__(256) disp_reach_kernel(const float2 *disp, uint32_t w, uint32_t rows, uint32_t y0, float h_total,
                                                         int *out) {
    int mn = 0x7FFFFFFF, mx = (int)0x80000000;
    const size_t n = (size_t)w * rows;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float dy = __ldg(&disp[i].y);
        if (!(fabsf(dy) <= 3.402823466e+38f)) dy = 0.0f;  // NaN / +-inf
        const float sy = fminf(fmaxf((float)(y0 + (uint32_t)(i / w)) - dy, -1.0f), h_total);
        const int f = __float2int_rd(sy);
        mn = min(mn, f);
        mx = max(mx, f);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(out, mn); atomicMax(out + 1, mx); }
}
}  // namespace
extern "C" int pfe_dev_disp_reach(pfe_ctx *ctx, const float *disp_band, uint32_t w, uint32_t rows, uint32_t y0, uint32_t h_total,
                                  int32_t *minmax_dev) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!disp_band || !minmax_dev || !w || !rows) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "disp_reach: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    const int init[2] = {0x7FFFFFFF, (int)0x80000000};
    void *stage;
    PFE_TRY(pfe_small_upload(ctx, init, sizeof(init), &stage));
    PFE_CUDA(ctx, cudaMemcpyAsync(minmax_dev, stage, sizeof(init), cudaMemcpyDeviceToDevice, ctx->stream));
    PFE_KERNEL(ctx, "disp_reach", disp_reach_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(
        (const float2 *)disp_band, w, rows, y0, (float)h_total, (int *)minmax_dev));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}

extern "C" int pfe_dev_liquify(pfe_ctx *ctx, float *field, uint32_t w, uint32_t h, int kind, float cx, float cy,
                               float radius, float strength, float a0, float a1, int32_t bbox[4]) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if (!field || !w || !h || kind < 0 || kind > 3) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "liquify: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    const float r = fmaxf(radius, 1.0f);
    const float sigma = r / 3.0f;
    const float s2 = 2.0f * sigma * sigma;
    int x0 = std::max(f2i_sat(floorf(cx - r)), 0), y0 = std::max(f2i_sat(floorf(cy - r)), 0);
    int x1 = std::min(f2i_sat(ceilf(cx + r)), (int)w), y1 = std::min(f2i_sat(ceilf(cy + r)), (int)h);
    if (bbox) { bbox[0] = x0; bbox[1] = y0; bbox[2] = x1; bbox[3] = y1; }
    if (x1 <= x0 || y1 <= y0) return PFE_OK;
    PFE_KERNEL(ctx, "liquify", liquify_kernel<<<dim3(pfe_div_up(x1 - x0, 32), pfe_div_up(y1 - y0, 8)), 256, 0, ctx->stream>>>(
        (float2 *)field, w, kind, cx, cy, r, s2, strength, a0, a1, x0, y0, x1 - x0, y1 - y0));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}
