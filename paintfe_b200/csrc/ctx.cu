// Context, memory helpers and the host-pointer tier of the C ABI.
// The host tier is plumbing only: copy in, call the pfe_dev_* entry point, copy out.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <vector>

#include "common.cuh"

int pfe_fail(pfe_ctx *ctx, int code, const char *what, cudaError_t e) {
    if (ctx) {
        ctx->err = what ? what : "";
        if (e != cudaSuccess) {
            ctx->err += ": ";
            ctx->err += cudaGetErrorString(e);
        }
    }
    return code;
}

int pfe_scratch(pfe_ctx *ctx, int slot, size_t bytes, void **out) {
    if (bytes > ctx->scratch_bytes[slot]) {
        if (ctx->scratch[slot]) {
            PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            PFE_CUDA(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return pfe_fail(ctx, PFE_ERR_OOM, "scratch cudaMalloc", e);
        }
        ctx->scratch_bytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return PFE_OK;
}

int pfe_small_upload(pfe_ctx *ctx, const void *host, size_t bytes, void **dev_out) {
    size_t need = (bytes + 255) & ~size_t(255);
    if (need > PFE_SMALL_BYTES) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "parameter table too large");
    if (ctx->small_cursor + need > PFE_SMALL_BYTES) {
        // the pinned staging ring is about to be reused: make sure earlier copies have drained
        PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->small_cursor = 0;
    }
    char *stage = (char *)ctx->pinned + ctx->small_cursor;
    char *dev = (char *)ctx->dev_small + ctx->small_cursor;
    memcpy(stage, host, bytes);
    PFE_CUDA(ctx, cudaMemcpyAsync(dev, stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->small_cursor += need;
    *dev_out = dev;
    return PFE_OK;
}

static cudaEvent_t take_event(pfe_ctx *c) {
    cudaEvent_t e = nullptr;
    if (!c->event_pool.empty()) { e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return e;
}
pfe_span::pfe_span(pfe_ctx *ctx, const char *n) : c(ctx), name(n) {
    if (!c->profiling || c->spans.size() >= (1u << 16)) { c = nullptr; return; }
    a = take_event(c);
    b = take_event(c);
    if (!a || !b) { c = nullptr; return; }
    cudaEventRecord(a, c->stream);
}
pfe_span::~pfe_span() {
    if (!c) return;
    cudaEventRecord(b, c->stream);
    c->spans.push_back({name, a, b});
}

extern "C" {

int pfe_abi_version(void) { return PFE_ABI_VERSION; }

int pfe_ctx_profile(pfe_ctx *c, int enable) {
    if (!c) return PFE_ERR_INVALID_ARG;
    c->profiling = enable != 0;
    return PFE_OK;
}

int pfe_ctx_profile_read(pfe_ctx *c, char *buf, size_t cap) {
    if (!c || !buf || cap < 4) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(c, cudaSetDevice(c->device));
    PFE_CUDA(c, cudaStreamSynchronize(c->stream));
    struct Agg { const char *name; uint64_t n; double ms; };
    std::vector<Agg> agg;
    for (auto &s : c->spans) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.a, s.b);
        c->event_pool.push_back(s.a);
        c->event_pool.push_back(s.b);
        size_t k = 0;
        for (; k < agg.size(); k++) if (strcmp(agg[k].name, s.name) == 0) break;
        if (k == agg.size()) agg.push_back({s.name, 0, 0.0});
        agg[k].n++;
        agg[k].ms += ms;
    }
    c->spans.clear();
    std::string out = "{";
    for (size_t k = 0; k < agg.size(); k++) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f}", k ? ", " : "", agg[k].name,
                 (unsigned long long)agg[k].n, agg[k].ms);
        out += tmp;
    }
    out += "}";
    if (out.size() + 1 > cap) return pfe_fail(c, PFE_ERR_INVALID_ARG, "profile buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return PFE_OK;
}

int pfe_ctx_create(int device, pfe_ctx **out) {
    if (!out) return PFE_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return PFE_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) return PFE_ERR_INVALID_ARG;
    pfe_ctx *c = new (std::nothrow) pfe_ctx();
    if (!c) return PFE_ERR_OOM;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return PFE_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    bool ok = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming) == cudaSuccess &&
              cudaMallocHost(&c->pinned, PFE_SMALL_BYTES) == cudaSuccess &&
              cudaMalloc(&c->dev_small, PFE_SMALL_BYTES) == cudaSuccess &&
              cudaMalloc((void **)&c->async_err, PFE_ASYNC_BLOCK_BYTES) == cudaSuccess && cudaMemset(c->async_err, 0, PFE_ASYNC_BLOCK_BYTES) == cudaSuccess;
    if (!ok) { pfe_ctx_destroy(c); return PFE_ERR_CUDA; }
    c->pinned_bytes = PFE_SMALL_BYTES;
    c->stream = c->own_stream;
    *out = c;
    return PFE_OK;
}

int pfe_ctx_destroy(pfe_ctx *c) {
    if (!c) return PFE_ERR_INVALID_ARG;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto &s : c->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < 4; i++) if (c->scratch[i]) cudaFree(c->scratch[i]);
    if (c->dev_small) cudaFree(c->dev_small);
    if (c->async_err) cudaFree(c->async_err);
    if (c->gauss_mem) cudaFree(c->gauss_mem);
    for (uint8_t *slab : c->chunks.slabs) cudaFree(slab);
    if (c->pinned) cudaFreeHost(c->pinned);
    for (int i = 0; i < 2; i++) {
        if (c->stage[i]) cudaFreeHost(c->stage[i]);
        if (c->stage_ev[i]) cudaEventDestroy(c->stage_ev[i]);
    }
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return PFE_OK;
}

// the device-resident weight tables are ordered on the stream that filled them: forget them on a switch
static void drop_stream_ordered_caches(pfe_ctx *c) {
    for (auto &g : c->gauss_slots) g.valid = false;
}

int pfe_ctx_set_stream(pfe_ctx *c, void *s) {
    if (!c) return PFE_ERR_INVALID_ARG;
    if (c->stream != (cudaStream_t)s) drop_stream_ordered_caches(c);
    c->stream = (cudaStream_t)s;
    return PFE_OK;
}

int pfe_ctx_use_own_stream(pfe_ctx *c) {
    if (!c) return PFE_ERR_INVALID_ARG;
    if (c->stream != c->own_stream) drop_stream_ordered_caches(c);
    c->stream = c->own_stream;
    return PFE_OK;
}

int pfe_ctx_sync(pfe_ctx *c) {
    if (!c) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(c, cudaSetDevice(c->device));
    PFE_CUDA(c, cudaStreamSynchronize(c->stream));
    return PFE_OK;
}

int pfe_ctx_check_async(pfe_ctx *c) {
    if (!c) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(c, cudaSetDevice(c->device));
    int flag = 0;
    PFE_CUDA(c, cudaMemcpyAsync(&flag, c->async_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PFE_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!flag) return PFE_OK;
    PFE_CUDA(c, cudaMemsetAsync(c->async_err, 0, sizeof(int), c->stream));
    if (flag & PFE_ASYNC_PEER_TIMEOUT)
        return pfe_fail(c, PFE_ERR_INVALID_ARG, "peer_wait: a neighbour's halo rows did not arrive before the timeout");
    return pfe_fail(c, PFE_ERR_INVALID_ARG, "warp_band: the source row window does not cover the warp's reach");
}

// -- device memory shared between the processes of one node (one process per GPU) ------------------------------
// The halo rows of a canvas split across GPUs travel as the producing kernel's own stores into the neighbour's
// buffer (pfe_dev_flatten_peer).  That needs the neighbour's allocation mapped here: CUDA IPC, which also enables
// peer access between the two devices (NVLink on an NVSwitch box).
int pfe_peer_alloc(pfe_ctx *c, size_t bytes, void **dptr, uint8_t handle[64]) {
    if (!c) return PFE_ERR_INVALID_ARG;
    if (!dptr || !handle || !bytes) return pfe_fail(c, PFE_ERR_INVALID_ARG, "peer_alloc: bad args");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    PFE_CUDA(c, cudaSetDevice(c->device));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return pfe_fail(c, PFE_ERR_OOM, "peer_alloc: cudaMalloc", e); }
    cudaIpcMemHandle_t hd;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hd, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
        return pfe_fail(c, PFE_ERR_CUDA, "peer_alloc: cudaIpcGetMemHandle", e);
    }
    memcpy(handle, &hd, 64);
    *dptr = p;
    return PFE_OK;
}

int pfe_peer_open(pfe_ctx *c, const uint8_t handle[64], void **dptr) {
    if (!c) return PFE_ERR_INVALID_ARG;
    if (!dptr || !handle) return pfe_fail(c, PFE_ERR_INVALID_ARG, "peer_open: bad args");
    PFE_CUDA(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return pfe_fail(c, PFE_ERR_CUDA, "peer_open: cudaIpcOpenMemHandle", e); }
    *dptr = p;
    return PFE_OK;
}

int pfe_peer_close(pfe_ctx *c, void *dptr) {
    if (!c) return PFE_ERR_INVALID_ARG;
    if (!dptr) return PFE_OK;
    PFE_CUDA(c, cudaSetDevice(c->device));
    PFE_CUDA(c, cudaIpcCloseMemHandle(dptr));
    return PFE_OK;
}

int pfe_peer_free(pfe_ctx *c, void *dptr) {
    if (!c) return PFE_ERR_INVALID_ARG;
    if (!dptr) return PFE_OK;
    PFE_CUDA(c, cudaSetDevice(c->device));
    PFE_CUDA(c, cudaDeviceSynchronize());
    PFE_CUDA(c, cudaFree(dptr));
    return PFE_OK;
}

const char *pfe_last_error(const pfe_ctx *c) { return c ? c->err.c_str() : "null context"; }
uint64_t pfe_ctx_launch_count(const pfe_ctx *c) { return c ? c->launches : 0; }

int pfe_host_alloc(size_t bytes, void **out) {
    if (!out) return PFE_ERR_INVALID_ARG;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return PFE_ERR_OOM; }
    return PFE_OK;
}
int pfe_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? PFE_OK : PFE_ERR_CUDA; }

int pfe_dev_alloc(pfe_ctx *c, size_t bytes, void **out) {
    if (!c || !out) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(c, cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return pfe_fail(c, PFE_ERR_OOM, "cudaMalloc", e); }
    return PFE_OK;
}
int pfe_dev_free(pfe_ctx *c, void *p) {
    if (!c) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(c, cudaSetDevice(c->device));
    PFE_CUDA(c, cudaStreamSynchronize(c->stream));
    PFE_CUDA(c, cudaFree(p));
    return PFE_OK;
}
int pfe_dev_upload(pfe_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || (bytes && (!dst || !src))) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(c, cudaSetDevice(c->device));
    PFE_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return PFE_OK;
}
int pfe_dev_download(pfe_ctx *c, void *dst, const void *src, size_t bytes) {
    if (!c || (bytes && (!dst || !src))) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(c, cudaSetDevice(c->device));
    PFE_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    PFE_CUDA(c, cudaStreamSynchronize(c->stream));
    return PFE_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Host tier.  One image in, one image out, optional selection mask.
// ---------------------------------------------------------------------------------------------
namespace {

struct Staged {
    uint8_t *src = nullptr, *dst = nullptr, *mask = nullptr;
};

int stage_in(pfe_ctx *c, const uint8_t *src, uint32_t w, uint32_t h, const uint8_t *mask, Staged *s) {
    if (!c || !src || w == 0 || h == 0) return c ? pfe_fail(c, PFE_ERR_INVALID_ARG, "null image or zero size") : PFE_ERR_INVALID_ARG;
    PFE_CUDA(c, cudaSetDevice(c->device));
    size_t n4 = (size_t)w * h * 4;
    void *a, *b, *m = nullptr;
    // one arena for src | dst | mask so a single grow covers the call
    size_t total = 2 * ((n4 + 255) & ~size_t(255)) + (mask ? (size_t)w * h : 0);
    PFE_TRY(pfe_scratch(c, PFE_SCRATCH_A, total, &a));
    b = (char *)a + ((n4 + 255) & ~size_t(255));
    if (mask) m = (char *)b + ((n4 + 255) & ~size_t(255));
    PFE_CUDA(c, cudaMemcpyAsync(a, src, n4, cudaMemcpyHostToDevice, c->stream));
    if (mask) PFE_CUDA(c, cudaMemcpyAsync(m, mask, (size_t)w * h, cudaMemcpyHostToDevice, c->stream));
    s->src = (uint8_t *)a;
    s->dst = (uint8_t *)b;
    s->mask = (uint8_t *)m;
    return PFE_OK;
}

int stage_out(pfe_ctx *c, const Staged &s, uint32_t w, uint32_t h, uint8_t *dst) {
    if (!dst) return pfe_fail(c, PFE_ERR_INVALID_ARG, "null dst");
    PFE_CUDA(c, cudaMemcpyAsync(dst, s.dst, (size_t)w * h * 4, cudaMemcpyDeviceToHost, c->stream));
    PFE_CUDA(c, cudaStreamSynchronize(c->stream));
    return PFE_OK;
}

}  // namespace

#define HOST_TIER(call_dev)                       \
    Staged s;                                     \
    PFE_TRY(stage_in(ctx, src, w, h, mask, &s));  \
    PFE_TRY(call_dev);                            \
    return stage_out(ctx, s, w, h, dst);

extern "C" {

int pfe_gaussian_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float sigma,
                      const uint8_t *mask, uint8_t *dst, uint32_t flags) {
    HOST_TIER(pfe_dev_gaussian_blur(ctx, s.src, w, h, sigma, s.mask, s.dst, flags));
}
int pfe_box_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius,
                 const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_box_blur(ctx, s.src, w, h, radius, s.mask, s.dst));
}
int pfe_motion_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float angle_deg,
                    float distance, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_motion_blur(ctx, s.src, w, h, angle_deg, distance, s.mask, s.dst));
}
int pfe_median(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius,
               const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_median(ctx, s.src, w, h, radius, s.mask, s.dst));
}
int pfe_sharpen(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, float radius,
                const uint8_t *mask, uint8_t *dst, uint32_t flags) {
    HOST_TIER(pfe_dev_sharpen(ctx, s.src, w, h, amount, radius, s.mask, s.dst, flags));
}
int pfe_vignette(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount,
                 float softness, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_vignette(ctx, s.src, w, h, amount, softness, s.mask, s.dst));
}

int pfe_glow(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius, float intensity,
             const uint8_t *mask, uint8_t *dst, uint32_t flags) {
    HOST_TIER(pfe_dev_glow(ctx, s.src, w, h, radius, intensity, s.mask, s.dst, flags));
}
int pfe_pixelate(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t block_size,
                 const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_pixelate(ctx, s.src, w, h, block_size, s.mask, s.dst));
}
int pfe_bulge(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, float ox, float oy,
              const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_bulge(ctx, s.src, w, h, amount, ox, oy, s.mask, s.dst));
}
int pfe_twist(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float angle_deg, float ox, float oy,
              const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_twist(ctx, s.src, w, h, angle_deg, ox, oy, s.mask, s.dst));
}
int pfe_add_noise(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, int noise_type,
                  int monochrome, uint32_t seed, float scale, uint32_t octaves, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_add_noise(ctx, s.src, w, h, amount, noise_type, monochrome, seed, scale, octaves, s.mask, s.dst));
}
int pfe_reduce_noise(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float strength, uint32_t radius,
                     const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_reduce_noise(ctx, s.src, w, h, strength, radius, s.mask, s.dst));
}

// the rest of src/ops/effects/ (effects3.cu)
int pfe_ink(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float edge_strength, float threshold,
            const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_ink(ctx, s.src, w, h, edge_strength, threshold, s.mask, s.dst));
}
int pfe_oil_painting(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius,
                     uint32_t levels, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_oil_painting(ctx, s.src, w, h, radius, levels, s.mask, s.dst));
}
int pfe_color_filter(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const uint8_t *color,
                     float intensity, int mode, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_color_filter(ctx, s.src, w, h, color, intensity, mode, s.mask, s.dst));
}
int pfe_contours(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float scale, float frequency,
                 float line_width, const uint8_t *color, uint32_t seed, uint32_t octaves, float blend,
                 const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_contours(ctx, s.src, w, h, scale, frequency, line_width, color, seed, octaves, blend, s.mask, s.dst));
}
int pfe_crystallize(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float cell_size,
                    uint32_t seed, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_crystallize(ctx, s.src, w, h, cell_size, seed, s.mask, s.dst));
}
int pfe_dents(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float scale, float amount,
              uint32_t seed, uint32_t octaves, float roughness, int pinch, int wrap, const uint8_t *mask,
              uint8_t *dst) {
    HOST_TIER(pfe_dev_dents(ctx, s.src, w, h, scale, amount, seed, octaves, roughness, pinch, wrap, s.mask, s.dst));
}
int pfe_halftone(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float dot_size, float angle_deg,
                 int shape, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_halftone(ctx, s.src, w, h, dot_size, angle_deg, shape, s.mask, s.dst));
}
int pfe_bokeh_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius,
                   const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_bokeh_blur(ctx, s.src, w, h, radius, s.mask, s.dst));
}
int pfe_zoom_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float center_x, float center_y,
                  float strength, uint32_t samples, const float *tint_rgba, float tint_strength,
                  const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_zoom_blur(ctx, s.src, w, h, center_x, center_y, strength, samples, tint_rgba, tint_strength, s.mask, s.dst));
}
int pfe_grid(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t cell_w, uint32_t cell_h,
             uint32_t line_width, const uint8_t *color, int style, float opacity, const uint8_t *mask,
             uint8_t *dst) {
    HOST_TIER(pfe_dev_grid(ctx, s.src, w, h, cell_w, cell_h, line_width, color, style, opacity, s.mask, s.dst));
}
int pfe_canvas_border(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t width,
                      const uint8_t *color, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_canvas_border(ctx, s.src, w, h, width, color, s.mask, s.dst));
}
int pfe_drop_shadow(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int32_t offset_x,
                    int32_t offset_y, float blur_radius, int widen_radius, const uint8_t *color,
                    float opacity, const uint8_t *mask, uint8_t *dst, uint32_t flags) {
    HOST_TIER(pfe_dev_drop_shadow(ctx, s.src, w, h, offset_x, offset_y, blur_radius, widen_radius, color, opacity, s.mask, s.dst, flags));
}
int pfe_outline(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t width,
                const uint8_t *color, int mode, int anti_alias, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_outline(ctx, s.src, w, h, width, color, mode, anti_alias, s.mask, s.dst));
}
int pfe_pixel_drag(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t seed, float amount,
                   uint32_t distance, float direction, const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_pixel_drag(ctx, s.src, w, h, seed, amount, distance, direction, s.mask, s.dst));
}
int pfe_rgb_displace(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const int32_t *offsets,
                     const uint8_t *mask, uint8_t *dst) {
    HOST_TIER(pfe_dev_rgb_displace(ctx, s.src, w, h, offsets, s.mask, s.dst));
}

int pfe_adjust(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const pfe_adjust_desc *d,
               const uint8_t *mask, const uint8_t *occupancy, uint8_t *dst) {
    if (!ctx || !d) return PFE_ERR_INVALID_ARG;
    Staged s;
    PFE_TRY(stage_in(ctx, src, w, h, mask, &s));
    void *occ = nullptr;
    if (occupancy) {
        size_t nb = (size_t)pfe_div_up(w, PFE_CHUNK_SIZE) * pfe_div_up(h, PFE_CHUNK_SIZE);
        PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, nb, &occ));
        PFE_CUDA(ctx, cudaMemcpyAsync(occ, occupancy, nb, cudaMemcpyHostToDevice, ctx->stream));
    }
    PFE_TRY(pfe_dev_adjust(ctx, s.src, w, h, d, s.mask, (const uint8_t *)occ, s.dst));
    return stage_out(ctx, s, w, h, dst);
}

int pfe_channel_minmax(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const uint8_t *mask,
                       uint8_t out[6]) {
    Staged s;
    PFE_TRY(stage_in(ctx, src, w, h, mask, &s));
    return pfe_dev_channel_minmax(ctx, s.src, w, h, s.mask, out);
}

int pfe_warp_displacement(pfe_ctx *ctx, const uint8_t *src, uint32_t sw, uint32_t sh, const float *disp,
                          uint32_t w, uint32_t h, uint8_t *dst) {
    if (!ctx || !src || !disp || !dst || !sw || !sh || !w || !h) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t sb = (size_t)sw * sh * 4, db = (size_t)w * h * 4, fb = (size_t)w * h * 8;
    void *a, *b, *f;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_A, sb, &a));
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, db, &b));
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, fb, &f));
    PFE_CUDA(ctx, cudaMemcpyAsync(a, src, sb, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaMemcpyAsync(f, disp, fb, cudaMemcpyHostToDevice, ctx->stream));
    PFE_TRY(pfe_dev_warp_displacement(ctx, (uint8_t *)a, sw, sh, (float *)f, w, h, (uint8_t *)b));
    PFE_CUDA(ctx, cudaMemcpyAsync(dst, b, db, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}

int pfe_warp_displacement_region(pfe_ctx *ctx, const uint8_t *src, uint32_t sw, uint32_t sh, const float *disp,
                                 const uint8_t *prev, const int32_t rect[4], uint32_t w, uint32_t h, uint8_t *dst) {
    if (!ctx || !src || !disp || !prev || !rect || !dst || !sw || !sh || !w || !h) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t sb = (size_t)sw * sh * 4, db = (size_t)w * h * 4, fb = (size_t)w * h * 8;
    void *a, *b, *f;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_A, sb, &a));
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, db, &b));
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, fb, &f));
    PFE_CUDA(ctx, cudaMemcpyAsync(a, src, sb, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaMemcpyAsync(b, prev, db, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaMemcpyAsync(f, disp, fb, cudaMemcpyHostToDevice, ctx->stream));
    PFE_TRY(pfe_dev_warp_displacement_region(ctx, (uint8_t *)a, sw, sh, (float *)f, (uint8_t *)b, rect, w, h, (uint8_t *)b));
    PFE_CUDA(ctx, cudaMemcpyAsync(dst, b, db, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}

int pfe_mesh_displacement(pfe_ctx *ctx, const float *orig, const float *def, uint32_t cols, uint32_t rows,
                          uint32_t w, uint32_t h, float *out) {
    if (!ctx || !def || !out || !w || !h) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t fb = (size_t)w * h * 8;
    void *f;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, fb, &f));
    PFE_TRY(pfe_dev_mesh_displacement(ctx, orig, def, cols, rows, w, h, (float *)f));
    PFE_CUDA(ctx, cudaMemcpyAsync(out, f, fb, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}

int pfe_mesh_warp(pfe_ctx *ctx, const uint8_t *src, uint32_t sw, uint32_t sh, const float *orig,
                  const float *def, uint32_t cols, uint32_t rows, uint32_t w, uint32_t h, uint8_t *dst) {
    if (!ctx || !src || !dst || !def || !sw || !sh || !w || !h) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t sb = (size_t)sw * sh * 4, db = (size_t)w * h * 4;
    void *a, *b;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_A, sb, &a));
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, db, &b));
    PFE_CUDA(ctx, cudaMemcpyAsync(a, src, sb, cudaMemcpyHostToDevice, ctx->stream));
    PFE_TRY(pfe_dev_mesh_warp(ctx, (uint8_t *)a, sw, sh, orig, def, cols, rows, w, h, 0, h, (uint8_t *)b));
    PFE_CUDA(ctx, cudaMemcpyAsync(dst, b, db, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}

int pfe_liquify(pfe_ctx *ctx, float *field, uint32_t w, uint32_t h, int kind, float cx, float cy,
                float radius, float strength, float a0, float a1, int32_t bbox[4]) {
    if (!ctx || !field || !w || !h) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t fb = (size_t)w * h * 8;
    void *f;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, fb, &f));
    PFE_CUDA(ctx, cudaMemcpyAsync(f, field, fb, cudaMemcpyHostToDevice, ctx->stream));
    PFE_TRY(pfe_dev_liquify(ctx, (float *)f, w, h, kind, cx, cy, radius, strength, a0, a1, bbox));
    PFE_CUDA(ctx, cudaMemcpyAsync(field, f, fb, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}

int pfe_brush_stamps(pfe_ctx *ctx, uint8_t *image, uint32_t w, uint32_t h, const pfe_brush_desc *brush,
                     const float *centres, uint32_t n, const uint8_t *sel) {
    if (!ctx || !image || !brush || (n && !centres) || !w || !h) return PFE_ERR_INVALID_ARG;
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t db = (size_t)w * h * 4;
    void *a, *m = nullptr;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_A, db, &a));
    PFE_CUDA(ctx, cudaMemcpyAsync(a, image, db, cudaMemcpyHostToDevice, ctx->stream));
    if (sel) {
        PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, (size_t)w * h, &m));
        PFE_CUDA(ctx, cudaMemcpyAsync(m, sel, (size_t)w * h, cudaMemcpyHostToDevice, ctx->stream));
    }
    PFE_TRY(pfe_dev_brush_stamps(ctx, (uint8_t *)a, w, h, brush, centres, n, (uint8_t *)m));
    PFE_CUDA(ctx, cudaMemcpyAsync(image, a, db, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}

// ---- flatten host tier --------------------------------------------------------------------------
// The call is PCIe-bound (4 bytes per layer per pixel in, 4 out), so it is pipelined in 64-row-
// aligned bands on three streams: while band k+1 of every layer is uploading (copy_stream), band k
// is flattened and H-blurred (ctx->stream); a band's V pass runs as soon as the H rows within
// +-radius of it exist, and its result starts downloading (d2h_stream) while later bands are still
// in flight. Only the last band's compute and download are exposed.
static int flatten_pipeline(pfe_ctx *ctx, const pfe_layer_desc *layers, uint32_t n, uint32_t w, uint32_t h,
                            const uint8_t *active, bool blur, float sigma, uint32_t flags, uint8_t *dst) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if ((!layers && n) || !w || !h || !dst) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "flatten: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n4 = (size_t)w * h * 4, n1 = (size_t)w * h;
    const size_t a4 = (n4 + 255) & ~size_t(255), a1 = (n1 + 255) & ~size_t(255);
    size_t total = a4;  // flattened image
    for (uint32_t i = 0; i < n; i++) {
        if (!layers[i].visible || layers[i].kind != PFE_LAYER_RASTER) continue;
        if (!layers[i].rgba) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "flatten: raster layer without pixels");
        total += a4 + (layers[i].mask ? a1 : 0);
    }
    const uint32_t chunks_x = pfe_div_up(w, PFE_CHUNK_SIZE);
    const size_t nb = (size_t)chunks_x * pfe_div_up(h, PFE_CHUNK_SIZE);
    if (active) total += (nb + 255) & ~size_t(255);
    void *base;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_A, total, &base));
    char *cur = (char *)base;
    uint8_t *flat = (uint8_t *)cur;
    cur += a4;
    std::vector<pfe_layer_desc> dl(layers, layers + n);
    for (uint32_t i = 0; i < n; i++) {
        pfe_layer_desc &L = dl[i];
        if (!L.visible || L.kind != PFE_LAYER_RASTER) { L.rgba = nullptr; L.mask = nullptr; continue; }
        L.rgba = (uint8_t *)cur;
        cur += a4;
        if (layers[i].mask) { L.mask = (uint8_t *)cur; cur += a1; }
    }
    uint8_t *active_dev = nullptr;
    if (active) active_dev = (uint8_t *)cur;
    uint8_t *out = flat;
    float *mid = nullptr;
    int radius = 0;
    if (blur) {
        void *b, *m;
        PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_B, n4, &b));
        PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_F32, (size_t)w * h * 16, &m));
        out = (uint8_t *)b;
        mid = (float *)m;
        radius = pfe_gauss_radius(sigma);
        if (radius > 4000) return pfe_fail(ctx, PFE_ERR_UNSUPPORTED, "gaussian: sigma too large");
    }
    // band height: ~8 bands, 64-row aligned; small images go through in one band
    uint32_t band_h = h;
    static const uint32_t want_bands = getenv("PFE_PIPE_BANDS") ? (uint32_t)std::max(1, atoi(getenv("PFE_PIPE_BANDS"))) : 8u;  // tuning aid
    if (n4 >= (size_t)8 << 20) band_h = std::max<uint32_t>(PFE_CHUNK_SIZE, ((h / want_bands + PFE_CHUNK_SIZE - 1) / PFE_CHUNK_SIZE) * PFE_CHUNK_SIZE);
    // Band list: regular bands, then the last stretch in shrinking pieces (down to one chunk row), because
    // what remains exposed after the last upload is the compute and download of the final band(s) only.
    std::vector<uint32_t> band_y0, band_rows;
    for (uint32_t y = 0; y < h;) {
        uint32_t rows = std::min(band_h, h - y);
        if (band_h < h && h - y <= band_h && rows > PFE_CHUNK_SIZE)
            rows = std::max<uint32_t>(PFE_CHUNK_SIZE, ((rows / 2 + PFE_CHUNK_SIZE - 1) / PFE_CHUNK_SIZE) * PFE_CHUNK_SIZE);
        band_y0.push_back(y);
        band_rows.push_back(rows);
        y += rows;
    }
    const uint32_t nbands = (uint32_t)band_y0.size();

    // the upload stream must not overwrite buffers that earlier work on ctx->stream still reads
    PFE_CUDA(ctx, cudaEventRecord(ctx->ev_copy, ctx->stream));
    PFE_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy, 0));
    PFE_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_copy, 0));
    if (active) PFE_CUDA(ctx, cudaMemcpyAsync(active_dev, active, nb, cudaMemcpyHostToDevice, ctx->copy_stream));

    // are all layer (and mask) buffers page-locked?
    bool all_pinned = getenv("PFE_PIPE_NO_BATCH") == nullptr;
    for (uint32_t i = 0; i < n && all_pinned; i++) {
        if (!dl[i].rgba) continue;
        for (const void *hp : {(const void *)layers[i].rgba, (const void *)layers[i].mask}) {
            if (!hp) continue;
            cudaPointerAttributes pa;
            if (cudaPointerGetAttributes(&pa, hp) != cudaSuccess || pa.type != cudaMemoryTypeHost) { cudaGetLastError(); all_pinned = false; }
        }
    }
    int rc = PFE_OK;
    std::vector<cudaEvent_t> events;
    auto new_event = [&]() -> cudaEvent_t {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        events.push_back(e);
        return e;
    };
    static const bool dbg = getenv("PFE_PIPE_DEBUG") != nullptr;  // tuning aid: where the call's time goes
    cudaEvent_t t0 = nullptr, t_up = nullptr, t_cmp = nullptr, t_dn = nullptr;
    if (dbg) {
        cudaEventCreate(&t0); cudaEventCreate(&t_up); cudaEventCreate(&t_cmp); cudaEventCreate(&t_dn);
        cudaEventRecord(t0, ctx->copy_stream);
    }
    uint32_t v_next = 0;  // first band whose V pass (or download) has not been issued yet
    // Inside the band loop a CUDA error must not return: copies on copy_stream / d2h_stream may still be touching the
    // caller's host buffers and the per-band events would leak.  Errors set rc and fall through to the common tail,
    // which synchronises the three streams and destroys the events.
    auto cuda_ok = [&](cudaError_t e, const char *what) -> bool {
        if (e == cudaSuccess) return true;
        if (rc == PFE_OK) rc = pfe_fail(ctx, PFE_ERR_CUDA, what, e);
        return false;
    };
    auto download_band = [&](uint32_t k) -> int {
        const uint32_t y0 = band_y0[k], rows = band_rows[k];
        cudaEvent_t e = new_event();
        if (!e) return pfe_fail(ctx, PFE_ERR_CUDA, "cudaEventCreate");
        if (!cuda_ok(cudaEventRecord(e, ctx->stream), "cudaEventRecord") ||
            !cuda_ok(cudaStreamWaitEvent(ctx->d2h_stream, e, 0), "cudaStreamWaitEvent") ||
            !cuda_ok(cudaMemcpyAsync(dst + (size_t)y0 * w * 4, out + (size_t)y0 * w * 4, (size_t)rows * w * 4,
                                     cudaMemcpyDeviceToHost, ctx->d2h_stream), "cudaMemcpyAsync (download)"))
            return rc;
        return PFE_OK;
    };
    for (uint32_t b = 0; b < nbands && rc == PFE_OK; b++) {
        const uint32_t y0 = band_y0[b], rows = band_rows[b];
        const size_t off4 = (size_t)y0 * w * 4, off1 = (size_t)y0 * w;
        std::vector<pfe_layer_desc> bl(dl);
        std::vector<void *> cp_dst, cp_src;
        std::vector<size_t> cp_len;
        for (uint32_t i = 0; i < n; i++) {
            if (!dl[i].rgba) continue;
            cp_dst.push_back((void *)(dl[i].rgba + off4)); cp_src.push_back((void *)(layers[i].rgba + off4)); cp_len.push_back((size_t)rows * w * 4);
            bl[i].rgba = dl[i].rgba + off4;
            if (dl[i].mask) {
                cp_dst.push_back((void *)(dl[i].mask + off1)); cp_src.push_back((void *)(layers[i].mask + off1)); cp_len.push_back((size_t)rows * w);
                bl[i].mask = dl[i].mask + off1;
            }
        }
        // One batched submission per band when every source is pinned (the copy engine then runs the band's
        // copies back to back instead of paying a gap per cudaMemcpyAsync); individual copies otherwise.
        bool batched = false;
        if (all_pinned && cp_dst.size() > 1) {
            cudaMemcpyAttributes attr;
            memset(&attr, 0, sizeof(attr));
            attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
            size_t attr_idx = 0, fail = 0;
            if (cudaMemcpyBatchAsync(cp_dst.data(), cp_src.data(), cp_len.data(), cp_dst.size(), &attr, &attr_idx, 1, &fail, ctx->copy_stream) == cudaSuccess)
                batched = true;
            else {
                cudaGetLastError();
                all_pinned = false;  // not supported here: stay with plain copies for the rest of the call
            }
        }
        if (!batched)
            for (size_t c = 0; c < cp_dst.size() && rc == PFE_OK; c++)
                cuda_ok(cudaMemcpyAsync(cp_dst[c], cp_src[c], cp_len[c], cudaMemcpyHostToDevice, ctx->copy_stream), "cudaMemcpyAsync (upload)");
        if (rc != PFE_OK) break;
        cudaEvent_t up = new_event();
        if (!up) { rc = pfe_fail(ctx, PFE_ERR_CUDA, "cudaEventCreate"); break; }
        if (!cuda_ok(cudaEventRecord(up, ctx->copy_stream), "cudaEventRecord") ||
            !cuda_ok(cudaStreamWaitEvent(ctx->stream, up, 0), "cudaStreamWaitEvent"))
            break;
        rc = pfe_dev_flatten(ctx, bl.data(), n, w, rows, active_dev ? active_dev + (size_t)(y0 / PFE_CHUNK_SIZE) * chunks_x : nullptr,
                             flat + off4);
        if (rc != PFE_OK) break;
        if (!blur) { rc = download_band(b); continue; }
        rc = pfe_gauss_h_rows(ctx, flat, mid, w, h, y0, rows, sigma, flags);
        const uint32_t h_done = y0 + rows;
        while (rc == PFE_OK && v_next < nbands) {
            const uint32_t vy0 = band_y0[v_next], vrows = band_rows[v_next];
            if (h_done < h && vy0 + vrows + (uint32_t)radius > h_done) break;  // its lower halo is not blurred yet
            rc = pfe_gauss_v_rows(ctx, mid, out, w, h, vy0, vrows, sigma, flags);
            if (rc == PFE_OK) rc = download_band(v_next);
            v_next++;
        }
    }
    if (dbg) { cudaEventRecord(t_up, ctx->copy_stream); cudaEventRecord(t_cmp, ctx->stream); cudaEventRecord(t_dn, ctx->d2h_stream); }
    cudaError_t e1 = cudaStreamSynchronize(ctx->copy_stream), e2 = cudaStreamSynchronize(ctx->stream),
                e3 = cudaStreamSynchronize(ctx->d2h_stream);
    if (dbg) {
        float a = 0, b = 0, c = 0;
        cudaEventElapsedTime(&a, t0, t_up); cudaEventElapsedTime(&b, t0, t_cmp); cudaEventElapsedTime(&c, t0, t_dn);
        fprintf(stderr, "[pfe pipe] uploads done %.2f ms, compute done %.2f ms, downloads done %.2f ms (bands %u)\n", a, b, c, nbands);
        cudaEventDestroy(t0); cudaEventDestroy(t_up); cudaEventDestroy(t_cmp); cudaEventDestroy(t_dn);
    }
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    if (rc != PFE_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return pfe_fail(ctx, PFE_ERR_CUDA, "flatten pipeline", e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3));
    return PFE_OK;
}

int pfe_flatten(pfe_ctx *ctx, const pfe_layer_desc *layers, uint32_t n, uint32_t w, uint32_t h,
                const uint8_t *active, uint8_t *dst) {
    return flatten_pipeline(ctx, layers, n, w, h, active, false, 0.0f, 0, dst);
}

int pfe_flatten_gaussian(pfe_ctx *ctx, const pfe_layer_desc *layers, uint32_t n, uint32_t w, uint32_t h,
                         const uint8_t *active, float sigma, uint8_t *dst, uint32_t flags) {
    return flatten_pipeline(ctx, layers, n, w, h, active, true, sigma, flags, dst);
}

}  // extern "C"
