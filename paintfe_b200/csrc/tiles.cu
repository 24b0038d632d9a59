// Tile-native compositing: layers stay what TiledImage makes them - a table of
// optional 64x64 RGBA chunks (src/canvas/tiled_image.rs:2-7) - on the host side of the ABI *and* on the
// device.  The flatten walks chunk tables the way CanvasState::composite_viewport does
// (src/canvas/canvas_state.rs:529-600): a chunk no visible raster layer populates is never touched, a
// layer without a chunk at (cx, cy) is skipped for that chunk, a mask without a chunk conceals nothing.
// Only populated chunks cross PCIe, so a sparse document costs what it contains, not w*h per layer.
//
// The per-pixel arithmetic is blend.cuh's, shared with the dense kernel; results are identical to the
// dense flatten of the same pixels (an absent chunk is 64x64 transparent pixels, and a transparent top
// pixel returns the base before any mode logic, canvas_state.rs:1253).
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "blend.cuh"

// A TiledImage on the device.  Chunks live in the context's pool (pfe_ctx::chunks) and are reference counted on the
// host: `slot[c]` is the pool slot of chunk c or -1; a clone shares every slot, and whoever writes a shared chunk first
// gets a private copy (Arc::make_mut, tiled_image.rs:330, :868).  `table` is what the kernels walk: the slot's address
// where the chunk is populated, null where it is not (a slot can be assigned but unpopulated after from_flat, whose
// alpha scan runs on the device).
struct pfe_tiled {
    pfe_ctx *owner = nullptr;
    uint32_t w = 0, h = 0, chunks_x = 0, chunks_y = 0;
    std::vector<int32_t> slot;        // host: pool slot per chunk, -1 = none
    const uint8_t **table = nullptr;  // device array: chunk pointer or null
    uint8_t *occupancy = nullptr;     // device bytes, 1 = populated
};

namespace {

constexpr int kMaxLayers = 32, kMaxAdj = 8;
constexpr size_t kChunkBytes = (size_t)PFE_CHUNK_SIZE * PFE_CHUNK_SIZE * 4;

struct TileLayer {
    const uint8_t *const *chunks;       // device table (raster layers)
    const uint8_t *const *mask_chunks;  // device table or null
    float opacity;
    uint8_t blend, kind, adj_slot, _pad;
};
struct TileParams {
    TileLayer layers[kMaxLayers];
    float adj[kMaxAdj][16];
    const uint8_t *active;  // one byte per chunk: some visible raster layer populates it
    uint8_t *dst;           // flat w*h RGBA8
    uint32_t n_layers, w, h, chunks_x, n_chunks, init_from_dst, vec_store;
    PackedConsts pc;
};

// active |= "this layer has a chunk here" (canvas_state.rs:529-550)
__global__ void tile_active_kernel(const uint8_t *const *table, uint8_t *active, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && table[i] != nullptr) active[i] = 1;
}

__global__ void __launch_bounds__(256, 3) flatten_tiles_kernel(const __grid_constant__ TileParams P) {
    __shared__ const uint8_t *s_px[kMaxLayers], *s_mask[kMaxLayers];
    __shared__ int s_list[kMaxLayers], s_n;
    blend_lut_init();
    const Lut lut = make_lut(P.pc);
    for (uint32_t chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        __syncthreads();  // table ready (first pass) / previous chunk's list no longer in use
        const uint32_t cy = chunk / P.chunks_x, cx = chunk - cy * P.chunks_x;
        const bool active = P.active[chunk] != 0;
        if (threadIdx.x < 32) {
            // the layers that do something in this chunk, bottom to top: raster layers with a chunk here and
            // adjustment layers (:579-600). One lane per layer (kMaxLayers == 32) so the table reads of all
            // layers are in flight together; a ballot compacts the survivors in layer order.
            const uint32_t li = threadIdx.x;
            const uint8_t *p = nullptr, *m = nullptr;
            bool use = false;
            if (active && li < P.n_layers) {
                const TileLayer &L = P.layers[li];
                if (L.kind != PFE_LAYER_RASTER) use = true;
                else if ((p = L.chunks[chunk]) != nullptr) {
                    use = true;
                    m = L.mask_chunks ? L.mask_chunks[chunk] : nullptr;
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, use);
            if (use) {
                s_list[__popc(b & ((1u << li) - 1u))] = (int)li;
                s_px[li] = p;
                s_mask[li] = m;
            }
            if (li == 0) s_n = __popc(b);
        }
        __syncthreads();
        const int n = s_n;
#pragma unroll 1
        for (int it = 0; it < 4; it++) {
            const uint32_t row = it * 16 + (threadIdx.x >> 4), col = (threadIdx.x & 15) * 4;
            const uint32_t gx = cx * PFE_CHUNK_SIZE + col, gy = cy * PFE_CHUNK_SIZE + row;
            if (gx >= P.w || gy >= P.h) continue;  // the chunk's zero padding past the canvas edge
            const uint32_t off = (row * PFE_CHUNK_SIZE + col) * 4;
            uint8_t *out = P.dst + ((size_t)gy * P.w + gx) * 4;
            const uint32_t valid = min(4u, P.w - gx);
            uint32_t acc[4] = {0u, 0u, 0u, 0u};
            if (active) {
                if (P.init_from_dst) {
                    if (P.vec_store) { const uint4 v = *reinterpret_cast<const uint4 *>(out); acc[0] = v.x; acc[1] = v.y; acc[2] = v.z; acc[3] = v.w; }
                    else for (uint32_t k = 0; k < valid; k++) acc[k] = reinterpret_cast<const uint32_t *>(out)[k];
                }
                // one layer ahead: the next raster chunk's 16 bytes are requested before this layer's math
                uint4 pend = make_uint4(0, 0, 0, 0);
                if (n > 0 && P.layers[s_list[0]].kind == PFE_LAYER_RASTER) pend = __ldg(reinterpret_cast<const uint4 *>(s_px[s_list[0]] + off));
                for (int i = 0; i < n; i++) {
                    const int li = s_list[i];
                    const TileLayer &L = P.layers[li];
                    const uint4 cur = pend;
                    if (i + 1 < n && P.layers[s_list[i + 1]].kind == PFE_LAYER_RASTER)
                        pend = __ldg(reinterpret_cast<const uint4 *>(s_px[s_list[i + 1]] + off));
                    if (L.kind != PFE_LAYER_RASTER) {
#pragma unroll
                        for (int k = 0; k < 4; k++) acc[k] = adj_px(acc[k], L.kind, P.adj[L.adj_slot], L.opacity);
                        continue;
                    }
                    uint32_t top[4] = {cur.x, cur.y, cur.z, cur.w};
                    if (s_mask[li]) {  // :660-665 - the mask is a TiledImage; its alpha conceals
                        const uint4 m = __ldg(reinterpret_cast<const uint4 *>(s_mask[li] + off));
                        const uint32_t mv[4] = {m.x >> 24, m.y >> 24, m.z >> 24, m.w >> 24};
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (mv[k] > 0) top[k] = (top[k] & 0x00FFFFFFu) | ((((top[k] >> 24) * (255u - mv[k])) / 255u) << 24);
                    }
                    if (__all_sync(__activemask(), (top[0] | top[1] | top[2] | top[3]) <= 0x00FFFFFFu)) continue;
                    blend_k<4>(acc, top, L.blend, L.opacity, pfe_clampf(L.opacity, 0.0f, 1.0f), lut);
                }
            } else if (P.init_from_dst) {
                continue;  // an inactive chunk was zeroed by the first launch of the chain
            }
            if (P.vec_store) *reinterpret_cast<uint4 *>(out) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
            else for (uint32_t k = 0; k < valid; k++) reinterpret_cast<uint32_t *>(out)[k] = acc[k];
        }
    }
}

// TiledImage::from_rgba_image on the device (tiled_image.rs:50-104): one CTA per chunk copies the chunk
// into its pool slot (zero padded past the canvas edge) and records whether any pixel has alpha != 0.
// On entry table[chunk] is the slot assigned to the chunk; on exit it is that slot or null.
__global__ void __launch_bounds__(256) tiles_from_flat_kernel(const uint32_t *flat, uint32_t w, uint32_t h, uint32_t chunks_x,
                                                              const uint8_t **table, uint8_t *occupancy) {
    const uint32_t chunk = blockIdx.x, cy = chunk / chunks_x, cx = chunk - cy * chunks_x;
    uint32_t *slot = reinterpret_cast<uint32_t *>(const_cast<uint8_t *>(table[chunk]));
    bool any = false;
    for (uint32_t i = threadIdx.x; i < PFE_CHUNK_SIZE * PFE_CHUNK_SIZE; i += blockDim.x) {
        const uint32_t ly = i / PFE_CHUNK_SIZE, lx = i % PFE_CHUNK_SIZE;
        const uint32_t gx = cx * PFE_CHUNK_SIZE + lx, gy = cy * PFE_CHUNK_SIZE + ly;
        const uint32_t v = (gx < w && gy < h) ? flat[(size_t)gy * w + gx] : 0u;
        slot[i] = v;
        any = any || (v >> 24) != 0;
    }
    const int has = __syncthreads_or(any ? 1 : 0);
    if (threadIdx.x == 0) {
        occupancy[chunk] = has ? 1 : 0;
        table[chunk] = has ? reinterpret_cast<const uint8_t *>(slot) : nullptr;
    }
}
// consecutive staged chunks -> their pool slots; copy of one slot to another (make_mut of a shared chunk)
__global__ void __launch_bounds__(256) scatter_chunks_kernel(const uint4 *src, uint8_t *const *dst, uint32_t n) {
    for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
        uint4 *d = reinterpret_cast<uint4 *>(dst[c]);
        const uint4 *s = src + (size_t)c * (kChunkBytes / 16);
        for (uint32_t i = threadIdx.x; i < kChunkBytes / 16; i += blockDim.x) d[i] = s[i];
    }
}
__global__ void __launch_bounds__(256) copy_chunks_kernel(const uint8_t *const *src, uint8_t *const *dst, uint32_t n) {
    for (uint32_t c = blockIdx.x; c < n; c += gridDim.x) {
        uint4 *d = reinterpret_cast<uint4 *>(dst[c]);
        const uint4 *s = reinterpret_cast<const uint4 *>(src[c]);  // null = a fresh chunk: zeros
        for (uint32_t i = threadIdx.x; i < kChunkBytes / 16; i += blockDim.x) d[i] = s ? s[i] : make_uint4(0, 0, 0, 0);
    }
}
// TiledImage::to_rgba_image (tiled_image.rs:271-293): absent chunks are transparent
__global__ void __launch_bounds__(256) tiles_to_flat_kernel(const uint8_t *const *table, uint32_t w, uint32_t h, uint32_t chunks_x,
                                                            uint32_t *flat) {
    const uint32_t chunk = blockIdx.x, cy = chunk / chunks_x, cx = chunk - cy * chunks_x;
    const uint32_t *slot = reinterpret_cast<const uint32_t *>(table[chunk]);
    for (uint32_t i = threadIdx.x; i < PFE_CHUNK_SIZE * PFE_CHUNK_SIZE; i += blockDim.x) {
        const uint32_t ly = i / PFE_CHUNK_SIZE, lx = i % PFE_CHUNK_SIZE;
        const uint32_t gx = cx * PFE_CHUNK_SIZE + lx, gy = cy * PFE_CHUNK_SIZE + ly;
        if (gx < w && gy < h) flat[(size_t)gy * w + gx] = slot ? slot[i] : 0u;
    }
}

template <class F>
void parallel_for(size_t n, F f) {
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 16u));
    if (n < 64 || nt == 1) { for (size_t i = 0; i < n; i++) f(i); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back([=]() { for (size_t i = t; i < n; i += nt) f(i); });
    for (auto &t : th) t.join();
}

constexpr size_t kStageBytes = 32u << 20;  // per pinned slice: 2048 chunks
int stage_init(pfe_ctx *ctx) {
    for (int i = 0; i < 2; i++) {
        if (!ctx->stage[i]) PFE_CUDA(ctx, cudaMallocHost(&ctx->stage[i], kStageBytes));
        if (!ctx->stage_ev[i]) PFE_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
    }
    return PFE_OK;
}
// Scattered host chunks -> consecutive device slots: gather a slice into pinned memory with all host
// threads while the previous slice is still on the wire.
int upload_chunks(pfe_ctx *ctx, const std::vector<const uint8_t *> &src, uint8_t *dev) {
    if (src.empty()) return PFE_OK;
    PFE_TRY(stage_init(ctx));
    const size_t per = kStageBytes / kChunkBytes;
    int buf = 0;
    for (size_t base = 0; base < src.size(); base += per, buf ^= 1) {
        const size_t cnt = std::min(per, src.size() - base);
        PFE_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[buf]));  // the copy that last used this slice has drained
        uint8_t *st = (uint8_t *)ctx->stage[buf];
        const uint8_t *const *s = src.data() + base;
        parallel_for(cnt, [=](size_t i) { memcpy(st + i * kChunkBytes, s[i], kChunkBytes); });
        PFE_CUDA(ctx, cudaMemcpyAsync(dev + base * kChunkBytes, st, cnt * kChunkBytes, cudaMemcpyHostToDevice, ctx->stream));
        PFE_CUDA(ctx, cudaEventRecord(ctx->stage_ev[buf], ctx->stream));
    }
    return PFE_OK;
}

int run_flatten_tiles(pfe_ctx *ctx, const pfe_tile_layer_desc *layers, uint32_t n, uint32_t w, uint32_t h, const uint8_t *active_dev,
                      uint8_t *dst) {
    TileParams P;
    memset(&P, 0, sizeof(P));
    P.pc = packed_consts();
    P.active = active_dev;
    P.dst = dst;
    P.w = w; P.h = h;
    P.chunks_x = pfe_div_up(w, PFE_CHUNK_SIZE);
    P.n_chunks = P.chunks_x * pfe_div_up(h, PFE_CHUNK_SIZE);
    P.vec_store = ((w & 3u) == 0 && ((uintptr_t)dst & 15) == 0) ? 1u : 0u;
    const size_t smem = kLutBytes;
    PFE_CUDA(ctx, cudaFuncSetAttribute(flatten_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = std::min<unsigned>(P.n_chunks, (unsigned)ctx->sm_count * 12);
    bool first = true;
    uint32_t adj_used = 0;
    auto flush = [&](bool force) -> int {
        if (P.n_layers == 0 && !(force && first)) return PFE_OK;
        P.init_from_dst = first ? 0u : 1u;
        PFE_KERNEL(ctx, "flatten_tiles", flatten_tiles_kernel<<<blocks, 256, smem, ctx->stream>>>(P));
        PFE_LAUNCHED(ctx);
        first = false;
        P.n_layers = 0;
        adj_used = 0;
        return PFE_OK;
    };
    for (uint32_t i = 0; i < n; i++) {
        const pfe_tile_layer_desc &S = layers[i];
        if (!S.visible) continue;
        if (S.kind > PFE_LAYER_ADJ_CHANNEL_MIXER) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "flatten_tiles: bad layer kind");
        if (S.kind == PFE_LAYER_RASTER && !S.chunks) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "flatten_tiles: raster layer without a chunk table");
        if (P.n_layers == kMaxLayers || (S.kind != PFE_LAYER_RASTER && adj_used == kMaxAdj)) PFE_TRY(flush(false));
        TileLayer &L = P.layers[P.n_layers++];
        L.chunks = S.chunks;
        L.mask_chunks = S.mask_chunks;
        L.opacity = S.opacity;
        L.blend = S.blend > 24 ? 0 : S.blend;
        L.kind = S.kind;
        L.adj_slot = 0;
        if (S.kind != PFE_LAYER_RASTER) {
            L.adj_slot = (uint8_t)adj_used;
            memcpy(P.adj[adj_used++], S.adj, sizeof(float) * 16);
        }
    }
    return flush(true);  // no visible layer: the result is transparent
}

int check_tiles_args(pfe_ctx *ctx, const pfe_tile_layer_desc *layers, uint32_t n, uint32_t w, uint32_t h, const void *dst) {
    if (!ctx) return PFE_ERR_INVALID_ARG;
    if ((!layers && n) || !dst || !w || !h || w > 0x7FFFFFFFu / 4 || h > 0x7FFFFFFFu / 4) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "flatten_tiles: bad args");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    return PFE_OK;
}

}  // namespace

extern "C" int pfe_dev_flatten_tiles(pfe_ctx *ctx, const pfe_tile_layer_desc *layers, uint32_t n, uint32_t w, uint32_t h,
                                     uint8_t *dst) {
    PFE_TRY(check_tiles_args(ctx, layers, n, w, h, dst));
    const uint32_t nch = pfe_div_up(w, PFE_CHUNK_SIZE) * pfe_div_up(h, PFE_CHUNK_SIZE);
    void *act;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, nch, &act));
    PFE_CUDA(ctx, cudaMemsetAsync(act, 0, nch, ctx->stream));
    for (uint32_t i = 0; i < n; i++)
        if (layers[i].visible && layers[i].kind == PFE_LAYER_RASTER && layers[i].chunks) {
            PFE_KERNEL(ctx, "tile_active", tile_active_kernel<<<pfe_div_up(nch, 256), 256, 0, ctx->stream>>>(layers[i].chunks, (uint8_t *)act, nch));
            PFE_LAUNCHED(ctx);
        }
    return run_flatten_tiles(ctx, layers, n, w, h, (const uint8_t *)act, dst);
}

// Host tier: chunk tables are host arrays of host pointers (what `Vec<Option<Arc<RgbaImage>>>` flattens
// to). Populated chunks of visible raster layers (and of their masks) are packed into one device pool.
extern "C" int pfe_flatten_tiles(pfe_ctx *ctx, const pfe_tile_layer_desc *layers, uint32_t n, uint32_t w, uint32_t h, uint8_t *dst) {
    PFE_TRY(check_tiles_args(ctx, layers, n, w, h, dst));
    const uint32_t nch = pfe_div_up(w, PFE_CHUNK_SIZE) * pfe_div_up(h, PFE_CHUNK_SIZE);
    std::vector<const uint8_t *> src;           // populated chunks in pool order
    std::vector<uint64_t> tables;                // per used table: nch entries, slot index + 1 or 0
    std::vector<uint8_t> active(nch, 0);
    std::vector<pfe_tile_layer_desc> dl(layers, layers + n);
    std::vector<std::pair<uint32_t, int>> fix;  // (layer, 0 = chunks / 1 = mask) -> table index in `tables`
    for (uint32_t i = 0; i < n; i++) {
        const pfe_tile_layer_desc &S = layers[i];
        if (!S.visible || S.kind != PFE_LAYER_RASTER) continue;
        if (!S.chunks) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "flatten_tiles: raster layer without a chunk table");
        for (int which = 0; which < 2; which++) {
            const uint8_t *const *t = which ? S.mask_chunks : S.chunks;
            if (!t) continue;
            fix.emplace_back(i, which);
            const size_t base = tables.size();
            tables.resize(base + nch, 0);
            for (uint32_t c = 0; c < nch; c++)
                if (t[c] && (which == 0 || S.chunks[c])) {  // a mask chunk under an absent layer chunk is never read
                    src.push_back(t[c]);
                    tables[base + c] = src.size();
                    if (which == 0) active[c] = 1;
                }
        }
    }
    // device layout (scratch A): [pool | tables | active | result]
    const size_t pool_b = src.size() * kChunkBytes, tab_b = tables.size() * 8, act_b = (nch + 255) & ~size_t(255);
    const size_t out_b = (size_t)w * h * 4;
    void *base;
    PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_A, pool_b + tab_b + act_b + out_b + 256, &base));
    uint8_t *pool = (uint8_t *)base, *tab_d = pool + pool_b, *act_d = tab_d + tab_b, *out_d = act_d + act_b;
    for (auto &e : tables) e = e ? (uint64_t)(uintptr_t)(pool + (e - 1) * kChunkBytes) : 0;
    PFE_TRY(upload_chunks(ctx, src, pool));
    if (tab_b) PFE_CUDA(ctx, cudaMemcpyAsync(tab_d, tables.data(), tab_b, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaMemcpyAsync(act_d, active.data(), nch, cudaMemcpyHostToDevice, ctx->stream));
    for (size_t k = 0; k < fix.size(); k++) {
        const uint8_t *const *t = reinterpret_cast<const uint8_t *const *>(tab_d + k * (size_t)nch * 8);
        if (fix[k].second) dl[fix[k].first].mask_chunks = t; else dl[fix[k].first].chunks = t;
    }
    PFE_TRY(run_flatten_tiles(ctx, dl.data(), n, w, h, act_d, out_d));
    PFE_CUDA(ctx, cudaMemcpyAsync(dst, out_d, out_b, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // also keeps `tables` / `active` alive until their copies ran
    return PFE_OK;
}

// ---- device-resident TiledImage -----------------------------------------------------------------
namespace {
uint8_t *slot_ptr(pfe_ctx *ctx, int32_t id) {
    return ctx->chunks.slabs[(uint32_t)id / pfe_ctx::ChunkPool::kPerSlab] + (size_t)((uint32_t)id % pfe_ctx::ChunkPool::kPerSlab) * kChunkBytes;
}
int slot_alloc(pfe_ctx *ctx, int32_t *out) {
    auto &cp = ctx->chunks;
    if (cp.free_list.empty()) {
        uint8_t *slab = nullptr;
        cudaError_t e = cudaMalloc(&slab, (size_t)pfe_ctx::ChunkPool::kPerSlab * kChunkBytes);
        if (e != cudaSuccess) { cudaGetLastError(); return pfe_fail(ctx, PFE_ERR_OOM, "tiled: chunk pool", e); }
        const uint32_t base = (uint32_t)cp.refs.size();
        cp.slabs.push_back(slab);
        cp.refs.resize(base + pfe_ctx::ChunkPool::kPerSlab, 0);
        for (uint32_t i = pfe_ctx::ChunkPool::kPerSlab; i-- > 0;) cp.free_list.push_back(base + i);
    }
    *out = (int32_t)cp.free_list.back();
    cp.free_list.pop_back();
    cp.refs[*out] = 1;
    return PFE_OK;
}
void slot_release(pfe_ctx *ctx, int32_t id) {
    if (id < 0) return;
    if (--ctx->chunks.refs[id] == 0) ctx->chunks.free_list.push_back((uint32_t)id);  // reuse is ordered by the stream
}
int check_tiled(pfe_ctx *ctx, const pfe_tiled *t) {
    if (!ctx || !t) return PFE_ERR_INVALID_ARG;
    if (t->owner != ctx) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "tiled: image belongs to another context");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    return PFE_OK;
}
// Give every chunk in `which` a slot nobody else holds.  keep = the chunk's pixels survive (shared chunks are copied,
// absent ones start as zeros); otherwise the caller overwrites the whole chunk.  Returns the chunks' slot addresses.
int make_private(pfe_ctx *ctx, pfe_tiled *t, const std::vector<uint32_t> &which, bool keep, std::vector<uint64_t> *addrs) {
    std::vector<uint64_t> copy_src, copy_dst;
    addrs->assign(which.size(), 0);
    for (size_t k = 0; k < which.size(); k++) {
        const uint32_t c = which[k];
        const int32_t old = t->slot[c];
        if (old >= 0 && ctx->chunks.refs[old] == 1) { (*addrs)[k] = (uint64_t)(uintptr_t)slot_ptr(ctx, old); continue; }
        int32_t fresh;
        PFE_TRY(slot_alloc(ctx, &fresh));
        if (keep) {
            copy_src.push_back(old >= 0 ? (uint64_t)(uintptr_t)slot_ptr(ctx, old) : 0);
            copy_dst.push_back((uint64_t)(uintptr_t)slot_ptr(ctx, fresh));
        }
        slot_release(ctx, old);
        t->slot[c] = fresh;
        (*addrs)[k] = (uint64_t)(uintptr_t)slot_ptr(ctx, fresh);
    }
    if (!copy_dst.empty()) {
        const size_t n = copy_dst.size();
        void *lists;
        PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_C, n * 16, &lists));
        PFE_CUDA(ctx, cudaMemcpyAsync(lists, copy_src.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
        PFE_CUDA(ctx, cudaMemcpyAsync((char *)lists + n * 8, copy_dst.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream));
        PFE_KERNEL(ctx, "tiled_copy_chunks", copy_chunks_kernel<<<(unsigned)std::min<size_t>(n, (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
            (const uint8_t *const *)lists, (uint8_t *const *)((char *)lists + n * 8), (uint32_t)n));
        PFE_LAUNCHED(ctx);
        PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // copy_src / copy_dst are host vectors
    }
    return PFE_OK;
}
}  // namespace

extern "C" int pfe_tiled_create(pfe_ctx *ctx, uint32_t w, uint32_t h, pfe_tiled **out) {
    if (!ctx || !out) return PFE_ERR_INVALID_ARG;
    *out = nullptr;
    if (!w || !h || w > 0x7FFFFFFFu / 4 || h > 0x7FFFFFFFu / 4) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "tiled_create: bad size");
    PFE_CUDA(ctx, cudaSetDevice(ctx->device));
    pfe_tiled *t = new (std::nothrow) pfe_tiled();
    if (!t) return PFE_ERR_OOM;
    t->owner = ctx;
    t->w = w; t->h = h;
    t->chunks_x = pfe_div_up(w, PFE_CHUNK_SIZE);
    t->chunks_y = pfe_div_up(h, PFE_CHUNK_SIZE);
    const size_t nch = (size_t)t->chunks_x * t->chunks_y;
    t->slot.assign(nch, -1);
    cudaError_t e = cudaMalloc(&t->table, nch * sizeof(void *));
    if (e == cudaSuccess) e = cudaMalloc(&t->occupancy, nch);
    if (e == cudaSuccess) e = cudaMemsetAsync(t->table, 0, nch * sizeof(void *), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(t->occupancy, 0, nch, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        pfe_tiled_destroy(ctx, t);
        return pfe_fail(ctx, PFE_ERR_OOM, "tiled_create", e);
    }
    *out = t;
    return PFE_OK;
}
extern "C" int pfe_tiled_destroy(pfe_ctx *ctx, pfe_tiled *t) {
    if (!ctx || !t) return PFE_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (t->owner == ctx)
        for (int32_t id : t->slot) slot_release(ctx, id);
    if (t->table) cudaFree((void *)t->table);
    if (t->occupancy) cudaFree(t->occupancy);
    delete t;
    return PFE_OK;
}
// A second image that shares every chunk of `src` (the snapshot an undo step keeps): no pixel is copied.
extern "C" int pfe_tiled_clone(pfe_ctx *ctx, const pfe_tiled *src, pfe_tiled **out) {
    if (!out) return PFE_ERR_INVALID_ARG;
    *out = nullptr;
    PFE_TRY(check_tiled(ctx, src));
    pfe_tiled *t = nullptr;
    PFE_TRY(pfe_tiled_create(ctx, src->w, src->h, &t));
    t->slot = src->slot;
    for (int32_t id : t->slot)
        if (id >= 0) ctx->chunks.refs[id]++;
    const size_t nch = t->slot.size();
    cudaError_t e = cudaMemcpyAsync(t->table, src->table, nch * sizeof(void *), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->occupancy, src->occupancy, nch, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) { pfe_tiled_destroy(ctx, t); return pfe_fail(ctx, PFE_ERR_CUDA, "tiled_clone", e); }
    *out = t;
    return PFE_OK;
}
// TiledImage::ensure_chunk_mut (tiled_image.rs:868) for a list of chunks: afterwards each is populated (a fresh chunk
// starts transparent) and held by this image alone, so device code may write it through pfe_tiled_table.
extern "C" int pfe_tiled_make_mut(pfe_ctx *ctx, pfe_tiled *t, const uint32_t *chunk_indices, uint32_t n) {
    PFE_TRY(check_tiled(ctx, t));
    if (n && !chunk_indices) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "tiled_make_mut: null list");
    std::vector<uint32_t> which(chunk_indices, chunk_indices + n);
    for (uint32_t c : which)
        if (c >= t->slot.size()) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "tiled_make_mut: chunk index out of range");
    // a slot that is assigned but unpopulated (from_flat found no alpha) holds stale pixels: treat it as absent
    std::vector<uint8_t> occ(t->slot.size());
    PFE_CUDA(ctx, cudaMemcpyAsync(occ.data(), t->occupancy, occ.size(), cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t c : which)
        if (!occ[c] && t->slot[c] >= 0) { slot_release(ctx, t->slot[c]); t->slot[c] = -1; }
    std::vector<uint64_t> addrs;
    PFE_TRY(make_private(ctx, t, which, true, &addrs));
    const uint8_t one = 1;
    for (size_t k = 0; k < which.size(); k++) {  // n is small (the chunks under one edit)
        PFE_CUDA(ctx, cudaMemcpyAsync((void *)(t->table + which[k]), &addrs[k], 8, cudaMemcpyHostToDevice, ctx->stream));
        PFE_CUDA(ctx, cudaMemcpyAsync(t->occupancy + which[k], &one, 1, cudaMemcpyHostToDevice, ctx->stream));
    }
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}
// Pool slot of every chunk (-1 = none): two images share a chunk exactly when their ids agree.  For tests and for a
// caller that wants to know what a snapshot costs.
extern "C" int pfe_tiled_chunk_ids(pfe_ctx *ctx, const pfe_tiled *t, int32_t *ids_out) {
    PFE_TRY(check_tiled(ctx, t));
    if (!ids_out) return pfe_fail(ctx, PFE_ERR_INVALID_ARG, "tiled_chunk_ids: null output");
    memcpy(ids_out, t->slot.data(), t->slot.size() * sizeof(int32_t));
    return PFE_OK;
}
extern "C" int pfe_tiled_upload(pfe_ctx *ctx, pfe_tiled *t, const uint8_t *const *host_chunk_table) {
    PFE_TRY(check_tiled(ctx, t));
    if (!host_chunk_table) return PFE_ERR_INVALID_ARG;
    const size_t nch = t->slot.size();
    // populated chunks get private slots (their old content is replaced), the others give theirs back
    std::vector<uint32_t> which;
    std::vector<const uint8_t *> src;
    for (size_t c = 0; c < nch; c++) {
        if (host_chunk_table[c]) { which.push_back((uint32_t)c); src.push_back(host_chunk_table[c]); }
        else { slot_release(ctx, t->slot[c]); t->slot[c] = -1; }
    }
    std::vector<uint64_t> addrs;
    PFE_TRY(make_private(ctx, t, which, false, &addrs));
    std::vector<uint64_t> table(nch, 0);
    std::vector<uint8_t> occ(nch, 0);
    for (size_t k = 0; k < which.size(); k++) { table[which[k]] = addrs[k]; occ[which[k]] = 1; }
    if (!src.empty()) {
        // staged uploads land consecutively in scratch; one kernel moves them to their slots
        void *stage_dev;
        PFE_TRY(pfe_scratch(ctx, PFE_SCRATCH_A, src.size() * (kChunkBytes + 8), &stage_dev));
        uint8_t *lists = (uint8_t *)stage_dev + src.size() * kChunkBytes;
        PFE_TRY(upload_chunks(ctx, src, (uint8_t *)stage_dev));
        PFE_CUDA(ctx, cudaMemcpyAsync(lists, addrs.data(), addrs.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        PFE_KERNEL(ctx, "tiled_scatter", scatter_chunks_kernel<<<(unsigned)std::min<size_t>(src.size(), (size_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
            (const uint4 *)stage_dev, (uint8_t *const *)lists, (uint32_t)src.size()));
        PFE_LAUNCHED(ctx);
    }
    PFE_CUDA(ctx, cudaMemcpyAsync(t->table, table.data(), nch * 8, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaMemcpyAsync(t->occupancy, occ.data(), nch, cudaMemcpyHostToDevice, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}
extern "C" int pfe_tiled_from_flat(pfe_ctx *ctx, pfe_tiled *t, const uint8_t *flat_dev) {
    PFE_TRY(check_tiled(ctx, t));
    if (!flat_dev) return PFE_ERR_INVALID_ARG;
    // which chunks end up populated is decided on the device (the alpha scan): every chunk gets a private slot to be
    // written into; the kernel nulls the table entries of the chunks that turn out transparent
    const size_t nch = t->slot.size();
    std::vector<uint32_t> all(nch);
    for (size_t c = 0; c < nch; c++) all[c] = (uint32_t)c;
    std::vector<uint64_t> addrs;
    PFE_TRY(make_private(ctx, t, all, false, &addrs));
    if (nch * 8 <= PFE_SMALL_BYTES / 2) {  // through the context's pinned ring: the call stays asynchronous
        void *stage;
        PFE_TRY(pfe_small_upload(ctx, addrs.data(), nch * 8, &stage));
        PFE_CUDA(ctx, cudaMemcpyAsync(t->table, stage, nch * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        PFE_CUDA(ctx, cudaMemcpyAsync(t->table, addrs.data(), nch * 8, cudaMemcpyHostToDevice, ctx->stream));
        PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // `addrs` is a host vector
    }
    PFE_KERNEL(ctx, "tiles_from_flat", tiles_from_flat_kernel<<<t->chunks_x * t->chunks_y, 256, 0, ctx->stream>>>(
        (const uint32_t *)flat_dev, t->w, t->h, t->chunks_x, t->table, t->occupancy));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}
extern "C" int pfe_tiled_to_flat(pfe_ctx *ctx, const pfe_tiled *t, uint8_t *flat_dev) {
    PFE_TRY(check_tiled(ctx, t));
    if (!flat_dev) return PFE_ERR_INVALID_ARG;
    PFE_KERNEL(ctx, "tiles_to_flat", tiles_to_flat_kernel<<<t->chunks_x * t->chunks_y, 256, 0, ctx->stream>>>(
        t->table, t->w, t->h, t->chunks_x, (uint32_t *)flat_dev));
    PFE_LAUNCHED(ctx);
    return PFE_OK;
}
extern "C" int pfe_tiled_download(pfe_ctx *ctx, const pfe_tiled *t, uint8_t *occupancy, uint8_t *tiles) {
    PFE_TRY(check_tiled(ctx, t));
    if (!occupancy) return PFE_ERR_INVALID_ARG;
    const size_t nch = t->slot.size();
    PFE_CUDA(ctx, cudaMemcpyAsync(occupancy, t->occupancy, nch, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (tiles)
        for (size_t c = 0; c < nch; c++)
            if (occupancy[c] && t->slot[c] >= 0)
                PFE_CUDA(ctx, cudaMemcpyAsync(tiles + c * kChunkBytes, slot_ptr(ctx, t->slot[c]), kChunkBytes, cudaMemcpyDeviceToHost, ctx->stream));
    PFE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PFE_OK;
}
extern "C" const uint8_t *const *pfe_tiled_table(const pfe_tiled *t) { return t ? t->table : nullptr; }
extern "C" const uint8_t *pfe_tiled_occupancy(const pfe_tiled *t) { return t ? t->occupancy : nullptr; }
