/*
 * pfe_b200.h — C ABI of libpfe_b200.so, the B200 (sm_100a) pixel engine that replaces
 * PaintFE's layer compositor and filter / adjustment / warp / brush-stamp kernels.
 *
 * Every entry point names the reference interface it stands in for (file:line under the
 * PaintFE tree, v1.3.9 @ 16410bd).  Shapes follow the reference's own conventions:
 *   - images are dense, row-major, straight-alpha RGBA8, exactly w*h*4 bytes
 *     (what `RgbaImage::as_raw()` / `RgbaImage::from_raw()` hold);
 *   - selection masks are `GrayImage` planes of the image's size, w*h bytes, 0 = unselected;
 *   - displacement fields are w*h*2 f32, (dx,dy) interleaved (`DisplacementField::data`);
 *   - the caller owns every buffer; nothing is allocated across the ABI; outputs are written
 *     into caller-provided buffers and may not alias inputs unless stated.
 *
 * Two tiers, same semantics:
 *   pfe_<op>      host pointers; copies in, runs the CUDA kernels, copies out, returns when the
 *                 result is in `dst` (what a `*_core(&RgbaImage, …) -> RgbaImage` call needs).
 *   pfe_dev_<op>  device pointers on the context's device; asynchronous on the context's
 *                 stream; for chaining ops without PCIe round trips.
 *
 * There is no CPU fallback.  Every function returns PFE_OK or a negative status; on error the
 * output buffer content is unspecified and pfe_last_error() describes the failure.
 * A pfe_ctx may be used from one thread at a time; distinct contexts are independent.
 */
#ifndef PFE_B200_H
#define PFE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFE_ABI_VERSION 1
#define PFE_CHUNK_SIZE 64u /* src/canvas/defs.rs:7 */

typedef enum pfe_status {
    PFE_OK = 0,
    PFE_ERR_INVALID_ARG = -1, /* null pointer, zero size, bad enum */
    PFE_ERR_UNSUPPORTED = -2, /* the `Option::None` of GpuRenderer::median_rgba etc. */
    PFE_ERR_CUDA = -3,        /* a CUDA runtime call failed; see pfe_last_error */
    PFE_ERR_NO_DEVICE = -4,   /* no CUDA device: the library never computes on the CPU */
    PFE_ERR_OOM = -5
} pfe_status;

typedef struct pfe_ctx pfe_ctx;

/* -- context ---------------------------------------------------------------------------
 * Replaces GpuRenderer::try_new / GpuContext (src/gpu/renderer.rs:245, src/gpu/context.rs:9).
 * One context = one device + one stream + scratch buffers. */
int pfe_ctx_create(int device, pfe_ctx **out);
int pfe_ctx_destroy(pfe_ctx *ctx);
/* Run on a caller-owned cudaStream_t (e.g. a framework's current stream). NULL is the legacy
 * default stream, as in CUDA itself; pfe_ctx_use_own_stream switches back to the context's own. */
int pfe_ctx_set_stream(pfe_ctx *ctx, void *cuda_stream);
int pfe_ctx_use_own_stream(pfe_ctx *ctx);
int pfe_ctx_sync(pfe_ctx *ctx);
/* Stream-asynchronous device-tier calls that can only detect a caller error on the device (today:
 * pfe_dev_warp_band, whose source-row window may turn out not to cover the warp's reach) record it in a
 * sticky flag instead of synchronising (so does pfe_dev_peer_wait for a timeout).  This call synchronises the context's stream, returns
 * PFE_ERR_INVALID_ARG (message in pfe_last_error) if any such error was recorded since the last call, PFE_OK
 * otherwise, and clears the flag. */
int pfe_ctx_check_async(pfe_ctx *ctx);
const char *pfe_last_error(const pfe_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t pfe_ctx_launch_count(const pfe_ctx *ctx);
int pfe_abi_version(void);
/* Per-kernel device timing: when enabled every kernel launch is bracketed by CUDA events on the
 * launching stream. pfe_ctx_profile_read synchronises, writes a JSON object
 * {"<kernel>": {"launches": n, "ms": total}, ...} into buf and clears the record. */
int pfe_ctx_profile(pfe_ctx *ctx, int enable);
int pfe_ctx_profile_read(pfe_ctx *ctx, char *buf, size_t cap);
/* Pinned host memory for callers that want full-rate PCIe copies. */
int pfe_host_alloc(size_t bytes, void **out);
int pfe_host_free(void *p);
/* Plain device memory helpers so non-CUDA hosts (Rust, Python) can use the pfe_dev_ tier. */
int pfe_dev_alloc(pfe_ctx *ctx, size_t bytes, void **out);
int pfe_dev_free(pfe_ctx *ctx, void *p);
int pfe_dev_upload(pfe_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes);
int pfe_dev_download(pfe_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes);

/* -- flatten ---------------------------------------------------------------------------
 * Replaces CanvasState::composite (src/canvas/canvas_state.rs:482-698), blend_pixel_static
 * (:1246-1505) and GpuRenderer::composite (src/gpu/renderer.rs:533-583). Bit-exact with the
 * CPU compositor (not with the premultiplied wgpu shader). */
typedef enum pfe_layer_kind {
    PFE_LAYER_RASTER = 0,
    PFE_LAYER_ADJ_EXPOSURE = 1,            /* adj[0] = gain = 2^ev (host powf), layers.rs:279 */
    PFE_LAYER_ADJ_BRIGHTNESS_CONTRAST = 2, /* adj[0] = brightness, adj[1] = contrast, :288 */
    PFE_LAYER_ADJ_INVERT = 3,              /* :298 */
    PFE_LAYER_ADJ_CHANNEL_MIXER = 4        /* adj[0..16] = red[4] green[4] blue[4] alpha[4], :299 */
} pfe_layer_kind;

typedef struct pfe_layer_desc {
    const uint8_t *rgba; /* w*h*4; NULL for adjustment layers. Layer::pixels flattened */
    const uint8_t *mask; /* w*h conceal plane = alpha of Layer::mask, or NULL (layers.rs:395-399) */
    float opacity;       /* Layer::opacity */
    uint8_t blend;       /* BlendMode::to_u8, layers.rs:125-153 (unknown ids = Normal, :183) */
    uint8_t visible;     /* layer_effectively_visible */
    uint8_t kind;        /* pfe_layer_kind */
    uint8_t _pad;
    float adj[16];
} pfe_layer_desc;

/* `active_chunks`: ceil(w/64)*ceil(h/64) bytes, row-major, non-zero where any visible layer
 * has a populated tile (canvas_state.rs:529-550); pixels of inactive chunks stay (0,0,0,0).
 * NULL = all chunks active (dense canvas). Only matters when adjustment layers are present. */
int pfe_flatten(pfe_ctx *ctx, const pfe_layer_desc *layers, uint32_t n_layers, uint32_t w,
                uint32_t h, const uint8_t *active_chunks, uint8_t *dst);
int pfe_dev_flatten(pfe_ctx *ctx, const pfe_layer_desc *layers_with_dev_ptrs, uint32_t n_layers,
                    uint32_t w, uint32_t h, const uint8_t *active_chunks_dev, uint8_t *dst_dev);

/* -- tile-native flatten and device-resident TiledImage --------------------------
 * The compositor's own data structure: a layer is a row-major table of ceil(w/64)*ceil(h/64) chunk
 * pointers, each to 64*64*4 bytes (row-major RGBA8, zero padded past the canvas edge) or NULL for an
 * unpopulated chunk - `Vec<Option<Arc<RgbaImage>>>`, src/canvas/tiled_image.rs:2-7, with
 * `chunk.as_raw().as_ptr()` per entry.  pfe_flatten_tiles is CanvasState::composite
 * (src/canvas/canvas_state.rs:482-698) on those tables: chunks no visible raster layer populates stay
 * transparent and are never read (:529-550), a layer without a chunk at a position is skipped there
 * (:596-598), `mask_chunks` is Layer::mask (its alpha conceals, :660-665; NULL = none / mask disabled).
 * Host tier: tables are host arrays of host pointers; only populated chunks are copied to the device.
 * Device tier: tables are device arrays of device pointers (pfe_tiled_table). dst is the flat w*h image. */
typedef struct pfe_tile_layer_desc {
    const uint8_t *const *chunks;      /* NULL for adjustment layers */
    const uint8_t *const *mask_chunks; /* or NULL */
    float opacity;
    uint8_t blend;
    uint8_t visible;
    uint8_t kind; /* pfe_layer_kind */
    uint8_t _pad;
    float adj[16];
} pfe_tile_layer_desc;
int pfe_flatten_tiles(pfe_ctx *ctx, const pfe_tile_layer_desc *layers, uint32_t n_layers, uint32_t w,
                      uint32_t h, uint8_t *dst);
int pfe_dev_flatten_tiles(pfe_ctx *ctx, const pfe_tile_layer_desc *layers_with_dev_tables,
                          uint32_t n_layers, uint32_t w, uint32_t h, uint8_t *dst_dev);
/* A TiledImage resident on the device: pointer table + occupancy bytes over chunks that live in the context's
 * reference-counted chunk pool, so that images cloned from one another share unchanged chunks the way
 * `Vec<Option<Arc<RgbaImage>>>` does (copy on write: Arc::make_mut, tiled_image.rs:330, ensure_chunk_mut :868).
 * An image belongs to the context that created it.
 * clone      a second image sharing every chunk of the first (an undo snapshot): no pixel is copied.
 * make_mut   ensure_chunk_mut for a list of chunk indices: each becomes populated (fresh chunks are transparent) and
 *            private to this image (a shared chunk is copied first), so device code may write it through the table.
 * chunk_ids  pool slot per chunk, -1 = none: two images share a chunk exactly when the ids agree.
 * upload     TiledImage -> device; only populated chunks cross PCIe.
 * from_flat  TiledImage::from_rgba_image (tiled_image.rs:50-104) on the device: a chunk is populated iff
 *            some pixel in it has alpha != 0.
 * to_flat    TiledImage::to_rgba_image (:271-293).
 * download   occupancy (one byte per chunk) and, if `tiles` != NULL, the populated chunks at
 *            tiles + index * 16384 (the layout pfe_flat_to_tiles produces). */
typedef struct pfe_tiled pfe_tiled;
int pfe_tiled_create(pfe_ctx *ctx, uint32_t w, uint32_t h, pfe_tiled **out);
int pfe_tiled_destroy(pfe_ctx *ctx, pfe_tiled *t);
int pfe_tiled_clone(pfe_ctx *ctx, const pfe_tiled *src, pfe_tiled **out);
int pfe_tiled_make_mut(pfe_ctx *ctx, pfe_tiled *t, const uint32_t *chunk_indices, uint32_t n);
int pfe_tiled_chunk_ids(pfe_ctx *ctx, const pfe_tiled *t, int32_t *ids_out);
int pfe_tiled_upload(pfe_ctx *ctx, pfe_tiled *t, const uint8_t *const *host_chunk_table);
int pfe_tiled_from_flat(pfe_ctx *ctx, pfe_tiled *t, const uint8_t *flat_dev);
int pfe_tiled_to_flat(pfe_ctx *ctx, const pfe_tiled *t, uint8_t *flat_dev);
int pfe_tiled_download(pfe_ctx *ctx, const pfe_tiled *t, uint8_t *occupancy, uint8_t *tiles);
const uint8_t *const *pfe_tiled_table(const pfe_tiled *t);
const uint8_t *pfe_tiled_occupancy(const pfe_tiled *t);

/* -- blurs -----------------------------------------------------------------------------
 * flags for the Gaussian family: */
#define PFE_GAUSS_EXACT 1u /* reference tap order with separate mul/add: bit-exact with the CPU
                              path. Default (0) uses FMA accumulation: within +-1 level. */

/* blur_with_selection_pub (src/ops/filters.rs:130-207) / parallel_gaussian_blur (:242-316) /
 * GpuRenderer::blur_rgba (src/gpu/renderer.rs:915). mask NULL = whole image. */
int pfe_gaussian_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float sigma,
                      const uint8_t *mask, uint8_t *dst, uint32_t flags);
int pfe_dev_gaussian_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float sigma,
                          const uint8_t *mask, uint8_t *dst, uint32_t flags);
/* box_blur_core (src/ops/effects/blur.rs:233-318): integer, bit-exact. */
int pfe_box_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius,
                 const uint8_t *mask, uint8_t *dst);
int pfe_dev_box_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius,
                     const uint8_t *mask, uint8_t *dst);
/* motion_blur_core (blur.rs:144-210). */
int pfe_motion_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float angle_deg,
                    float distance, const uint8_t *mask, uint8_t *dst);
int pfe_dev_motion_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float angle_deg,
                        float distance, const uint8_t *mask, uint8_t *dst);
/* median_core (src/ops/effects/noise.rs:357-410) / GpuRenderer::median_rgba (renderer.rs:945).
 * Any radius from 1 to 20000 (radius 0 is treated as 1, noise.rs:364; PFE_ERR_UNSUPPORTED beyond); integer,
 * bit-exact. */
int pfe_median(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius,
               const uint8_t *mask, uint8_t *dst);
int pfe_dev_median(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius,
                   const uint8_t *mask, uint8_t *dst);
/* sharpen_core / unsharp mask (src/ops/effects/stylize.rs:96-141). */
int pfe_sharpen(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount,
                float radius, const uint8_t *mask, uint8_t *dst, uint32_t flags);
int pfe_dev_sharpen(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount,
                    float radius, const uint8_t *mask, uint8_t *dst, uint32_t flags);
/* vignette_core (stylize.rs:170-191). */
int pfe_vignette(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount,
                 float softness, const uint8_t *mask, uint8_t *dst);
int pfe_dev_vignette(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount,
                     float softness, const uint8_t *mask, uint8_t *dst);

/* -- further Effect-API kernels ------------------------------------------
 * glow_core (src/ops/effects/stylize.rs:26-76): blur(sigma=radius) then screen with the source;
 * fused into the Gaussian V pass. */
int pfe_glow(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius, float intensity,
             const uint8_t *mask, uint8_t *dst, uint32_t flags);
int pfe_dev_glow(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius, float intensity,
                 const uint8_t *mask, uint8_t *dst, uint32_t flags);
/* pixelate_core (src/ops/effects/distort.rs:333-373): block centre sample; block_size < 2 -> 2. */
int pfe_pixelate(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t block_size,
                 const uint8_t *mask, uint8_t *dst);
int pfe_dev_pixelate(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t block_size,
                     const uint8_t *mask, uint8_t *dst);
/* bulge_core_at / twist_core_at (distort.rs:400-437, :464-493); origin in 0..1 ((0.5,0.5) for the
 * plain _core forms). bulge is bit-exact; twist needs sin/cos per pixel: within +-1 level. */
int pfe_bulge(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, float origin_x,
              float origin_y, const uint8_t *mask, uint8_t *dst);
int pfe_dev_bulge(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, float origin_x,
                  float origin_y, const uint8_t *mask, uint8_t *dst);
int pfe_twist(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float angle_deg, float origin_x,
              float origin_y, const uint8_t *mask, uint8_t *dst);
int pfe_dev_twist(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float angle_deg, float origin_x,
                  float origin_y, const uint8_t *mask, uint8_t *dst);
/* add_noise_core (src/ops/effects/noise.rs:73-143). noise_type: 0 Uniform, 1 Gaussian, 2 Perlin.
 * Uniform and Perlin are bit-exact (integer hash + strict f32); monochrome Gaussian uses ln/cos
 * per pixel: within +-1 level. */
int pfe_add_noise(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, int noise_type,
                  int monochrome, uint32_t seed, float scale, uint32_t octaves, const uint8_t *mask, uint8_t *dst);
int pfe_dev_add_noise(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float amount, int noise_type,
                      int monochrome, uint32_t seed, float scale, uint32_t octaves, const uint8_t *mask,
                      uint8_t *dst);
/* reduce_noise_core (bilateral, noise.rs:172-262); exp per tap: within +-1 level. */
int pfe_reduce_noise(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float strength, uint32_t radius,
                     const uint8_t *mask, uint8_t *dst);
int pfe_dev_reduce_noise(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float strength,
                         uint32_t radius, const uint8_t *mask, uint8_t *dst);

/* -- the rest of src/ops/effects/ (each pinned by a golden in tests/visual_filters.rs) --------------- */
/* ink_core (src/ops/effects/artistic.rs:31-99): Sobel on Rec.709 luminance, thresholded to black/white. */
int pfe_ink(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float edge_strength, float threshold,
            const uint8_t *mask, uint8_t *dst);
int pfe_dev_ink(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float edge_strength,
                float threshold, const uint8_t *mask, uint8_t *dst);
/* oil_painting_core (artistic.rs:123-217): most common intensity bin in a (2r+1)^2 window; radius clamps to
 * 1..10, levels to 2..64; integer arithmetic. */
int pfe_oil_painting(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius,
                     uint32_t levels, const uint8_t *mask, uint8_t *dst);
int pfe_dev_oil_painting(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t radius,
                         uint32_t levels, const uint8_t *mask, uint8_t *dst);
/* color_filter_core (artistic.rs:266-307). color = RGBA; mode: 0 Multiply, 1 Screen, 2 Overlay, 3 SoftLight. */
int pfe_color_filter(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const uint8_t *color,
                     float intensity, int mode, const uint8_t *mask, uint8_t *dst);
int pfe_dev_color_filter(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const uint8_t *color,
                         float intensity, int mode, const uint8_t *mask, uint8_t *dst);
/* contours_core (src/ops/effects/contours.rs:56-112): iso-lines of a turbulence field blended over the image. */
int pfe_contours(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float scale, float frequency,
                 float line_width, const uint8_t *color, uint32_t seed, uint32_t octaves, float blend,
                 const uint8_t *mask, uint8_t *dst);
int pfe_dev_contours(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float scale,
                     float frequency, float line_width, const uint8_t *color, uint32_t seed,
                     uint32_t octaves, float blend, const uint8_t *mask, uint8_t *dst);
/* crystallize_core (src/ops/effects/distort.rs:26-169): jittered-grid Voronoi cells filled with their mean colour. */
int pfe_crystallize(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float cell_size,
                    uint32_t seed, const uint8_t *mask, uint8_t *dst);
int pfe_dev_crystallize(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float cell_size,
                        uint32_t seed, const uint8_t *mask, uint8_t *dst);
/* dents_core (distort.rs:248-310): turbulence displacement, bilinear sample. */
int pfe_dents(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float scale, float amount,
              uint32_t seed, uint32_t octaves, float roughness, int pinch, int wrap, const uint8_t *mask,
              uint8_t *dst);
int pfe_dev_dents(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float scale, float amount,
                  uint32_t seed, uint32_t octaves, float roughness, int pinch, int wrap,
                  const uint8_t *mask, uint8_t *dst);
/* halftone_core (src/ops/effects/stylize.rs:242-277). shape: 0 Circle, 1 Square, 2 Diamond, 3 Line. */
int pfe_halftone(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float dot_size, float angle_deg,
                 int shape, const uint8_t *mask, uint8_t *dst);
int pfe_dev_halftone(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float dot_size,
                     float angle_deg, int shape, const uint8_t *mask, uint8_t *dst);
/* bokeh_blur_core (src/ops/effects/blur.rs:22-115): equal-weight disc, integer sums; radius < 0.5 copies. */
int pfe_bokeh_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius,
                   const uint8_t *mask, uint8_t *dst);
int pfe_dev_bokeh_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float radius,
                       const uint8_t *mask, uint8_t *dst);
/* zoom_blur_core (blur.rs:322-427): nearest-neighbour samples toward (center_x, center_y) in 0..1; tint_rgba =
 * 4 floats in 0..1 or NULL. */
int pfe_zoom_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float center_x, float center_y,
                  float strength, uint32_t samples, const float *tint_rgba, float tint_strength,
                  const uint8_t *mask, uint8_t *dst);
int pfe_dev_zoom_blur(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, float center_x,
                      float center_y, float strength, uint32_t samples, const float *tint_rgba,
                      float tint_strength, const uint8_t *mask, uint8_t *dst);
/* grid_core (src/ops/effects/render.rs:52-92). style: 0 Lines, 1 Checkerboard. */
int pfe_grid(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t cell_w, uint32_t cell_h,
             uint32_t line_width, const uint8_t *color, int style, float opacity, const uint8_t *mask,
             uint8_t *dst);
int pfe_dev_grid(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t cell_w,
                 uint32_t cell_h, uint32_t line_width, const uint8_t *color, int style, float opacity,
                 const uint8_t *mask, uint8_t *dst);
/* canvas_border_core (render.rs:114-165). */
int pfe_canvas_border(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t width,
                      const uint8_t *color, const uint8_t *mask, uint8_t *dst);
int pfe_dev_canvas_border(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t width,
                          const uint8_t *color, const uint8_t *mask, uint8_t *dst);
/* shadow_core (render.rs:220-352): offset alpha, optional max-spread, Gaussian blur, shadow under the source.
 * flags as for pfe_gaussian_blur (PFE_GAUSS_EXACT = bit-exact blur). */
int pfe_drop_shadow(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int32_t offset_x,
                    int32_t offset_y, float blur_radius, int widen_radius, const uint8_t *color,
                    float opacity, const uint8_t *mask, uint8_t *dst, uint32_t flags);
int pfe_dev_drop_shadow(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int32_t offset_x,
                        int32_t offset_y, float blur_radius, int widen_radius, const uint8_t *color,
                        float opacity, const uint8_t *mask, uint8_t *dst, uint32_t flags);
/* outline_core (render.rs:403-572). mode: 0 Outside, 1 Inside, 2 Center. */
int pfe_outline(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t width,
                const uint8_t *color, int mode, int anti_alias, const uint8_t *mask, uint8_t *dst);
int pfe_dev_outline(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t width,
                    const uint8_t *color, int mode, int anti_alias, const uint8_t *mask, uint8_t *dst);
/* pixel_drag_core (src/ops/effects/glitch.rs:44-99). */
int pfe_pixel_drag(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t seed, float amount,
                   uint32_t distance, float direction, const uint8_t *mask, uint8_t *dst);
int pfe_dev_pixel_drag(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t seed,
                       float amount, uint32_t distance, float direction, const uint8_t *mask, uint8_t *dst);
/* rgb_displace_core (glitch.rs:142-197). offsets = {r_dx, r_dy, g_dx, g_dy, b_dx, b_dy}. */
int pfe_rgb_displace(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const int32_t *offsets,
                     const uint8_t *mask, uint8_t *dst);
int pfe_dev_rgb_displace(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const int32_t *offsets,
                         const uint8_t *mask, uint8_t *dst);

/* -- per-pixel adjustments ---------------------------------------------------------------
 * Ops 0..31 follow src/ops/adjustments.rs (round-to-nearest, selection mask honoured,
 * apply_pixel_transform[_from_flat] :21-108). Ops 32.. follow the inline Rhai bindings in
 * src/ops/scripting.rs:869-1075 (truncating casts, alpha untouched, no mask). Both exist in
 * the reference and are pinned by different goldens. */
typedef enum pfe_adjust_op {
    PFE_ADJ_INVERT = 0,              /* invert_colors, adjustments.rs:115; invert_rgba */
    PFE_ADJ_INVERT_ALPHA = 1,        /* :122 */
    PFE_ADJ_SEPIA = 2,               /* :133 */
    PFE_ADJ_DESATURATE = 3,          /* desaturate_layer, filters.rs:320 */
    PFE_ADJ_BRIGHTNESS_CONTRAST = 4, /* :265; params = {brightness, contrast} */
    PFE_ADJ_HSL = 5,                 /* :300; params = {hue_deg, saturation, lightness} */
    PFE_ADJ_EXPOSURE = 6,            /* :352; params = {gain = 2^ev} */
    PFE_ADJ_LUT_RGB = 7,             /* levels :398-446; luts = 256 bytes applied to r,g,b */
    PFE_ADJ_LUT_RGBA = 8,            /* curves :549-626, per-channel levels :490; luts = 4*256 */
    PFE_ADJ_TEMPERATURE_TINT = 9,    /* :518; params = {temperature, tint} */
    PFE_ADJ_HIGHLIGHTS_SHADOWS = 10, /* :371; params = {shadows, highlights} */
    PFE_ADJ_THRESHOLD = 11,          /* :1240; params = {level} */
    PFE_ADJ_POSTERIZE = 12,          /* :1267; params = {max(levels, 2)} */
    PFE_ADJ_COLOR_BALANCE = 13,      /* :1294; params = {shadows r,g,b, midtones r,g,b, highlights r,g,b} */
    PFE_ADJ_GRADIENT_MAP = 14,       /* :1344; luts = 256 RGBA entries (1024 bytes) indexed by luminance */
    PFE_ADJ_BLACK_AND_WHITE = 15,    /* :1373; params = {r_weight, g_weight, b_weight} */
    PFE_ADJ_VIBRANCE = 16,           /* :1408; params = {amount / 100} */
    PFE_ADJ_S_INVERT = 32,              /* apply_invert, scripting.rs:869 */
    PFE_ADJ_S_DESATURATE = 33,          /* apply_desaturate :883 */
    PFE_ADJ_S_SEPIA = 34,               /* apply_sepia() :900 */
    PFE_ADJ_S_SEPIA_STRENGTH = 35,      /* apply_sepia(strength) :921; params = {strength in [0,1]} */
    PFE_ADJ_S_BRIGHTNESS_CONTRAST = 36, /* :944 */
    PFE_ADJ_S_HSL = 37,                 /* :964 */
    PFE_ADJ_S_EXPOSURE = 38,            /* :1040; params = {gain} */
    PFE_ADJ_S_LUT_RGB = 39              /* apply_levels :1054 */
} pfe_adjust_op;

typedef struct pfe_adjust_desc {
    int32_t op;          /* pfe_adjust_op */
    float params[12];
    const uint8_t *luts; /* HOST pointer in both tiers (<= 1 KiB, copied with the launch) */
} pfe_adjust_desc;

/* `occupancy` (chunk bitmap like active_chunks) selects apply_pixel_transform's in-place tile
 * walk: pixels of unpopulated chunks are copied through untouched (tiled_image.rs:905-933).
 * NULL = the _from_flat form: every pixel is transformed. src may equal dst. */
int pfe_adjust(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, const pfe_adjust_desc *d,
               const uint8_t *mask, const uint8_t *occupancy, uint8_t *dst);
int pfe_dev_adjust(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h,
                   const pfe_adjust_desc *d, const uint8_t *mask, const uint8_t *occupancy_dev,
                   uint8_t *dst);
/* LUT builders (host arithmetic, identical to the reference's): build_levels_lut
 * adjustments.rs:424, scripting apply_levels LUT scripting.rs:1054, build_stretch_lut :232,
 * build_curves_lut :634 (n_points (x,y) pairs), build_multi_channel_luts :576. */
void pfe_build_levels_lut(float in_black, float in_white, float gamma, float out_black,
                          float out_white, uint8_t lut[256]);
void pfe_build_levels_lut_script(float in_black, float in_white, float gamma, uint8_t lut[256]);
void pfe_build_stretch_lut(uint8_t min, uint8_t max, uint8_t lut[256]);
void pfe_build_curves_lut(const float *points_xy, int n_points, uint8_t lut[256]);
void pfe_compose_curve_luts(const uint8_t in_rgb_r_g_b_a[5 * 256], uint8_t out_r_g_b_a[4 * 256]);
/* Per-channel min/max over pixels with alpha != 0 that are selected (auto_levels,
 * adjustments.rs:144-196). out = {min_r,max_r,min_g,max_g,min_b,max_b}. */
int pfe_channel_minmax(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h,
                       const uint8_t *mask, uint8_t out[6]);
int pfe_dev_channel_minmax(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h,
                           const uint8_t *mask, uint8_t out_host[6]);

/* -- geometry ----------------------------------------------------------------
 * Flips and quarter turns: imageops::{flip_horizontal, flip_vertical, rotate90, rotate270, rotate180} as
 * called from src/ops/transform.rs:326-334 and the Rhai bindings src/ops/scripting.rs:645-740; the same
 * pixels as TiledImage::{flip,rotate}_*_chunked behind flip_canvas_* / rotate_canvas_* (transform.rs:62-131).
 * dst is w*h, or h wide and w tall for the quarter turns. No selection mask (the reference has none here). */
typedef enum pfe_orient_op {
    PFE_ORIENT_FLIP_H = 0,
    PFE_ORIENT_FLIP_V = 1,
    PFE_ORIENT_ROTATE_90CW = 2,
    PFE_ORIENT_ROTATE_90CCW = 3,
    PFE_ORIENT_ROTATE_180 = 4
} pfe_orient_op;
int pfe_orient(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int op, uint8_t *dst);
int pfe_dev_orient(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, int op, uint8_t *dst);
/* resize_canvas / resize_canvas_layers (transform.rs:382-463): anchor_x/y in {0 start, 1 centre, 2 end}. */
int pfe_resize_canvas(pfe_ctx *ctx, const uint8_t *src, uint32_t old_w, uint32_t old_h, uint32_t new_w,
                      uint32_t new_h, uint32_t anchor_x, uint32_t anchor_y, const uint8_t fill_rgba[4],
                      uint8_t *dst);
int pfe_dev_resize_canvas(pfe_ctx *ctx, const uint8_t *src, uint32_t old_w, uint32_t old_h, uint32_t new_w,
                          uint32_t new_h, uint32_t anchor_x, uint32_t anchor_y, const uint8_t fill_rgba[4],
                          uint8_t *dst);
/* apply_affine (transform.rs:826-946) behind affine_transform_layer[_from_flat] (:750-820) and
 * rotate_canvas_arbitrary (:134-186): rotations in degrees, perspective focal = 1.5 * max(canvas);
 * nearest != 0 is Interpolation::Nearest, anything else the bilinear branch. dst = canvas_w * canvas_h. */
int pfe_affine(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h, uint32_t canvas_w,
               uint32_t canvas_h, float rotation_z, float rotation_x, float rotation_y, float scale,
               float offset_x, float offset_y, int nearest, uint8_t *dst);
int pfe_dev_affine(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h, uint32_t canvas_w,
                   uint32_t canvas_h, float rotation_z, float rotation_x, float rotation_y, float scale,
                   float offset_x, float offset_y, int nearest, uint8_t *dst);
/* imageops::resize as resize_image / resize_layers call it (transform.rs:347-378; Interpolation::to_filter
 * :47-55). `image` crate 0.25.9 arithmetic (not vendored in the reference; restated from its published
 * source and pinned by golden/transforms/resize_*.png). */
typedef enum pfe_resize_filter {
    PFE_RESIZE_NEAREST = 0,     /* Interpolation::Nearest  -> FilterType::Nearest */
    PFE_RESIZE_TRIANGLE = 1,    /* Interpolation::Bilinear -> FilterType::Triangle */
    PFE_RESIZE_CATMULL_ROM = 2, /* Interpolation::Bicubic  -> FilterType::CatmullRom */
    PFE_RESIZE_LANCZOS3 = 3     /* Interpolation::Lanczos3 -> FilterType::Lanczos3 */
} pfe_resize_filter;
int pfe_resize(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t new_w, uint32_t new_h,
               int filter, uint8_t *dst);
int pfe_dev_resize(pfe_ctx *ctx, const uint8_t *src, uint32_t w, uint32_t h, uint32_t new_w, uint32_t new_h,
                   int filter, uint8_t *dst);

/* -- warps -----------------------------------------------------------------------------
 * warp_displacement_full (src/ops/transform.rs:1288-1345) / GpuLiquifyPipeline::warp_into
 * (src/gpu/compute/liquify.rs:176): dst(x,y) = bilinear src(x-dx, y-dy), zero outside. */
int pfe_warp_displacement(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h,
                          const float *disp, uint32_t w, uint32_t h, uint8_t *dst);
int pfe_dev_warp_displacement(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h,
                              const float *disp, uint32_t w, uint32_t h, uint8_t *dst);
/* generate_displacement_from_mesh (:1670-1706; original_points != NULL) and _fast (:1712-1739;
 * original_points == NULL) / GpuMeshWarpDisplacementPipeline::generate_displacement
 * (src/gpu/compute/mesh_warp.rs:131). points: (rows+1)*(cols+1) [x,y] pairs, HOST pointers in
 * both tiers (<= PFE_MESH_MAX_POINTS). */
#define PFE_MESH_MAX_POINTS 256
int pfe_mesh_displacement(pfe_ctx *ctx, const float *original_points, const float *deformed_points,
                          uint32_t cols, uint32_t rows, uint32_t w, uint32_t h, float *out_disp);
int pfe_dev_mesh_displacement(pfe_ctx *ctx, const float *original_points,
                              const float *deformed_points, uint32_t cols, uint32_t rows, uint32_t w,
                              uint32_t h, float *out_disp_dev);
/* warp_mesh_catmull_rom (:1743-1761), fused: the w*h*2 field is never materialised.
 * Band form: produce only output rows [y0, y0+rows_out) of the full w x h result, reading the
 * full source (used by the multi-GPU row-band split); y0=0, rows_out=h for the whole image. */
int pfe_mesh_warp(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h,
                  const float *original_points, const float *deformed_points, uint32_t cols,
                  uint32_t rows, uint32_t w, uint32_t h, uint8_t *dst);
int pfe_dev_mesh_warp(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h,
                      const float *original_points, const float *deformed_points, uint32_t cols,
                      uint32_t rows, uint32_t w, uint32_t h, uint32_t y0, uint32_t rows_out,
                      uint8_t *dst_band);
/* warp_displacement_region (src/ops/transform.rs:1206-1285; the liquify tool's incremental preview): dst = prev
 * everywhere, except inside dirty_rect = {x0, y0, x1, y1} (half-open, clamped like the reference: x0.max(0),
 * (x1 as u32).min(w) - so a negative x1 means "to the right edge"), where it is the displacement warp of src.
 * prev and dst are w*h*4 bytes and may be the same buffer; an inverted rect changes nothing. */
int pfe_warp_displacement_region(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h, const float *disp,
                                 const uint8_t *prev, const int32_t dirty_rect[4], uint32_t w, uint32_t h, uint8_t *dst);
int pfe_dev_warp_displacement_region(pfe_ctx *ctx, const uint8_t *src, uint32_t src_w, uint32_t src_h, const float *disp,
                                     const uint8_t *prev, const int32_t dirty_rect[4], uint32_t w, uint32_t h,
                                     uint8_t *dst);
/* Row-band form of both warps for a canvas split across GPUs: produce output rows
 * [y0, y0+rows_out) of the w x h result from a WINDOW of source rows [src_y0, src_y0+src_nrows) of
 * the src_w x src_h source (own band + halo). disp_band != NULL: displacement warp with the band's
 * rows_out*w*2 field; disp_band == NULL: fused mesh warp (points are HOST pointers). Asynchronous on the
 * context's stream. A bilinear tap inside the image that fell outside the window reads as transparent and
 * sets the context's sticky error flag: pfe_ctx_check_async reports it (PFE_ERR_INVALID_ARG). */
int pfe_dev_warp_band(pfe_ctx *ctx, const uint8_t *src_rows, uint32_t src_w, uint32_t src_h, uint32_t src_y0,
                      uint32_t src_nrows, const float *disp_band, const float *original_points,
                      const float *deformed_points, uint32_t cols, uint32_t rows, uint32_t w, uint32_t h,
                      uint32_t y0, uint32_t rows_out, uint8_t *dst_band);
/* Row-band form of parallel_gaussian_blur (src/ops/filters.rs:242-316) for a canvas split across GPUs
 * (SURVEY §8e).  `ext` is this GPU's band EXTENDED by the halo rows its neighbours sent: ext_rows rows of w
 * pixels, band rows preceded by up to ceil(3 sigma) rows from above and followed by as many from below (fewer
 * only where the band is that close to the true image border, so that clamp-to-edge at ext's first / last row
 * IS the image's).  The f32 intermediate of the extended band lives in the context's scratch:
 *   pfe_dev_gaussian_band_h  filters rows [y0, y0+rows) of ext horizontally into it - call it for the band's own
 *                            rows while the halo is still in flight, then for the halo rows once they have landed;
 *   pfe_dev_gaussian_band_v  filters vertically and writes output rows [y0, y0+rows) of ext (normally the band's
 *                            own rows) to dst_rows (rows*w*4 bytes, row 0 = ext row y0).
 * Both are asynchronous on the context's stream; the scratch is only valid between one sequence of _h calls
 * and the _v call that follows them (any other Gaussian-family call on the context overwrites it).  The
 * result is bit-identical to the band's rows of the unsplit blur in both modes. */
int pfe_dev_gaussian_band_h(pfe_ctx *ctx, const uint8_t *ext, uint32_t w, uint32_t ext_rows, uint32_t y0,
                            uint32_t rows, float sigma, uint32_t flags);
int pfe_dev_gaussian_band_v(pfe_ctx *ctx, uint32_t w, uint32_t ext_rows, uint32_t y0, uint32_t rows, float sigma,
                            uint8_t *dst_rows, uint32_t flags);
/* Halo rows over peer memory (one process per GPU on one NVSwitch node).  No counterpart in the reference,
 * which is single-device (src/gpu/context.rs:123 refuses large sizes, src/cli.rs:159 is a serial loop): these entry
 * points exist so that ONE canvas can be split across GPUs below the ABI.  Instead of flattening a
 * band's edge rows and then sending them, the flatten kernel itself stores them a second time, straight into the
 * neighbour GPU's extended band, and the last CTA to finish releases a flag in the neighbour's memory: the transfer
 * is the kernel's own NVLink stores, tile by tile, and there is no send/receive at all.
 *   pfe_peer_alloc   cudaMalloc (zero-filled) + a 64-byte handle another PROCESS of this node can open;
 *   pfe_peer_open    map a neighbour's allocation here (also enables peer access between the two devices);
 *   pfe_peer_close / pfe_peer_free   unmap / release - close every mapping before its owner frees;
 *   pfe_dev_flatten_peer  pfe_dev_flatten that also writes every result at the same offset from peer_dst and
 *                    then, if peer_flag != NULL, stores flag_value to *peer_flag with release semantics at system
 *                    scope (whoever sees the flag sees the rows).  peer_dst / peer_flag need not be remote;
 *   pfe_dev_peer_signal  the hand-over alone, for rows that were put into the neighbour's buffer by other means (a
 *                    device-to-device copy to the mapped address): everything enqueued on the context's stream before
 *                    it is visible to whoever sees the flag;
 *   pfe_dev_peer_wait  stream-ordered wait until every flags[0..n) (in THIS device's memory, n <= 32) has reached
 *                    value (wrap-around compare: flags are step counters that only grow).  If that takes longer than
 *                    timeout_ms the wait gives up and sets the sticky error pfe_ctx_check_async reports.
 * A buffer a neighbour writes into must not be reused for the next step before the neighbour can know it was read:
 * callers alternate between two buffers, so that the flag of step k+1 orders the reads of step k before the writes
 * of step k+2 (paintfe_b200/dist.py PeerHalo). */
int pfe_peer_alloc(pfe_ctx *ctx, size_t bytes, void **dptr, uint8_t handle[64]);
int pfe_peer_open(pfe_ctx *ctx, const uint8_t handle[64], void **dptr);
int pfe_peer_close(pfe_ctx *ctx, void *dptr);
int pfe_peer_free(pfe_ctx *ctx, void *dptr);
int pfe_dev_flatten_peer(pfe_ctx *ctx, const pfe_layer_desc *layers, uint32_t n_layers, uint32_t w, uint32_t h,
                         const uint8_t *active_chunks, uint8_t *dst, uint8_t *peer_dst, uint32_t *peer_flag,
                         uint32_t flag_value);
int pfe_dev_peer_signal(pfe_ctx *ctx, uint32_t *peer_flag, uint32_t flag_value);
int pfe_dev_peer_wait(pfe_ctx *ctx, const uint32_t *flags, uint32_t n, uint32_t value, uint32_t timeout_ms);
/* Reach of a band's displacement field, for sizing the halo of pfe_dev_warp_band: minmax_dev[0..1] (DEVICE
 * memory, int32) = min / max over the band's rows [y0, y0+rows) of floor(clamp(y - dy, -1, h_total)), with
 * non-finite dy counted as 0. Asynchronous: the result stays on the device so that it can go straight into
 * the ranks' max-reduction. */
int pfe_dev_disp_reach(pfe_ctx *ctx, const float *disp_band, uint32_t w, uint32_t rows, uint32_t y0,
                       uint32_t h_total, int32_t *minmax_dev);
/* DisplacementField::apply_push / expand / contract / twirl (:1051-1200), in place on a
 * w*h*2 field. a0,a1: push = (delta_x, delta_y); twirl = (clockwise ? 1 : 0, -).
 * bbox_out (may be NULL) = (x0, y0, x1, y1) like the reference's return value. */
typedef enum pfe_liquify_kind { PFE_LIQ_PUSH = 0, PFE_LIQ_EXPAND = 1, PFE_LIQ_CONTRACT = 2, PFE_LIQ_TWIRL = 3 } pfe_liquify_kind;
int pfe_liquify(pfe_ctx *ctx, float *field, uint32_t w, uint32_t h, int kind, float center_x,
                float center_y, float radius, float strength, float a0, float a1, int32_t bbox_out[4]);
int pfe_dev_liquify(pfe_ctx *ctx, float *field_dev, uint32_t w, uint32_t h, int kind, float center_x,
                    float center_y, float radius, float strength, float a0, float a1,
                    int32_t bbox_out[4]);

/* -- brush stamps ------------------------------------------------------------------------
 * ToolsPanel::draw_circle_no_dirty (src/ui/panels/tools/behavior/raster/brush_render.rs:135-400)
 * for the circle tip in every BrushMode and the eraser, plus rebuild_brush_lut (:27-50) and
 * draw_line_no_dirty's stamp placement (:762-838). One launch applies a whole list of stamps
 * in order, in place. */
typedef struct pfe_brush_desc {
    float size;          /* ToolProperties::size (diameter) */
    float hardness;
    float flow;
    int32_t anti_aliased;
    float color[4];      /* straight RGBA, 0..1 (primary or secondary, chosen by the caller) */
    int32_t is_eraser;
    int32_t mode;        /* BrushMode: 0 Normal, 1 Dodge, 2 Burn, 3 Sponge (brush_render.rs:362-395) */
} pfe_brush_desc;
int pfe_brush_stamps(pfe_ctx *ctx, uint8_t *image, uint32_t w, uint32_t h, const pfe_brush_desc *brush,
                     const float *centres_xy, uint32_t n_stamps, const uint8_t *selection_mask);
int pfe_dev_brush_stamps(pfe_ctx *ctx, uint8_t *image_dev, uint32_t w, uint32_t h,
                         const pfe_brush_desc *brush, const float *centres_xy_host, uint32_t n_stamps,
                         const uint8_t *selection_mask_dev);
/* Stamp centres of draw_line_no_dirty (circle tip: 1 px stepping). Returns the count written
 * (<= cap). Host-only helper. */
int pfe_brush_line_centres(uint32_t w, uint32_t h, float x0, float y0, float x1, float y1,
                           float *centres_xy, int cap);
void pfe_brush_lut(const pfe_brush_desc *brush, uint8_t lut[256]);

/* -- tiles -----------------------------------------------------------------------------
 * TiledImage marshalling (src/canvas/tiled_image.rs). chunk_table[cy*chunks_per_row+cx] points
 * to a 64*64*4-byte tile or is NULL for an unpopulated chunk (`Vec<Option<Arc<RgbaImage>>>`,
 * :2-7). Host-side, multithreaded memcpy; no device work. */
int pfe_tiles_to_flat(const uint8_t *const *chunk_table, uint32_t w, uint32_t h, uint8_t *flat); /* to_rgba_image :271 */
/* from_rgba_image :50-104: writes occupancy (1 where the chunk has any alpha != 0) and, when
 * tiles_out != NULL, copies populated chunks into tiles_out + idx*16384 (zero-padded edges). */
int pfe_flat_to_tiles(const uint8_t *flat, uint32_t w, uint32_t h, uint8_t *occupancy, uint8_t *tiles_out);

/* -- fused headline pipeline -------------------------------------------------------------
 * composite() followed by parallel_gaussian_blur on the result, without leaving the device
 * (cli.rs:282-285 + scripting apply_blur). Equivalent to pfe_flatten then pfe_gaussian_blur. */
int pfe_flatten_gaussian(pfe_ctx *ctx, const pfe_layer_desc *layers, uint32_t n_layers, uint32_t w,
                         uint32_t h, const uint8_t *active_chunks, float sigma, uint8_t *dst,
                         uint32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* PFE_B200_H */
