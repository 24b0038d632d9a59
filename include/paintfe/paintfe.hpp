// paintfe.hpp — host-side mirror of PaintFE's compositor / ops interface over libpfe_b200.so.
//
// PaintFE is Rust; this image has no Rust toolchain, so the host side above the C ABI is written in
// C++17 with the reference's own names, argument meaning and error behaviour, so that a call site
// (the Rhai Effect API in src/ops/scripting.rs, cli::run_one in src/cli.rs, the `*_gpu` wrappers)
// reads the same and the reference's tests port line by line (tests/cpp/mirror_tests.cpp).
// INTEGRATION.md shows the Rust `extern "C"` binding that does the same job inside PaintFE itself.
//
// Every compute call goes to the CUDA library; there is no CPU implementation behind this header.
// Errors: the reference's ops are infallible (`-> RgbaImage`) and early-return on bad layer indices
// (e.g. src/ops/filters.rs:16-18); here a library failure throws paintfe::Error, an out-of-range
// layer index is a silent no-op exactly as in the reference.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../pfe_b200.h"

namespace paintfe {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// One pfe_ctx per thread, like GpuRenderer's single wgpu device (src/gpu/context.rs:9-15).
class Engine {
public:
    explicit Engine(int device = 0) {
        int rc = pfe_ctx_create(device, &ctx_);
        if (rc != PFE_OK) throw Error(rc, "pfe_ctx_create failed (no CUDA device? the engine has no CPU path)");
    }
    ~Engine() { if (ctx_) pfe_ctx_destroy(ctx_); }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    pfe_ctx *ctx() const { return ctx_; }
    void check(int rc, const char *what) const {
        if (rc != PFE_OK) throw Error(rc, std::string(what) + ": " + pfe_last_error(ctx_));
    }
    static Engine &current() {
        thread_local Engine e(device_index());
        return e;
    }
    static int &device_index() { static int d = 0; return d; }

private:
    pfe_ctx *ctx_ = nullptr;
};

// ---- image::RgbaImage / image::GrayImage (image 0.25.9) as plain containers -------------------
struct Rgba { uint8_t v[4]; uint8_t &operator[](int i) { return v[i]; } uint8_t operator[](int i) const { return v[i]; }
              bool operator==(const Rgba &o) const { return std::memcmp(v, o.v, 4) == 0; } };

class RgbaImage {
public:
    RgbaImage() = default;
    RgbaImage(uint32_t w, uint32_t h) : w_(w), h_(h), d_((size_t)w * h * 4, 0) {}
    static RgbaImage from_pixel(uint32_t w, uint32_t h, Rgba p) {
        RgbaImage im(w, h);
        for (size_t i = 0; i < (size_t)w * h; i++) std::memcpy(&im.d_[i * 4], p.v, 4);
        return im;
    }
    static std::optional<RgbaImage> from_raw(uint32_t w, uint32_t h, std::vector<uint8_t> raw) {
        if (raw.size() != (size_t)w * h * 4) return std::nullopt;
        RgbaImage im;
        im.w_ = w; im.h_ = h; im.d_ = std::move(raw);
        return im;
    }
    uint32_t width() const { return w_; }
    uint32_t height() const { return h_; }
    const std::vector<uint8_t> &as_raw() const { return d_; }
    std::vector<uint8_t> &as_mut() { return d_; }
    std::vector<uint8_t> into_raw() && { return std::move(d_); }
    Rgba get_pixel(uint32_t x, uint32_t y) const { Rgba p; std::memcpy(p.v, &d_[((size_t)y * w_ + x) * 4], 4); return p; }
    void put_pixel(uint32_t x, uint32_t y, Rgba p) { std::memcpy(&d_[((size_t)y * w_ + x) * 4], p.v, 4); }
    bool operator==(const RgbaImage &o) const { return w_ == o.w_ && h_ == o.h_ && d_ == o.d_; }

private:
    uint32_t w_ = 0, h_ = 0;
    std::vector<uint8_t> d_;
};

class GrayImage {
public:
    GrayImage() = default;
    GrayImage(uint32_t w, uint32_t h) : w_(w), h_(h), d_((size_t)w * h, 0) {}
    uint32_t width() const { return w_; }
    uint32_t height() const { return h_; }
    const std::vector<uint8_t> &as_raw() const { return d_; }
    uint8_t get_pixel(uint32_t x, uint32_t y) const { return d_[(size_t)y * w_ + x]; }
    void put_pixel(uint32_t x, uint32_t y, uint8_t v) { d_[(size_t)y * w_ + x] = v; }

private:
    uint32_t w_ = 0, h_ = 0;
    std::vector<uint8_t> d_;
};

namespace detail {
// A GrayImage mask only applies when it has the image's size; the reference's per-pixel
// `x < mask_w && y < mask_h` guard (e.g. effects.rs:36-40) is the identity in that case.
inline const uint8_t *mask_ptr(const GrayImage *m, uint32_t w, uint32_t h) {
    if (!m) return nullptr;
    if (m->width() != w || m->height() != h) throw Error(PFE_ERR_INVALID_ARG, "selection mask must have the image's size");
    return m->as_raw().data();
}
}  // namespace detail

// ================================================================================================
namespace canvas {

constexpr uint32_t CHUNK_SIZE = PFE_CHUNK_SIZE;  // src/canvas/defs.rs:7

// src/canvas/layers.rs:2-29, ids :125-185
enum class BlendMode : uint8_t {
    Normal = 0, Multiply, Screen, Additive, Reflect, Glow, ColorBurn, ColorDodge, Overlay, Difference, Negation,
    Lighten, Darken, Xor, Overwrite, HardLight, SoftLight, Exclusion, Subtract, Divide, LinearBurn, VividLight,
    LinearLight, PinLight, HardMix
};
inline uint8_t to_u8(BlendMode m) { return (uint8_t)m; }
inline BlendMode from_u8(uint8_t v) { return v <= 24 ? (BlendMode)v : BlendMode::Normal; }

// src/canvas/tiled_image.rs: sparse 64x64 chunks, `Vec<Option<Arc<RgbaImage>>>`, copy-on-write.
class TiledImage {
public:
    using Chunk = std::vector<uint8_t>;  // 64*64*4 bytes
    TiledImage() : TiledImage(1, 1) {}
    TiledImage(uint32_t width, uint32_t height) {  // TiledImage::new, :13-37
        uint64_t total = (uint64_t)width * height;
        if (total > 256000000ull || width == 0 || height == 0) { width = 1; height = 1; }
        width_ = width; height_ = height;
        chunks_per_row_ = (width + CHUNK_SIZE - 1) / CHUNK_SIZE;
        chunks_.assign((size_t)chunks_per_row_ * ((height + CHUNK_SIZE - 1) / CHUNK_SIZE), nullptr);
    }
    static TiledImage from_rgba_image(const RgbaImage &src) { return from_raw_rgba(src.width(), src.height(), src.as_raw().data()); }
    static TiledImage from_raw_rgba(uint32_t w, uint32_t h, const uint8_t *data) {  // :50-161
        TiledImage t(w, h);
        if (t.width_ != w || t.height_ != h) return t;
        std::vector<uint8_t> occ(t.chunks_.size()), tiles(t.chunks_.size() * kChunkBytes);
        if (pfe_flat_to_tiles(data, w, h, occ.data(), tiles.data()) != PFE_OK) throw Error(PFE_ERR_INVALID_ARG, "pfe_flat_to_tiles");
        for (size_t i = 0; i < occ.size(); i++)
            if (occ[i]) t.chunks_[i] = std::make_shared<Chunk>(tiles.begin() + i * kChunkBytes, tiles.begin() + (i + 1) * kChunkBytes);
        return t;
    }
    RgbaImage to_rgba_image() const {  // :271-293
        std::vector<const uint8_t *> table(chunks_.size());
        for (size_t i = 0; i < chunks_.size(); i++) table[i] = chunks_[i] ? chunks_[i]->data() : nullptr;
        RgbaImage out(width_, height_);
        if (pfe_tiles_to_flat(table.data(), width_, height_, out.as_mut().data()) != PFE_OK) throw Error(PFE_ERR_INVALID_ARG, "pfe_tiles_to_flat");
        return out;
    }
    uint32_t width() const { return width_; }
    uint32_t height() const { return height_; }
    Rgba get_pixel(uint32_t x, uint32_t y) const {
        Rgba p{{0, 0, 0, 0}};
        if (x >= width_ || y >= height_) return p;
        const auto &c = chunks_[(size_t)(y / CHUNK_SIZE) * chunks_per_row_ + x / CHUNK_SIZE];
        if (c) std::memcpy(p.v, c->data() + ((size_t)(y % CHUNK_SIZE) * CHUNK_SIZE + x % CHUNK_SIZE) * 4, 4);
        return p;
    }
    void put_pixel(uint32_t x, uint32_t y, Rgba p) {
        if (x >= width_ || y >= height_) return;
        Chunk &c = ensure_chunk_mut(x / CHUNK_SIZE, y / CHUNK_SIZE);
        std::memcpy(c.data() + ((size_t)(y % CHUNK_SIZE) * CHUNK_SIZE + x % CHUNK_SIZE) * 4, p.v, 4);
    }
    const Chunk *get_chunk(uint32_t cx, uint32_t cy) const {  // :853
        size_t i = (size_t)cy * chunks_per_row_ + cx;
        return (cx < chunks_per_row_ && i < chunks_.size() && chunks_[i]) ? chunks_[i].get() : nullptr;
    }
    Chunk &ensure_chunk_mut(uint32_t cx, uint32_t cy) {  // :868, Arc::make_mut copy-on-write
        auto &c = chunks_[(size_t)cy * chunks_per_row_ + cx];
        if (!c) c = std::make_shared<Chunk>(kChunkBytes, 0);
        else if (c.use_count() > 1) c = std::make_shared<Chunk>(*c);
        return *c;
    }
    std::vector<std::pair<uint32_t, uint32_t>> chunk_keys() const {  // :884
        std::vector<std::pair<uint32_t, uint32_t>> k;
        for (size_t i = 0; i < chunks_.size(); i++)
            if (chunks_[i]) k.emplace_back((uint32_t)(i % chunks_per_row_), (uint32_t)(i / chunks_per_row_));
        return k;
    }
    std::vector<const uint8_t *> chunk_table() const {  // `chunk.as_raw().as_ptr()` per entry, null where None
        std::vector<const uint8_t *> t(chunks_.size());
        for (size_t i = 0; i < chunks_.size(); i++) t[i] = chunks_[i] ? chunks_[i]->data() : nullptr;
        return t;
    }
    std::vector<uint8_t> occupancy() const {
        std::vector<uint8_t> o(chunks_.size());
        for (size_t i = 0; i < chunks_.size(); i++) o[i] = chunks_[i] ? 1 : 0;
        return o;
    }

private:
    static constexpr size_t kChunkBytes = (size_t)CHUNK_SIZE * CHUNK_SIZE * 4;
    uint32_t width_, height_, chunks_per_row_;
    std::vector<std::shared_ptr<Chunk>> chunks_;
};

// src/canvas/layers.rs:262-325
struct AdjustmentLayerData {
    enum Kind { Exposure = 1, BrightnessContrast = 2, Invert = 3, ChannelMixer = 4 } kind = Invert;
    float ev = 0, brightness = 0, contrast = 0;
    std::array<float, 4> red{1, 0, 0, 0}, green{0, 1, 0, 0}, blue{0, 0, 1, 0}, alpha{0, 0, 0, 1};
};

// src/canvas/layers.rs:389-421 (fields on the compositing path)
struct Layer {
    std::string name;
    bool visible = true;
    float opacity = 1.0f;
    BlendMode blend_mode = BlendMode::Normal;
    TiledImage pixels;
    std::optional<TiledImage> mask;  // alpha = concealment, 0 = reveal (layers.rs:395-397)
    bool mask_enabled = false;
    std::optional<AdjustmentLayerData> adjustment;  // LayerContent::Adjustment
    std::optional<uint64_t> folder_id;              // layers.rs:392
    Layer(std::string n, uint32_t w, uint32_t h, Rgba fill) : name(std::move(n)), pixels(w, h) {
        if (fill[3] > 0) pixels = TiledImage::from_rgba_image(RgbaImage::from_pixel(w, h, fill));
    }
};

// src/canvas/layers.rs:378-387
struct LayerFolder {
    uint64_t id = 0;
    std::string name;
    bool visible = true, collapsed = false;
};

// src/canvas/canvas_state.rs: the compositing subset of CanvasState
class CanvasState {
public:
    uint32_t width, height;
    std::vector<Layer> layers;
    std::vector<LayerFolder> layer_folders;
    // canvas_state.rs:216-227: hidden itself, or a member of a hidden folder
    bool layer_effectively_visible(size_t idx) const {
        if (idx >= layers.size() || !layers[idx].visible) return false;
        if (!layers[idx].folder_id) return true;
        for (const LayerFolder &f : layer_folders)
            if (f.id == *layers[idx].folder_id) return f.visible;
        return true;
    }
    std::optional<GrayImage> selection_mask;
    CanvasState(uint32_t w, uint32_t h) : width(w), height(h) {  // CanvasState::new: one white background layer
        layers.emplace_back("Background", w, h, Rgba{{255, 255, 255, 255}});
    }
    // CanvasState::composite, canvas_state.rs:482-698: the chunk tables go to the library as they are
    // (no densification on the host); only populated chunks are uploaded.
    RgbaImage composite() const {
        std::vector<std::vector<const uint8_t *>> tables;
        tables.reserve(layers.size() * 2);
        std::vector<pfe_tile_layer_desc> descs;
        for (size_t li = 0; li < layers.size(); li++) {
            const Layer &L = layers[li];
            const bool vis = layer_effectively_visible(li);
            pfe_tile_layer_desc d{};
            d.opacity = L.opacity;
            d.blend = to_u8(L.blend_mode);
            d.visible = vis ? 1 : 0;
            if (L.adjustment) {
                const auto &a = *L.adjustment;
                d.kind = (uint8_t)a.kind;
                if (a.kind == AdjustmentLayerData::Exposure) d.adj[0] = std::pow(2.0f, a.ev);  // 2.0f32.powf(ev)
                if (a.kind == AdjustmentLayerData::BrightnessContrast) { d.adj[0] = a.brightness; d.adj[1] = a.contrast; }
                if (a.kind == AdjustmentLayerData::ChannelMixer)
                    for (int i = 0; i < 4; i++) { d.adj[i] = a.red[i]; d.adj[4 + i] = a.green[i]; d.adj[8 + i] = a.blue[i]; d.adj[12 + i] = a.alpha[i]; }
            } else if (vis) {
                tables.push_back(L.pixels.chunk_table());
                d.chunks = tables.back().data();
                if (L.mask_enabled && L.mask) {
                    tables.push_back(L.mask->chunk_table());
                    d.mask_chunks = tables.back().data();
                }
            } else {
                d.visible = 0;
            }
            descs.push_back(d);
        }
        RgbaImage out(width, height);
        Engine &e = Engine::current();
        e.check(pfe_flatten_tiles(e.ctx(), descs.data(), (uint32_t)descs.size(), width, height, out.as_mut().data()), "pfe_flatten_tiles");
        return out;
    }
    // The same composite through the dense entry point (pfe_flatten + active-chunk bitmap): what a caller
    // that already holds flat layers uses; kept so both boundaries stay covered by the ported tests.
    RgbaImage composite_dense() const {
        const size_t nch = (size_t)((width + CHUNK_SIZE - 1) / CHUNK_SIZE) * ((height + CHUNK_SIZE - 1) / CHUNK_SIZE);
        std::vector<uint8_t> active(nch, 0);
        std::vector<RgbaImage> flats;
        std::vector<std::vector<uint8_t>> masks;
        std::vector<pfe_layer_desc> descs;
        flats.reserve(layers.size());
        masks.reserve(layers.size());
        bool any_adjustment = false;
        for (size_t li = 0; li < layers.size(); li++) {
            const Layer &L = layers[li];
            const bool vis = layer_effectively_visible(li);
            pfe_layer_desc d{};
            d.opacity = L.opacity;
            d.blend = to_u8(L.blend_mode);
            d.visible = vis ? 1 : 0;
            if (L.adjustment) {
                const auto &a = *L.adjustment;
                d.kind = (uint8_t)a.kind;
                if (a.kind == AdjustmentLayerData::Exposure) d.adj[0] = std::pow(2.0f, a.ev);  // 2.0f32.powf(ev)
                if (a.kind == AdjustmentLayerData::BrightnessContrast) { d.adj[0] = a.brightness; d.adj[1] = a.contrast; }
                if (a.kind == AdjustmentLayerData::ChannelMixer)
                    for (int i = 0; i < 4; i++) { d.adj[i] = a.red[i]; d.adj[4 + i] = a.green[i]; d.adj[8 + i] = a.blue[i]; d.adj[12 + i] = a.alpha[i]; }
                any_adjustment = any_adjustment || vis;
            } else if (vis) {
                flats.push_back(L.pixels.to_rgba_image());
                d.rgba = flats.back().as_raw().data();
                auto occ = L.pixels.occupancy();
                for (size_t i = 0; i < nch && i < occ.size(); i++) active[i] |= occ[i];  // :529-550
                if (L.mask_enabled && L.mask) {
                    RgbaImage m = L.mask->to_rgba_image();
                    masks.emplace_back((size_t)width * height);
                    for (size_t i = 0; i < masks.back().size(); i++) masks.back()[i] = m.as_raw()[i * 4 + 3];
                    d.mask = masks.back().data();
                }
            } else {
                d.visible = 0;
            }
            descs.push_back(d);
        }
        RgbaImage out(width, height);
        Engine &e = Engine::current();
        e.check(pfe_flatten(e.ctx(), descs.data(), (uint32_t)descs.size(), width, height,
                            any_adjustment ? active.data() : nullptr, out.as_mut().data()), "pfe_flatten");
        return out;
    }
};

}  // namespace canvas

// ================================================================================================
namespace ops {
using canvas::CanvasState;
using canvas::TiledImage;

namespace detail {
template <class F>
inline RgbaImage img_op(const RgbaImage &flat, const GrayImage *mask, const char *what, F f) {
    if (flat.width() == 0 || flat.height() == 0) return flat;  // `if w == 0 || h == 0 { return flat.clone() }`
    RgbaImage out(flat.width(), flat.height());
    Engine &e = Engine::current();
    e.check(f(e.ctx(), flat.as_raw().data(), flat.width(), flat.height(),
              paintfe::detail::mask_ptr(mask, flat.width(), flat.height()), out.as_mut().data()), what);
    return out;
}
inline void commit_to_layer(CanvasState &s, size_t idx, const RgbaImage &r) {  // effects.rs:100-106
    if (idx >= s.layers.size()) return;
    s.layers[idx].pixels = TiledImage::from_rgba_image(r);
}
}  // namespace detail

namespace filters {
// src/ops/filters.rs:130 / :237
inline RgbaImage blur_with_selection_pub(const RgbaImage &flat, float sigma, const GrayImage *mask, bool exact = true) {
    return detail::img_op(flat, mask, "pfe_gaussian_blur", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_gaussian_blur(c, s, w, h, sigma, m, d, exact ? PFE_GAUSS_EXACT : 0u);
    });
}
inline RgbaImage parallel_gaussian_blur_pub(const RgbaImage &src, float sigma, bool exact = true) { return blur_with_selection_pub(src, sigma, nullptr, exact); }
// src/ops/filters.rs:15-25
inline void gaussian_blur_layer(CanvasState &state, size_t layer_idx, float sigma) {
    if (layer_idx >= state.layers.size()) return;
    RgbaImage flat = state.layers[layer_idx].pixels.to_rgba_image();
    detail::commit_to_layer(state, layer_idx, blur_with_selection_pub(flat, sigma, state.selection_mask ? &*state.selection_mask : nullptr));
}
}  // namespace filters

namespace effects {
// src/ops/effects/blur.rs:144, :233; noise.rs:357; stylize.rs:96, :170
inline RgbaImage motion_blur_core(const RgbaImage &flat, float angle_deg, float distance, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_motion_blur", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_motion_blur(c, s, w, h, angle_deg, distance, m, d);
    });
}
inline RgbaImage box_blur_core(const RgbaImage &flat, float radius, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_box_blur", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_box_blur(c, s, w, h, radius, m, d);
    });
}
inline RgbaImage median_core(const RgbaImage &flat, uint32_t radius, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_median", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_median(c, s, w, h, radius, m, d);
    });
}
inline RgbaImage sharpen_core(const RgbaImage &flat, float amount, float radius, const GrayImage *mask, bool exact = true) {
    return detail::img_op(flat, mask, "pfe_sharpen", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_sharpen(c, s, w, h, amount, radius, m, d, exact ? PFE_GAUSS_EXACT : 0u);
    });
}
inline RgbaImage vignette_core(const RgbaImage &flat, float amount, float softness, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_vignette", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_vignette(c, s, w, h, amount, softness, m, d);
    });
}
// widened scope: stylize.rs:26, distort.rs:333/396/460, noise.rs:73/172
inline RgbaImage glow_core(const RgbaImage &flat, float radius, float intensity, const GrayImage *mask, bool exact = true) {
    return detail::img_op(flat, mask, "pfe_glow", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_glow(c, s, w, h, radius, intensity, m, d, exact ? PFE_GAUSS_EXACT : 0u);
    });
}
inline RgbaImage pixelate_core(const RgbaImage &flat, uint32_t block_size, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_pixelate", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_pixelate(c, s, w, h, block_size, m, d);
    });
}
inline RgbaImage bulge_core_at(const RgbaImage &flat, float amount, std::pair<float, float> origin, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_bulge", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_bulge(c, s, w, h, amount, origin.first, origin.second, m, d);
    });
}
inline RgbaImage bulge_core(const RgbaImage &flat, float amount, const GrayImage *mask) { return bulge_core_at(flat, amount, {0.5f, 0.5f}, mask); }
inline RgbaImage twist_core_at(const RgbaImage &flat, float angle_deg, std::pair<float, float> origin, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_twist", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_twist(c, s, w, h, angle_deg, origin.first, origin.second, m, d);
    });
}
inline RgbaImage twist_core(const RgbaImage &flat, float angle_deg, const GrayImage *mask) { return twist_core_at(flat, angle_deg, {0.5f, 0.5f}, mask); }
enum class NoiseType { Uniform = 0, Gaussian = 1, Perlin = 2 };
inline RgbaImage add_noise_core(const RgbaImage &flat, float amount, NoiseType t, bool monochrome, uint32_t seed, float scale,
                                uint32_t octaves, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_add_noise", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_add_noise(c, s, w, h, amount, (int)t, monochrome ? 1 : 0, seed, scale, octaves, m, d);
    });
}
inline RgbaImage reduce_noise_core(const RgbaImage &flat, float strength, uint32_t radius, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_reduce_noise", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_reduce_noise(c, s, w, h, strength, radius, m, d);
    });
}
// the rest of src/ops/effects/: artistic.rs:31/123/266, contours.rs:56, distort.rs:26/248, stylize.rs:242,
// blur.rs:22/322, render.rs:52/114/220/403, glitch.rs:44/142
inline RgbaImage ink_core(const RgbaImage &flat, float edge_strength, float threshold, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_ink", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_ink(c, s, w, h, edge_strength, threshold, m, d);
    });
}
inline RgbaImage oil_painting_core(const RgbaImage &flat, uint32_t radius, uint32_t levels, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_oil_painting", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_oil_painting(c, s, w, h, radius, levels, m, d);
    });
}
enum class ColorFilterMode { Multiply = 0, Screen = 1, Overlay = 2, SoftLight = 3 };
inline RgbaImage color_filter_core(const RgbaImage &flat, std::array<uint8_t, 4> filter_color, float intensity, ColorFilterMode mode, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_color_filter", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_color_filter(c, s, w, h, filter_color.data(), intensity, (int)mode, m, d);
    });
}
inline RgbaImage contours_core(const RgbaImage &flat, float scale, float frequency, float line_width, std::array<uint8_t, 4> line_color, uint32_t seed,
                              uint32_t octaves, float blend, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_contours", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_contours(c, s, w, h, scale, frequency, line_width, line_color.data(), seed, octaves, blend, m, d);
    });
}
inline RgbaImage crystallize_core(const RgbaImage &flat, float cell_size, uint32_t seed, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_crystallize", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_crystallize(c, s, w, h, cell_size, seed, m, d);
    });
}
inline RgbaImage dents_core(const RgbaImage &flat, float scale, float amount, uint32_t seed, uint32_t octaves, float roughness, bool pinch, bool wrap,
                           const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_dents", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_dents(c, s, w, h, scale, amount, seed, octaves, roughness, pinch ? 1 : 0, wrap ? 1 : 0, m, d);
    });
}
enum class HalftoneShape { Circle = 0, Square = 1, Diamond = 2, Line = 3 };
inline RgbaImage halftone_core(const RgbaImage &flat, float dot_size, float angle_deg, HalftoneShape shape, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_halftone", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_halftone(c, s, w, h, dot_size, angle_deg, (int)shape, m, d);
    });
}
inline RgbaImage bokeh_blur_core(const RgbaImage &flat, float radius, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_bokeh_blur", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_bokeh_blur(c, s, w, h, radius, m, d);
    });
}
inline RgbaImage zoom_blur_core(const RgbaImage &flat, float center_x, float center_y, float strength, uint32_t samples, std::array<float, 4> tint_color,
                               float tint_strength, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_zoom_blur", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_zoom_blur(c, s, w, h, center_x, center_y, strength, samples, tint_color.data(), tint_strength, m, d);
    });
}
enum class GridStyle { Lines = 0, Checkerboard = 1 };
inline RgbaImage grid_core(const RgbaImage &flat, uint32_t cell_w, uint32_t cell_h, uint32_t line_width, std::array<uint8_t, 4> color, GridStyle style,
                          float opacity, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_grid", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_grid(c, s, w, h, cell_w, cell_h, line_width, color.data(), (int)style, opacity, m, d);
    });
}
inline RgbaImage canvas_border_core(const RgbaImage &flat, uint32_t width, std::array<uint8_t, 4> color, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_canvas_border", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_canvas_border(c, s, w, h, width, color.data(), m, d);
    });
}
inline RgbaImage shadow_core(const RgbaImage &flat, int32_t offset_x, int32_t offset_y, float blur_radius, bool widen_radius, std::array<uint8_t, 4> color,
                            float opacity, const GrayImage *mask, bool exact = true) {
    return detail::img_op(flat, mask, "pfe_drop_shadow", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_drop_shadow(c, s, w, h, offset_x, offset_y, blur_radius, widen_radius ? 1 : 0, color.data(), opacity, m, d, exact ? PFE_GAUSS_EXACT : 0u);
    });
}
enum class OutlineMode { Outside = 0, Inside = 1, Center = 2 };
inline RgbaImage outline_core(const RgbaImage &flat, uint32_t width, std::array<uint8_t, 4> color, OutlineMode mode, bool anti_alias, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_outline", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_outline(c, s, w, h, width, color.data(), (int)mode, anti_alias ? 1 : 0, m, d);
    });
}
inline RgbaImage pixel_drag_core(const RgbaImage &flat, uint32_t seed, float amount, uint32_t distance, float direction, const GrayImage *mask) {
    return detail::img_op(flat, mask, "pfe_pixel_drag", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_pixel_drag(c, s, w, h, seed, amount, distance, direction, m, d);
    });
}
inline RgbaImage rgb_displace_core(const RgbaImage &flat, std::pair<int32_t, int32_t> r_off, std::pair<int32_t, int32_t> g_off,
                                   std::pair<int32_t, int32_t> b_off, const GrayImage *mask) {
    const int32_t off[6] = {r_off.first, r_off.second, g_off.first, g_off.second, b_off.first, b_off.second};
    return detail::img_op(flat, mask, "pfe_rgb_displace", [&](pfe_ctx *c, const uint8_t *s, uint32_t w, uint32_t h, const uint8_t *m, uint8_t *d) {
        return pfe_rgb_displace(c, s, w, h, off, m, d);
    });
}
}  // namespace effects

namespace adjustments {
namespace detail2 {
inline RgbaImage adjust_flat(const RgbaImage &flat, const pfe_adjust_desc &d, const GrayImage *mask, const std::vector<uint8_t> *occ) {
    if (flat.width() == 0 || flat.height() == 0) return flat;
    RgbaImage out(flat.width(), flat.height());
    Engine &e = Engine::current();
    e.check(pfe_adjust(e.ctx(), flat.as_raw().data(), flat.width(), flat.height(), &d,
                       paintfe::detail::mask_ptr(mask, flat.width(), flat.height()), occ ? occ->data() : nullptr, out.as_mut().data()), "pfe_adjust");
    return out;
}
inline pfe_adjust_desc desc(int op, std::initializer_list<float> p = {}, const uint8_t *luts = nullptr) {
    pfe_adjust_desc d{};
    d.op = op;
    int i = 0;
    for (float v : p) d.params[i++] = v;
    d.luts = luts;
    return d;
}
// apply_pixel_transform (adjustments.rs:21-42): in place on populated chunks only
inline void in_place(CanvasState &s, size_t idx, const pfe_adjust_desc &d) {
    if (idx >= s.layers.size()) return;
    auto occ = s.layers[idx].pixels.occupancy();
    RgbaImage flat = s.layers[idx].pixels.to_rgba_image();
    RgbaImage out = adjust_flat(flat, d, s.selection_mask ? &*s.selection_mask : nullptr, &occ);
    // chunk population is unchanged by an in-place transform: rebuild tiles only where populated
    TiledImage t = s.layers[idx].pixels;
    for (auto [cx, cy] : t.chunk_keys()) {
        auto &c = t.ensure_chunk_mut(cx, cy);
        uint32_t x0 = cx * canvas::CHUNK_SIZE, y0 = cy * canvas::CHUNK_SIZE;
        uint32_t cw = std::min(canvas::CHUNK_SIZE, out.width() - x0), ch = std::min(canvas::CHUNK_SIZE, out.height() - y0);
        for (uint32_t ly = 0; ly < ch; ly++)
            std::memcpy(c.data() + (size_t)ly * canvas::CHUNK_SIZE * 4, out.as_raw().data() + ((size_t)(y0 + ly) * out.width() + x0) * 4, (size_t)cw * 4);
    }
    s.layers[idx].pixels = std::move(t);
}
// apply_pixel_transform_from_flat (adjustments.rs:46-108)
inline void from_flat(CanvasState &s, size_t idx, const RgbaImage &original_flat, const pfe_adjust_desc &d) {
    if (idx >= s.layers.size()) return;
    if (original_flat.width() == 0 || original_flat.height() == 0) return;
    s.layers[idx].pixels = TiledImage::from_rgba_image(adjust_flat(original_flat, d, s.selection_mask ? &*s.selection_mask : nullptr, nullptr));
}
}  // namespace detail2

inline void invert_colors(CanvasState &s, size_t i) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_INVERT)); }                    // :115
inline void sepia(CanvasState &s, size_t i) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_SEPIA)); }                             // :133
inline void invert_alpha(CanvasState &s, size_t i) {                                                                               // :122
    if (i >= s.layers.size()) return;
    detail2::from_flat(s, i, s.layers[i].pixels.to_rgba_image(), detail2::desc(PFE_ADJ_INVERT_ALPHA));
}
inline void brightness_contrast(CanvasState &s, size_t i, float b, float c) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_BRIGHTNESS_CONTRAST, {b, c})); }
inline void brightness_contrast_from_flat(CanvasState &s, size_t i, float b, float c, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_BRIGHTNESS_CONTRAST, {b, c})); }
inline void hue_saturation_lightness(CanvasState &s, size_t i, float h, float sat, float l) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_HSL, {h, sat, l})); }
inline void hue_saturation_lightness_from_flat(CanvasState &s, size_t i, float h, float sat, float l, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_HSL, {h, sat, l})); }
inline void exposure_adjust(CanvasState &s, size_t i, float ev) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_EXPOSURE, {std::pow(2.0f, ev)})); }
inline void exposure_from_flat(CanvasState &s, size_t i, float ev, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_EXPOSURE, {std::pow(2.0f, ev)})); }
inline void levels_from_flat(CanvasState &s, size_t i, float ib, float iw, float g, float ob, float ow, const RgbaImage &flat) {    // :411
    uint8_t lut[256];
    pfe_build_levels_lut(ib, iw, g, ob, ow, lut);
    detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_LUT_RGB, {}, lut));
}
// adjustments.rs:374-421, :517-545, :1240-1444
inline void highlights_shadows(CanvasState &s, size_t i, float shadows, float highlights) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_HIGHLIGHTS_SHADOWS, {shadows, highlights})); }
inline void highlights_shadows_from_flat(CanvasState &s, size_t i, float shadows, float highlights, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_HIGHLIGHTS_SHADOWS, {shadows, highlights})); }
inline void temperature_tint(CanvasState &s, size_t i, float t, float tint) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_TEMPERATURE_TINT, {t, tint})); }
inline void temperature_tint_from_flat(CanvasState &s, size_t i, float t, float tint, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_TEMPERATURE_TINT, {t, tint})); }
inline void threshold(CanvasState &s, size_t i, float level) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_THRESHOLD, {level})); }
inline void threshold_from_flat(CanvasState &s, size_t i, float level, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_THRESHOLD, {level})); }
inline void posterize(CanvasState &s, size_t i, uint32_t levels) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_POSTERIZE, {(float)std::max(levels, 2u)})); }
inline void posterize_from_flat(CanvasState &s, size_t i, uint32_t levels, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_POSTERIZE, {(float)std::max(levels, 2u)})); }
using Rgb3 = std::array<float, 3>;
inline pfe_adjust_desc color_balance_desc(Rgb3 sh, Rgb3 mid, Rgb3 hi) { return detail2::desc(PFE_ADJ_COLOR_BALANCE, {sh[0], sh[1], sh[2], mid[0], mid[1], mid[2], hi[0], hi[1], hi[2]}); }
inline void color_balance(CanvasState &s, size_t i, Rgb3 sh, Rgb3 mid, Rgb3 hi) { detail2::in_place(s, i, color_balance_desc(sh, mid, hi)); }
inline void color_balance_from_flat(CanvasState &s, size_t i, Rgb3 sh, Rgb3 mid, Rgb3 hi, const RgbaImage &flat) { detail2::from_flat(s, i, flat, color_balance_desc(sh, mid, hi)); }
using GradientLut = std::array<std::array<uint8_t, 4>, 256>;
inline void gradient_map(CanvasState &s, size_t i, const GradientLut &lut) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_GRADIENT_MAP, {}, &lut[0][0])); }
inline void gradient_map_from_flat(CanvasState &s, size_t i, const GradientLut &lut, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_GRADIENT_MAP, {}, &lut[0][0])); }
inline void black_and_white(CanvasState &s, size_t i, float rw, float gw, float bw) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_BLACK_AND_WHITE, {rw, gw, bw})); }
inline void black_and_white_from_flat(CanvasState &s, size_t i, float rw, float gw, float bw, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_BLACK_AND_WHITE, {rw, gw, bw})); }
inline void vibrance(CanvasState &s, size_t i, float amount) { detail2::in_place(s, i, detail2::desc(PFE_ADJ_VIBRANCE, {amount / 100.0f})); }
inline void vibrance_from_flat(CanvasState &s, size_t i, float amount, const RgbaImage &flat) { detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_VIBRANCE, {amount / 100.0f})); }
// curves_from_flat_multi, :563: channel_points = [RGB, R, G, B, A], each (points, enabled)
using CurvePoints = std::vector<std::pair<float, float>>;
inline void curves_from_flat_multi(CanvasState &s, size_t i, const std::array<std::pair<CurvePoints, bool>, 5> &ch, const RgbaImage &flat) {
    uint8_t in[5 * 256], out[4 * 256];
    for (int c = 0; c < 5; c++) {
        if (ch[c].second) pfe_build_curves_lut(reinterpret_cast<const float *>(ch[c].first.data()), (int)ch[c].first.size(), in + c * 256);
        else for (int k = 0; k < 256; k++) in[c * 256 + k] = (uint8_t)k;
    }
    pfe_compose_curve_luts(in, out);
    detail2::from_flat(s, i, flat, detail2::desc(PFE_ADJ_LUT_RGBA, {}, out));
}
}  // namespace adjustments

namespace transform {
// src/ops/transform.rs:1015-1200
struct DisplacementField {
    uint32_t width, height;
    std::vector<float> data;  // (dx, dy) pairs
    DisplacementField(uint32_t w, uint32_t h) : width(w), height(h), data((size_t)w * h * 2, 0.0f) {}
    std::pair<float, float> get(uint32_t x, uint32_t y) const { size_t i = ((size_t)y * width + x) * 2; return {data[i], data[i + 1]}; }
    void add(uint32_t x, uint32_t y, float dx, float dy) { size_t i = ((size_t)y * width + x) * 2; data[i] += dx; data[i + 1] += dy; }
    using BBox = std::array<int32_t, 4>;
    BBox apply_push(float cx, float cy, float dx, float dy, float radius, float strength) { return brush(PFE_LIQ_PUSH, cx, cy, radius, strength, dx, dy); }
    BBox apply_expand(float cx, float cy, float radius, float strength) { return brush(PFE_LIQ_EXPAND, cx, cy, radius, strength, 0, 0); }
    BBox apply_contract(float cx, float cy, float radius, float strength) { return brush(PFE_LIQ_CONTRACT, cx, cy, radius, strength, 0, 0); }
    BBox apply_twirl(float cx, float cy, float radius, float strength, bool clockwise) { return brush(PFE_LIQ_TWIRL, cx, cy, radius, strength, clockwise ? 1.0f : 0.0f, 0); }

private:
    BBox brush(int kind, float cx, float cy, float r, float s, float a0, float a1) {
        BBox b{};
        Engine &e = Engine::current();
        e.check(pfe_liquify(e.ctx(), data.data(), width, height, kind, cx, cy, r, s, a0, a1, b.data()), "pfe_liquify");
        return b;
    }
};
using Points = std::vector<std::array<float, 2>>;
// :1288
inline RgbaImage warp_displacement_full(const RgbaImage &src, const DisplacementField &d) {
    RgbaImage out(d.width, d.height);
    Engine &e = Engine::current();
    e.check(pfe_warp_displacement(e.ctx(), src.as_raw().data(), src.width(), src.height(), d.data.data(), d.width, d.height, out.as_mut().data()), "pfe_warp_displacement");
    return out;
}
// :1670 and :1712
inline DisplacementField generate_displacement_from_mesh(const Points &orig, const Points &def, size_t cols, size_t rows, uint32_t w, uint32_t h) {
    DisplacementField f(w, h);
    Engine &e = Engine::current();
    e.check(pfe_mesh_displacement(e.ctx(), orig[0].data(), def[0].data(), (uint32_t)cols, (uint32_t)rows, w, h, f.data.data()), "pfe_mesh_displacement");
    return f;
}
inline void generate_displacement_from_mesh_fast(const Points &def, size_t cols, size_t rows, uint32_t w, uint32_t h, std::vector<float> &out) {
    out.resize((size_t)w * h * 2);
    Engine &e = Engine::current();
    e.check(pfe_mesh_displacement(e.ctx(), nullptr, def[0].data(), (uint32_t)cols, (uint32_t)rows, w, h, out.data()), "pfe_mesh_displacement");
}
// :1743 — fused on the device: the displacement field is never materialised
inline RgbaImage warp_mesh_catmull_rom(const RgbaImage &src, const Points &orig, const Points &def, size_t cols, size_t rows, uint32_t w, uint32_t h) {
    RgbaImage out(w, h);
    Engine &e = Engine::current();
    e.check(pfe_mesh_warp(e.ctx(), src.as_raw().data(), src.width(), src.height(), orig[0].data(), def[0].data(), (uint32_t)cols, (uint32_t)rows, w, h, out.as_mut().data()), "pfe_mesh_warp");
    return out;
}
// :467 flatten_image
inline void flatten_image(CanvasState &state) {
    RgbaImage comp = state.composite();
    state.layers.clear();
    state.layers.emplace_back("Background", state.width, state.height, Rgba{{0, 0, 0, 0}});
    state.layers[0].pixels = TiledImage::from_rgba_image(comp);
}

// ---- geometry: src/ops/transform.rs:14-131, :134-186, :347-463, :750-820 (selection-region variants
// `try_transform_selected_region` / `flip_layer_selected_region_*` are GUI paths and not mirrored)
enum class Interpolation { Nearest = 0, Bilinear = 1, Bicubic = 2, Lanczos3 = 3 };
namespace detail3 {
template <class F>
inline RgbaImage reshape(const RgbaImage &flat, uint32_t nw, uint32_t nh, const char *what, F f) {
    RgbaImage out(nw, nh);
    if (flat.width() == 0 || flat.height() == 0 || nw == 0 || nh == 0) return out;
    Engine &e = Engine::current();
    e.check(f(e.ctx(), flat.as_raw().data(), out.as_mut().data()), what);
    return out;
}
inline RgbaImage orient(const RgbaImage &flat, int op) {
    const bool turn = op == PFE_ORIENT_ROTATE_90CW || op == PFE_ORIENT_ROTATE_90CCW;
    return reshape(flat, turn ? flat.height() : flat.width(), turn ? flat.width() : flat.height(), "pfe_orient",
                   [&](pfe_ctx *c, const uint8_t *s, uint8_t *d) { return pfe_orient(c, s, flat.width(), flat.height(), op, d); });
}
inline void orient_all(CanvasState &s, int op) {
    for (auto &l : s.layers) l.pixels = TiledImage::from_rgba_image(orient(l.pixels.to_rgba_image(), op));
    if (op == PFE_ORIENT_ROTATE_90CW || op == PFE_ORIENT_ROTATE_90CCW) std::swap(s.width, s.height);
}
inline RgbaImage apply_affine(const RgbaImage &src, uint32_t cw, uint32_t ch, float rz, float rx, float ry, float scale,
                              std::pair<float, float> offset, Interpolation interp) {
    return reshape(src, cw, ch, "pfe_affine", [&](pfe_ctx *c, const uint8_t *s, uint8_t *d) {
        return pfe_affine(c, s, src.width(), src.height(), cw, ch, rz, rx, ry, scale, offset.first, offset.second, interp == Interpolation::Nearest ? 1 : 0, d);
    });
}
}  // namespace detail3
inline void flip_canvas_horizontal(CanvasState &s) { detail3::orient_all(s, PFE_ORIENT_FLIP_H); }
inline void flip_canvas_vertical(CanvasState &s) { detail3::orient_all(s, PFE_ORIENT_FLIP_V); }
inline void rotate_canvas_90cw(CanvasState &s) { detail3::orient_all(s, PFE_ORIENT_ROTATE_90CW); }
inline void rotate_canvas_90ccw(CanvasState &s) { detail3::orient_all(s, PFE_ORIENT_ROTATE_90CCW); }
inline void rotate_canvas_180(CanvasState &s) { detail3::orient_all(s, PFE_ORIENT_ROTATE_180); }
inline void flip_layer_horizontal(CanvasState &s, size_t i) { if (i < s.layers.size()) s.layers[i].pixels = TiledImage::from_rgba_image(detail3::orient(s.layers[i].pixels.to_rgba_image(), PFE_ORIENT_FLIP_H)); }
inline void flip_layer_vertical(CanvasState &s, size_t i) { if (i < s.layers.size()) s.layers[i].pixels = TiledImage::from_rgba_image(detail3::orient(s.layers[i].pixels.to_rgba_image(), PFE_ORIENT_FLIP_V)); }
inline void resize_image(CanvasState &s, uint32_t nw, uint32_t nh, Interpolation interp) {
    for (auto &l : s.layers) {
        RgbaImage flat = l.pixels.to_rgba_image();
        l.pixels = TiledImage::from_rgba_image(detail3::reshape(flat, nw, nh, "pfe_resize", [&](pfe_ctx *c, const uint8_t *src, uint8_t *d) {
            return pfe_resize(c, src, flat.width(), flat.height(), nw, nh, (int)interp, d);
        }));
    }
    s.width = nw;
    s.height = nh;
}
inline void resize_canvas(CanvasState &s, uint32_t nw, uint32_t nh, std::pair<uint32_t, uint32_t> anchor, Rgba fill) {
    for (auto &l : s.layers) {
        RgbaImage flat = l.pixels.to_rgba_image();
        l.pixels = TiledImage::from_rgba_image(detail3::reshape(flat, nw, nh, "pfe_resize_canvas", [&](pfe_ctx *c, const uint8_t *src, uint8_t *d) {
            return pfe_resize_canvas(c, src, flat.width(), flat.height(), nw, nh, anchor.first, anchor.second, fill.v, d);
        }));
    }
    s.width = nw;
    s.height = nh;
}
inline void affine_transform_layer(CanvasState &s, size_t i, float rz, float rx, float ry, float scale, std::pair<float, float> offset) {
    if (i >= s.layers.size()) return;
    s.layers[i].pixels = TiledImage::from_rgba_image(detail3::apply_affine(s.layers[i].pixels.to_rgba_image(), s.width, s.height, rz, rx, ry, scale, offset, Interpolation::Bilinear));
}
inline void affine_transform_layer_from_flat(CanvasState &s, size_t i, float rz, float rx, float ry, float scale, std::pair<float, float> offset, const RgbaImage &flat) {
    if (i >= s.layers.size()) return;
    s.layers[i].pixels = TiledImage::from_rgba_image(detail3::apply_affine(flat, s.width, s.height, rz, rx, ry, scale, offset, Interpolation::Bilinear));
}
inline void rotate_canvas_arbitrary(CanvasState &s, float degrees, Interpolation interp) {
    if (std::fabs(degrees) < 0.001f) return;
    for (auto &l : s.layers) {
        l.pixels = TiledImage::from_rgba_image(detail3::apply_affine(l.pixels.to_rgba_image(), s.width, s.height, degrees, 0, 0, 1.0f, {0, 0}, interp));
        if (l.mask) l.mask = TiledImage::from_rgba_image(detail3::apply_affine(l.mask->to_rgba_image(), s.width, s.height, degrees, 0, 0, 1.0f, {0, 0}, interp));
    }
}
}  // namespace transform
}  // namespace ops

// ================================================================================================
namespace gpu {
// src/gpu/renderer.rs:915-947 — same shapes; median's Option::None maps to PFE_ERR_UNSUPPORTED.
class GpuRenderer {
public:
    static std::optional<GpuRenderer> try_new(const std::string & /*preferred*/) {
        try { Engine::current(); } catch (const Error &) { return std::nullopt; }
        return GpuRenderer();
    }
    std::vector<uint8_t> blur_rgba(const std::vector<uint8_t> &data, uint32_t w, uint32_t h, float sigma) const {
        std::vector<uint8_t> out(data.size());
        Engine &e = Engine::current();
        e.check(pfe_gaussian_blur(e.ctx(), data.data(), w, h, sigma, nullptr, out.data(), 0), "blur_rgba");
        return out;
    }
    std::vector<uint8_t> brightness_contrast_rgba(const std::vector<uint8_t> &data, uint32_t w, uint32_t h, float b, float c) const {
        return adjust(data, w, h, ops::adjustments::detail2::desc(PFE_ADJ_BRIGHTNESS_CONTRAST, {b, c}));
    }
    std::vector<uint8_t> hsl_rgba(const std::vector<uint8_t> &data, uint32_t w, uint32_t h, float hue, float sat, float light) const {
        return adjust(data, w, h, ops::adjustments::detail2::desc(PFE_ADJ_HSL, {hue, sat, light}));
    }
    std::vector<uint8_t> invert_rgba(const std::vector<uint8_t> &data, uint32_t w, uint32_t h) const {
        return adjust(data, w, h, ops::adjustments::detail2::desc(PFE_ADJ_INVERT));
    }
    std::optional<std::vector<uint8_t>> median_rgba(const std::vector<uint8_t> &data, uint32_t w, uint32_t h, uint32_t radius) const {
        std::vector<uint8_t> out(data.size());
        Engine &e = Engine::current();
        int rc = pfe_median(e.ctx(), data.data(), w, h, radius, nullptr, out.data());
        if (rc == PFE_ERR_UNSUPPORTED) return std::nullopt;
        e.check(rc, "median_rgba");
        return out;
    }
    // composite(canvas_w, canvas_h, layer_info) with layers supplied directly (renderer.rs:533)
    std::optional<std::vector<uint8_t>> composite(uint32_t w, uint32_t h, const std::vector<pfe_layer_desc> &layers) const {
        std::vector<uint8_t> out((size_t)w * h * 4);
        Engine &e = Engine::current();
        e.check(pfe_flatten(e.ctx(), layers.data(), (uint32_t)layers.size(), w, h, nullptr, out.data()), "composite");
        return out;
    }

private:
    static std::vector<uint8_t> adjust(const std::vector<uint8_t> &data, uint32_t w, uint32_t h, const pfe_adjust_desc &d) {
        std::vector<uint8_t> out(data.size());
        Engine &e = Engine::current();
        e.check(pfe_adjust(e.ctx(), data.data(), w, h, &d, nullptr, nullptr, out.data()), "pfe_adjust");
        return out;
    }
};
}  // namespace gpu

}  // namespace paintfe
