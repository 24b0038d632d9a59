// Ports of the reference's own integration tests onto the C++ mirror (include/paintfe/paintfe.hpp),
// so the drop-in claim is checked at the API the reference's callers use.
//   tests/visual_blend.rs, visual_filters.rs, visual_adjustments.rs, transform_ops.rs, gpu_pipelines.rs
// usage: mirror_tests <dir with golden .rgba files: <category>__<name>.rgba = w(u32) h(u32) pixels>
#include <cstdio>
#include <fstream>
#include <functional>
#include <string>

#include "../../include/paintfe/paintfe.hpp"

using namespace paintfe;
using namespace paintfe::canvas;

static std::string g_dir;
static int g_fail = 0, g_run = 0;

static RgbaImage load_golden(const std::string &cat, const std::string &name) {
    std::ifstream f(g_dir + "/" + cat + "__" + name + ".rgba", std::ios::binary);
    if (!f) throw std::runtime_error("missing golden " + cat + "/" + name);
    uint32_t wh[2];
    f.read((char *)wh, 8);
    std::vector<uint8_t> d((size_t)wh[0] * wh[1] * 4);
    f.read((char *)d.data(), (std::streamsize)d.size());
    return *RgbaImage::from_raw(wh[0], wh[1], std::move(d));
}
static int max_diff(const RgbaImage &a, const RgbaImage &b) {
    if (a.width() != b.width() || a.height() != b.height()) return 999;
    int m = 0;
    for (size_t i = 0; i < a.as_raw().size(); i++) m = std::max(m, std::abs((int)a.as_raw()[i] - (int)b.as_raw()[i]));
    return m;
}
static void assert_golden(const std::string &cat, const std::string &name, const RgbaImage &img, int tol = 0) {  // common/mod.rs:211
    int d = max_diff(img, load_golden(cat, name));
    if (d > tol) throw std::runtime_error(cat + "/" + name + ": max channel diff " + std::to_string(d));
}
static void check(bool ok, const char *what) { if (!ok) throw std::runtime_error(what); }
static void run(const char *name, const std::function<void()> &f) {
    g_run++;
    try { f(); } catch (const std::exception &e) { g_fail++; std::printf("FAILED %s: %s\n", name, e.what()); }
}

// tests/common/mod.rs:272-316
static RgbaImage create_test_gradient(uint32_t w, uint32_t h) {
    RgbaImage img(w, h);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            uint8_t r = w > 1 ? (uint8_t)(x * 255 / (w - 1)) : 128, b = h > 1 ? (uint8_t)(y * 255 / (h - 1)) : 128;
            img.put_pixel(x, y, Rgba{{r, (uint8_t)(255 - r), b, 255}});
        }
    return img;
}
static RgbaImage create_solid(uint32_t w, uint32_t h, Rgba c) {
    RgbaImage img(w, h);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) img.put_pixel(x, y, c);
    return img;
}
static RgbaImage create_test_checkerboard(uint32_t w, uint32_t h) {
    RgbaImage img(w, h);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            uint8_t v = ((x / 8 + y / 8) % 2 == 0) ? 255 : 0;
            img.put_pixel(x, y, Rgba{{v, v, v, 255}});
        }
    return img;
}
static CanvasState canvas_from_image(const RgbaImage &img) {
    CanvasState s(img.width(), img.height());
    s.layers[0].pixels = TiledImage::from_rgba_image(img);
    return s;
}
static RgbaImage gradient_32() {  // transform_ops.rs:25
    RgbaImage img(32, 32);
    for (uint32_t y = 0; y < 32; y++)
        for (uint32_t x = 0; x < 32; x++) img.put_pixel(x, y, Rgba{{(uint8_t)(x * 8), (uint8_t)(y * 8), 128, 255}});
    return img;
}
static ops::transform::Points uniform_grid(size_t cols, size_t rows, float w, float h) {
    ops::transform::Points p;
    for (size_t r = 0; r <= rows; r++)
        for (size_t c = 0; c <= cols; c++) p.push_back({(float)c / (float)cols * w, (float)r / (float)rows * h});
    return p;
}

// visual_blend.rs:19-47
static RgbaImage make_blend_test(BlendMode mode) {
    const uint32_t w = 64, h = 64;
    RgbaImage fg(w, h);
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            uint8_t r = (uint8_t)(((float)x / (float)w) * 255.0f), g = (uint8_t)(((float)y / (float)h) * 255.0f);
            uint8_t a = (uint8_t)(((float)(x + y) / (float)(w + h - 2)) * 200.0f + 55.0f);
            fg.put_pixel(x, y, Rgba{{r, g, 128, a}});
        }
    CanvasState state(w, h);
    state.layers[0].pixels = TiledImage::from_rgba_image(create_test_checkerboard(w, h));
    Layer top("Foreground", w, h, Rgba{{0, 0, 0, 0}});
    top.blend_mode = mode;
    top.pixels = TiledImage::from_rgba_image(fg);
    state.layers.push_back(std::move(top));
    return state.composite();
}

int main(int argc, char **argv) {
    if (argc < 2) { std::printf("usage: mirror_tests <golden_raw_dir>\n"); return 2; }
    g_dir = argv[1];
    const char *blend_names[25] = {"normal", "multiply", "screen", "additive", "reflect", "glow", "color_burn", "color_dodge", "overlay",
                                   "difference", "negation", "lighten", "darken", "xor", "overwrite", "hard_light", "soft_light",
                                   "exclusion", "subtract", "divide", "linear_burn", "vivid_light", "linear_light", "pin_light", "hard_mix"};
    for (int m = 0; m < 25; m++)
        run(blend_names[m], [&] { assert_golden("blend", blend_names[m], make_blend_test((BlendMode)m)); });
    run("normal_half_opacity", [] {
        CanvasState state(64, 64);
        state.layers[0].pixels = TiledImage::from_rgba_image(create_test_checkerboard(64, 64));
        Layer fg("Foreground", 64, 64, Rgba{{0, 0, 0, 0}});
        fg.opacity = 0.5f;
        fg.pixels = TiledImage::from_rgba_image(create_test_gradient(64, 64));
        state.layers.push_back(std::move(fg));
        assert_golden("blend", "normal_half_opacity", state.composite());
    });
    run("hidden_layer_is_skipped", [] {  // visual_blend.rs: hidden layer test
        CanvasState state(64, 64);
        state.layers[0].pixels = TiledImage::from_rgba_image(create_test_checkerboard(64, 64));
        Layer fg("Hidden", 64, 64, Rgba{{255, 0, 0, 255}});
        fg.visible = false;
        state.layers.push_back(std::move(fg));
        check(state.composite() == create_test_checkerboard(64, 64), "hidden layer changed the composite");
    });
    run("layer_mask_and_adjustment_layer", [] {
        CanvasState state(64, 64);
        Layer fg("Masked", 64, 64, Rgba{{0, 0, 255, 255}});
        TiledImage mask(64, 64);
        for (uint32_t y = 0; y < 64; y++) for (uint32_t x = 0; x < 32; x++) mask.put_pixel(x, y, Rgba{{0, 0, 0, 255}});
        fg.mask = mask;
        fg.mask_enabled = true;
        state.layers.push_back(std::move(fg));
        Layer adj("Invert", 64, 64, Rgba{{0, 0, 0, 0}});
        adj.adjustment = AdjustmentLayerData{};
        state.layers.push_back(std::move(adj));
        RgbaImage out = state.composite();
        check(out == state.composite_dense(), "tile-native and dense composites differ");
        check(out.get_pixel(5, 5) == Rgba{{0, 0, 0, 255}}, "concealed half: white background inverted");
        check(out.get_pixel(40, 5) == Rgba{{255, 255, 0, 255}}, "revealed half: blue inverted");
    });

    // layer_ops.rs:151-316
    auto solid_layer = [](const char *name, Rgba c) {
        Layer l(name, 32, 32, Rgba{{0, 0, 0, 0}});
        l.pixels = TiledImage::from_rgba_image(RgbaImage::from_pixel(32, 32, c));
        return l;
    };
    run("hidden_layer_not_composited", [&] {
        CanvasState state(32, 32);
        Layer red = solid_layer("Red", Rgba{{255, 0, 0, 255}});
        red.visible = false;
        state.layers.push_back(std::move(red));
        check(state.composite().get_pixel(16, 16) == Rgba{{255, 255, 255, 255}}, "hidden layer showed");
    });
    run("hidden_folder_hides_member_layers", [&] {
        CanvasState state(32, 32);
        state.layer_folders.push_back(canvas::LayerFolder{1, "Group", false, false});
        Layer red = solid_layer("Red", Rgba{{255, 0, 0, 255}});
        red.folder_id = 1;
        state.layers.push_back(std::move(red));
        check(state.composite().get_pixel(16, 16) == Rgba{{255, 255, 255, 255}}, "layer of a hidden folder showed");
        check(state.composite_dense().get_pixel(16, 16) == Rgba{{255, 255, 255, 255}}, "dense: layer of a hidden folder showed");
        check(!state.layer_effectively_visible(1), "layer_effectively_visible");
    });
    run("layer_opacity_affects_composite", [&] {
        CanvasState state(32, 32);
        Layer black = solid_layer("Black50", Rgba{{0, 0, 0, 255}});
        black.opacity = 0.5f;
        state.layers.push_back(std::move(black));
        check(std::abs((int)state.composite().get_pixel(16, 16)[0] - 128) <= 2, "expected ~128 gray");
    });
    run("layer_reorder_changes_composite", [&] {
        CanvasState state(32, 32);
        state.layers.push_back(solid_layer("Red", Rgba{{255, 0, 0, 255}}));
        state.layers.push_back(solid_layer("Blue", Rgba{{0, 0, 255, 255}}));
        check(state.composite().get_pixel(16, 16)[2] == 255, "blue on top");
        std::swap(state.layers[1], state.layers[2]);
        check(state.composite().get_pixel(16, 16)[0] == 255, "red on top after swap");
    });
    run("flatten_multiple_layers", [&] {
        CanvasState state(32, 32);
        state.layers.push_back(solid_layer("Red", Rgba{{255, 0, 0, 128}}));
        RgbaImage before = state.composite();
        ops::transform::flatten_image(state);
        check(state.layers.size() == 1, "flatten should produce one layer");
        check(state.composite() == before, "composite should be unchanged after flatten");
    });
    run("flatten_preserves_hidden_layer_exclusion", [&] {
        CanvasState state(32, 32);
        Layer green = solid_layer("Green", Rgba{{0, 255, 0, 255}});
        green.visible = false;
        state.layers.push_back(std::move(green));
        RgbaImage before = state.composite();
        ops::transform::flatten_image(state);
        check(state.layers.size() == 1 && state.composite().get_pixel(16, 16) == Rgba{{255, 255, 255, 255}} && state.composite() == before, "hidden layer leaked into flatten");
    });

    // visual_filters.rs
    const RgbaImage img = create_test_gradient(64, 64);
    run("gaussian_blur_s2", [&] { assert_golden("filters", "gaussian_blur_s2", ops::filters::parallel_gaussian_blur_pub(img, 2.0f)); });
    run("gaussian_blur_s5", [&] { assert_golden("filters", "gaussian_blur_s5", ops::filters::parallel_gaussian_blur_pub(img, 5.0f)); });
    run("gaussian_blur_s5_fast", [&] { assert_golden("filters", "gaussian_blur_s5", ops::filters::parallel_gaussian_blur_pub(img, 5.0f, false), 1); });
    run("motion_blur_45_10", [&] { assert_golden("filters", "motion_blur_45_10", ops::effects::motion_blur_core(img, 45.0f, 10.0f, nullptr)); });
    run("box_blur_r3", [&] { assert_golden("filters", "box_blur_r3", ops::effects::box_blur_core(img, 3.0f, nullptr)); });
    run("median_r2", [&] { assert_golden("filters", "median_r2", ops::effects::median_core(img, 2, nullptr)); });
    run("sharpen_a1_r1", [&] { assert_golden("filters", "sharpen_a1_r1", ops::effects::sharpen_core(img, 1.0f, 1.0f, nullptr)); });
    run("vignette_08_05", [&] { assert_golden("filters", "vignette_08_05", ops::effects::vignette_core(img, 0.8f, 0.5f, nullptr)); });
    run("glow_r3_i05", [&] { assert_golden("filters", "glow_r3_i05", ops::effects::glow_core(img, 3.0f, 0.5f, nullptr)); });
    run("pixelate_8", [&] { assert_golden("filters", "pixelate_8", ops::effects::pixelate_core(img, 8, nullptr)); });
    run("bulge_05", [&] { assert_golden("filters", "bulge_05", ops::effects::bulge_core(img, 0.5f, nullptr)); });
    run("twist_45", [&] { assert_golden("filters", "twist_45", ops::effects::twist_core(img, 45.0f, nullptr), 1); });
    run("add_noise_uniform", [&] { assert_golden("filters", "add_noise_uniform", ops::effects::add_noise_core(img, 30.0f, ops::effects::NoiseType::Uniform, false, 42, 1.0f, 1, nullptr)); });
    run("add_noise_gaussian", [&] { assert_golden("filters", "add_noise_gaussian_mono", ops::effects::add_noise_core(img, 30.0f, ops::effects::NoiseType::Gaussian, true, 42, 1.0f, 1, nullptr), 1); });
    run("add_noise_perlin", [&] { assert_golden("filters", "add_noise_perlin", ops::effects::add_noise_core(img, 50.0f, ops::effects::NoiseType::Perlin, false, 42, 5.0f, 3, nullptr)); });
    run("reduce_noise", [&] { assert_golden("filters", "reduce_noise", ops::effects::reduce_noise_core(img, 0.5f, 2, nullptr), 1); });
    // the rest of visual_filters.rs (:44-283)
    using namespace ops::effects;
    run("bokeh_blur_r5", [&] { assert_golden("filters", "bokeh_blur_r5", bokeh_blur_core(img, 5.0f, nullptr)); });
    run("zoom_blur", [&] { assert_golden("filters", "zoom_blur", zoom_blur_core(img, 0.5f, 0.5f, 0.3f, 8, {0.0f, 0.0f, 0.0f, 0.0f}, 0.0f, nullptr)); });
    run("crystallize_s16", [&] { assert_golden("filters", "crystallize_s16", crystallize_core(img, 16.0f, 42, nullptr)); });
    run("dents", [&] { assert_golden("filters", "dents", dents_core(img, 20.0f, 10.0f, 42, 2, 0.5f, false, false, nullptr)); });
    run("halftone_circle", [&] { assert_golden("filters", "halftone_circle", halftone_core(img, 4.0f, 45.0f, HalftoneShape::Circle, nullptr)); });
    run("grid_lines_16", [&] { assert_golden("filters", "grid_lines_16", grid_core(img, 16, 16, 1, {0, 0, 0, 255}, GridStyle::Lines, 1.0f, nullptr)); });
    run("drop_shadow", [&] {
        RgbaImage sq = create_solid(64, 64, Rgba{{0, 0, 0, 0}});
        for (uint32_t y = 16; y < 48; y++) for (uint32_t x = 16; x < 48; x++) sq.put_pixel(x, y, Rgba{{255, 255, 255, 255}});
        assert_golden("filters", "drop_shadow", shadow_core(sq, 5, 5, 3.0f, false, {0, 0, 0, 255}, 0.8f, nullptr));
    });
    run("outline_outside", [&] {
        RgbaImage sq = create_solid(64, 64, Rgba{{0, 0, 0, 0}});
        for (uint32_t y = 16; y < 48; y++) for (uint32_t x = 16; x < 48; x++) sq.put_pixel(x, y, Rgba{{255, 0, 0, 255}});
        assert_golden("filters", "outline_outside", outline_core(sq, 2, {0, 0, 255, 255}, OutlineMode::Outside, true, nullptr));
    });
    run("contours", [&] { assert_golden("filters", "contours", contours_core(img, 10.0f, 5.0f, 1.0f, {0, 0, 0, 255}, 42, 2, 0.5f, nullptr)); });
    run("canvas_border_core_applies_edges_only", [&] {
        RgbaImage out = canvas_border_core(create_solid(8, 8, Rgba{{10, 20, 30, 255}}), 2, {200, 100, 50, 255}, nullptr);
        check(out.get_pixel(0, 0) == Rgba{{200, 100, 50, 255}}, "edge pixel should be the border colour");
        check(out.get_pixel(3, 3) == Rgba{{10, 20, 30, 255}}, "interior pixel should be unchanged");
    });
    run("pixel_drag", [&] { assert_golden("filters", "pixel_drag", pixel_drag_core(img, 42, 50.0f, 20, 0.0f, nullptr)); });
    run("rgb_displace", [&] { assert_golden("filters", "rgb_displace", rgb_displace_core(img, {5, 0}, {0, 0}, {-5, 0}, nullptr)); });
    run("ink", [&] { assert_golden("filters", "ink", ink_core(img, 1.0f, 0.5f, nullptr)); });
    run("oil_painting", [&] { assert_golden("filters", "oil_painting", oil_painting_core(img, 3, 20, nullptr)); });
    run("color_filter_multiply", [&] { assert_golden("filters", "color_filter_multiply", color_filter_core(img, {255, 128, 0, 255}, 0.5f, ColorFilterMode::Multiply, nullptr)); });
    run("color_filter_identity", [&] { check(color_filter_core(img, {255, 255, 255, 255}, 0.0f, ColorFilterMode::Multiply, nullptr) == img, "intensity 0 must be identity"); });
    run("gaussian_sigma0_identity", [&] { check(ops::filters::parallel_gaussian_blur_pub(img, 0.0f) == img, "sigma 0 must be identity"); });  // visual_filters.rs:296
    run("sharpen_amount0_identity", [&] { check(ops::effects::sharpen_core(img, 0.0f, 1.0f, nullptr) == img, "amount 0 must be identity"); });
    run("selection_mask_limits_blur", [&] {
        GrayImage m(64, 64);
        for (uint32_t y = 0; y < 64; y++) for (uint32_t x = 0; x < 32; x++) m.put_pixel(x, y, 255);
        RgbaImage out = ops::filters::blur_with_selection_pub(create_test_checkerboard(64, 64), 3.0f, &m);
        check(out.get_pixel(50, 20) == create_test_checkerboard(64, 64).get_pixel(50, 20), "unselected pixel changed");
        check(!(out.get_pixel(8, 20) == create_test_checkerboard(64, 64).get_pixel(8, 20)), "selected edge pixel unchanged");
    });

    // visual_adjustments.rs
    auto extract = [](const CanvasState &s) { return s.layers[0].pixels.to_rgba_image(); };
    run("invert_colors", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::invert_colors(s, 0); assert_golden("adjustments", "invert_colors", extract(s)); });
    run("invert_colors_roundtrip", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::invert_colors(s, 0); ops::adjustments::invert_colors(s, 0); check(extract(s) == img, "invert twice"); });
    run("invert_alpha", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::invert_alpha(s, 0); assert_golden("adjustments", "invert_alpha", extract(s)); });
    run("sepia", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::sepia(s, 0); assert_golden("adjustments", "sepia", extract(s)); });
    run("brightness_contrast", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::brightness_contrast_from_flat(s, 0, 30.0f, 20.0f, img); assert_golden("adjustments", "brightness_30_contrast_20", extract(s)); });
    run("hsl_adjustment", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::hue_saturation_lightness_from_flat(s, 0, 30.0f, -20.0f, 10.0f, img); assert_golden("adjustments", "hsl_h30_s-20_l10", extract(s)); });
    run("exposure", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::exposure_from_flat(s, 0, 1.0f, img); assert_golden("adjustments", "exposure_1ev", extract(s)); });
    run("levels", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::levels_from_flat(s, 0, 20.0f, 235.0f, 1.2f, 0.0f, 255.0f, img); assert_golden("adjustments", "levels", extract(s)); });
    run("highlights_shadows", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::highlights_shadows_from_flat(s, 0, 30.0f, -20.0f, img); assert_golden("adjustments", "highlights_shadows", extract(s)); });
    run("temperature_tint", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::temperature_tint_from_flat(s, 0, 30.0f, 10.0f, img); assert_golden("adjustments", "temperature_tint", extract(s)); });
    run("threshold", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::threshold_from_flat(s, 0, 128.0f, img); assert_golden("adjustments", "threshold_128", extract(s)); });
    run("posterize", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::posterize_from_flat(s, 0, 4, img); assert_golden("adjustments", "posterize_4", extract(s)); });
    run("color_balance", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::color_balance_from_flat(s, 0, {10.0f, 0.0f, -10.0f}, {0.0f, 0.0f, 0.0f}, {-10.0f, 0.0f, 10.0f}, img); assert_golden("adjustments", "color_balance", extract(s)); });
    run("color_balance_identity", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::color_balance_from_flat(s, 0, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, {0.0f, 0.0f, 0.0f}, img); check(extract(s) == img, "all-zero colour balance must be identity"); });
    run("gradient_map", [&] {  // visual_adjustments.rs:299-316
        ops::adjustments::GradientLut lut;
        for (int i = 0; i < 256; i++) {
            float t = (float)i / 255.0f;
            lut[i] = {(uint8_t)(t * 255.0f), (uint8_t)(t * t * 200.0f), (uint8_t)(t * t * t * 150.0f), 255};
        }
        CanvasState s = canvas_from_image(img);
        ops::adjustments::gradient_map_from_flat(s, 0, lut, img);
        assert_golden("adjustments", "gradient_map", extract(s));
    });
    run("black_and_white", [&] {
        RgbaImage bands(64, 64);
        const Rgba colors[8] = {{{255, 0, 0, 255}}, {{0, 255, 0, 255}}, {{0, 0, 255, 255}}, {{0, 255, 255, 255}}, {{255, 0, 255, 255}}, {{255, 255, 0, 255}}, {{255, 255, 255, 255}}, {{0, 0, 0, 255}}};
        for (uint32_t y = 0; y < 64; y++) for (uint32_t x = 0; x < 64; x++) bands.put_pixel(x, y, colors[std::min<uint32_t>(x * 8 / 64, 7)]);
        CanvasState s = canvas_from_image(bands);
        ops::adjustments::black_and_white_from_flat(s, 0, 0.3f, 0.59f, 0.11f, bands);
        assert_golden("adjustments", "black_and_white", extract(s));
    });
    run("vibrance", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::vibrance_from_flat(s, 0, 50.0f, img); assert_golden("adjustments", "vibrance_50", extract(s)); });
    run("vibrance_identity", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::vibrance_from_flat(s, 0, 0.0f, img); check(extract(s) == img, "vibrance 0 must be identity"); });
    run("curves_identity", [&] {  // visual_adjustments.rs:228-246
        CanvasState s = canvas_from_image(img);
        std::array<std::pair<ops::adjustments::CurvePoints, bool>, 5> ch;
        for (auto &c : ch) c = {{{0.0f, 0.0f}, {255.0f, 255.0f}}, false};
        ops::adjustments::curves_from_flat_multi(s, 0, ch, img);
        check(extract(s) == img, "disabled curves must be identity");
    });
    run("bad_layer_index_is_noop", [&] { CanvasState s = canvas_from_image(img); ops::adjustments::invert_colors(s, 7); ops::filters::gaussian_blur_layer(s, 7, 2.0f); check(extract(s) == img, "out-of-range layer index"); });


    // visual_transforms.rs
    {
        using namespace ops::transform;
        const RgbaImage ti = create_test_gradient(64, 48);
        auto ex0 = [](const CanvasState &s) { return s.layers[0].pixels.to_rgba_image(); };
        run("flip_canvas_h", [&] { CanvasState s = canvas_from_image(ti); flip_canvas_horizontal(s); assert_golden("transforms", "flip_canvas_h", ex0(s)); });
        run("flip_canvas_v", [&] { CanvasState s = canvas_from_image(ti); flip_canvas_vertical(s); assert_golden("transforms", "flip_canvas_v", ex0(s)); });
        run("flip_canvas_h_roundtrip", [&] { CanvasState s = canvas_from_image(ti); flip_canvas_horizontal(s); flip_canvas_horizontal(s); check(ex0(s) == ti, "flip h x 2 should be identity"); });
        run("rotate_90cw", [&] {
            CanvasState s = canvas_from_image(ti);
            rotate_canvas_90cw(s);
            check(ex0(s).width() == 48 && ex0(s).height() == 64 && s.width == 48 && s.height == 64, "90cw swaps the dimensions");
            assert_golden("transforms", "rotate_90cw", ex0(s));
        });
        run("rotate_90ccw", [&] { CanvasState s = canvas_from_image(ti); rotate_canvas_90ccw(s); assert_golden("transforms", "rotate_90ccw", ex0(s)); });
        run("rotate_180", [&] { CanvasState s = canvas_from_image(ti); rotate_canvas_180(s); assert_golden("transforms", "rotate_180", ex0(s)); });
        run("rotate_90cw_x4_identity", [&] { CanvasState s = canvas_from_image(ti); for (int i = 0; i < 4; i++) rotate_canvas_90cw(s); check(ex0(s) == ti, "4 x 90cw should be identity"); });
        run("rotate_90cw_then_ccw_identity", [&] { CanvasState s = canvas_from_image(ti); rotate_canvas_90cw(s); rotate_canvas_90ccw(s); check(ex0(s) == ti, "90cw + 90ccw should be identity"); });
        run("resize_2x_nearest", [&] { CanvasState s = canvas_from_image(ti); resize_image(s, 128, 96, Interpolation::Nearest); assert_golden("transforms", "resize_2x_nearest", ex0(s)); });
        run("resize_half_bilinear", [&] { CanvasState s = canvas_from_image(ti); resize_image(s, 32, 24, Interpolation::Bilinear); assert_golden("transforms", "resize_half_bilinear", ex0(s)); });
        run("resize_half_lanczos", [&] { CanvasState s = canvas_from_image(ti); resize_image(s, 32, 24, Interpolation::Lanczos3); assert_golden("transforms", "resize_half_lanczos", ex0(s)); });
        run("resize_canvas_center", [&] { CanvasState s = canvas_from_image(ti); resize_canvas(s, 96, 80, {1, 1}, Rgba{{0, 0, 0, 0}}); assert_golden("transforms", "resize_canvas_center", ex0(s)); });
        run("resize_canvas_topleft", [&] { CanvasState s = canvas_from_image(ti); resize_canvas(s, 80, 64, {0, 0}, Rgba{{255, 0, 0, 255}}); assert_golden("transforms", "resize_canvas_topleft", ex0(s)); });
        run("flip_layer_h", [&] { CanvasState s = canvas_from_image(ti); flip_layer_horizontal(s, 0); assert_golden("transforms", "flip_layer_h", ex0(s)); });
        run("flip_layer_v", [&] { CanvasState s = canvas_from_image(ti); flip_layer_vertical(s, 0); assert_golden("transforms", "flip_layer_v", ex0(s)); });
        run("affine_rotate_45", [&] { CanvasState s = canvas_from_image(ti); affine_transform_layer(s, 0, 45.0f * (3.14159265358979323846f / 180.0f), 0.0f, 0.0f, 1.0f, {0.0f, 0.0f}); assert_golden("transforms", "affine_rotate_45", ex0(s)); });
        run("affine_identity", [&] { CanvasState s = canvas_from_image(ti); affine_transform_layer(s, 0, 0.0f, 0.0f, 0.0f, 1.0f, {0.0f, 0.0f}); check(max_diff(ex0(s), ti) <= 1, "identity affine within 1 level"); });
        // transform_ops.rs:280-303
        const RgbaImage g32 = create_test_gradient(32, 32);
        run("affine_rotate_90_golden", [&] { CanvasState s = canvas_from_image(g32); affine_transform_layer(s, 0, 1.57079632679489661923f, 0.0f, 0.0f, 1.0f, {0.0f, 0.0f}); assert_golden("transform", "affine_rotate_90", s.composite()); });
        run("affine_scale_half_golden", [&] { CanvasState s = canvas_from_image(g32); affine_transform_layer(s, 0, 0.0f, 0.0f, 0.0f, 0.5f, {0.0f, 0.0f}); assert_golden("transform", "affine_scale_half", s.composite()); });
    }
    // transform_ops.rs
    using namespace ops::transform;
    run("displacement_identity_preserves_image", [] { check(warp_displacement_full(gradient_32(), DisplacementField(32, 32)) == gradient_32(), "identity warp"); });
    run("displacement_translate_shifts_pixels", [] {
        DisplacementField f(32, 32);
        for (uint32_t y = 0; y < 32; y++) for (uint32_t x = 0; x < 32; x++) f.add(x, y, 5.0f, 0.0f);
        check(warp_displacement_full(gradient_32(), f).get_pixel(10, 16) == gradient_32().get_pixel(5, 16), "shifted pixel mismatch");
    });
    run("displacement_field_golden", [] {
        DisplacementField f(32, 32);
        auto bb = f.apply_push(16.0f, 16.0f, 3.0f, 0.0f, 10.0f, 0.8f);
        check(bb == DisplacementField::BBox{6, 6, 26, 26}, "apply_push bbox");
        assert_golden("transform", "displacement_radial_push", warp_displacement_full(gradient_32(), f), 1);
    });
    run("warp_displacement_full_golden", [] {
        DisplacementField f(32, 32);
        for (uint32_t y = 0; y < 32; y++)
            for (uint32_t x = 0; x < 32; x++) {
                float dx = (float)x - 16.0f, dy = (float)y - 16.0f;
                float r = std::max(std::sqrt(dx * dx + dy * dy), 0.001f);
                float strength = std::max(1.0f - r / 16.0f, 0.0f);
                f.add(x, y, -dy * strength * 0.5f, dx * strength * 0.5f);
            }
        assert_golden("transform", "displacement_swirl", warp_displacement_full(gradient_32(), f));
    });
    run("mesh_warp_identity", [] {
        auto g = uniform_grid(2, 2, 32.0f, 32.0f);
        check(max_diff(warp_mesh_catmull_rom(gradient_32(), g, g, 2, 2, 32, 32), gradient_32()) <= 2, "identity mesh");
    });
    run("mesh_warp_deformed_golden", [] {
        auto o = uniform_grid(2, 2, 32.0f, 32.0f);
        auto d = o;
        d[4] = {20.0f, 20.0f};
        assert_golden("transform", "mesh_warp_deformed", warp_mesh_catmull_rom(gradient_32(), o, d, 2, 2, 32, 32));
        auto field = generate_displacement_from_mesh(o, d, 2, 2, 32, 32);
        assert_golden("transform", "mesh_warp_deformed", warp_displacement_full(gradient_32(), field));
    });
    run("flatten_image_collapses_layers", [&] {
        CanvasState s = canvas_from_image(create_test_checkerboard(64, 64));
        Layer fg("fg", 64, 64, Rgba{{0, 0, 0, 0}});
        fg.opacity = 0.5f;
        fg.pixels = TiledImage::from_rgba_image(img);
        s.layers.push_back(std::move(fg));
        RgbaImage before = s.composite();
        flatten_image(s);
        check(s.layers.size() == 1 && s.layers[0].pixels.to_rgba_image() == before, "flatten_image");
    });

    // gpu_pipelines.rs (same loose bounds as the reference's GPU tests)
    run("gpu_renderer_methods", [&] {
        auto r = gpu::GpuRenderer::try_new("");
        check(r.has_value(), "GpuRenderer::try_new");
        auto inv = r->invert_rgba(img.as_raw(), 64, 64);
        check(inv[0] == 255 - img.as_raw()[0] && inv[3] == 255, "invert_rgba");
        auto bl = r->blur_rgba(img.as_raw(), 64, 64, 2.0f);
        check(max_diff(*RgbaImage::from_raw(64, 64, bl), load_golden("filters", "gaussian_blur_s2")) <= 1, "blur_rgba");
        check(r->median_rgba(img.as_raw(), 64, 64, 2).has_value(), "median_rgba r=2");
        // the reference's GPU median declines radii above 7 (noise.rs:311-322) and the caller falls back to median_core,
        // which sorts any window; this library serves every radius up to 20000 itself and declines only beyond
        {
            auto big = r->median_rgba(img.as_raw(), 64, 64, 200);
            check(big.has_value(), "median_rgba r=200");
            // a window that covers the whole image from every pixel... is still a per-pixel rank over clamped samples
            check(big->size() == img.as_raw().size(), "median_rgba r=200 size");
        }
        check(!r->median_rgba(img.as_raw(), 64, 64, 30000).has_value(), "median_rgba r=30000 -> None");
        check(r->hsl_rgba(img.as_raw(), 64, 64, 0.0f, 0.0f, 0.0f) == img.as_raw(), "hsl identity");
        check(r->brightness_contrast_rgba(img.as_raw(), 64, 64, 0.0f, 0.0f) == img.as_raw(), "b/c identity");
    });
    // TiledImage semantics
    run("tiled_image_semantics", [] {
        TiledImage t(130, 70);
        check(t.chunk_keys().empty() && t.get_pixel(129, 69) == Rgba{{0, 0, 0, 0}}, "empty image");
        t.put_pixel(129, 69, Rgba{{1, 2, 3, 4}});
        check(t.chunk_keys().size() == 1 && t.chunk_keys()[0] == std::make_pair(2u, 1u), "one chunk populated");
        TiledImage c = t;  // COW clone shares chunks
        c.put_pixel(129, 69, Rgba{{9, 9, 9, 9}});
        check(t.get_pixel(129, 69) == Rgba{{1, 2, 3, 4}}, "copy-on-write");
        TiledImage huge(20000, 20000);
        check(huge.width() == 1 && huge.height() == 1, "dimension clamp (tiled_image.rs:17)");
        RgbaImage flat(70, 70);
        flat.put_pixel(3, 3, Rgba{{10, 20, 30, 0}});   // transparent chunk: dropped, RGB lost
        flat.put_pixel(69, 69, Rgba{{10, 20, 30, 40}});
        TiledImage r = TiledImage::from_rgba_image(flat);
        check(r.chunk_keys().size() == 1 && r.to_rgba_image().get_pixel(3, 3) == Rgba{{0, 0, 0, 0}}, "from_rgba_image drops transparent chunks");
    });

    std::printf("%d tests, %d failed\n", g_run, g_fail);
    return g_fail ? 1 : 0;
}
