"""Pins the CPU oracle against the reference's own golden PNGs (tolerance 0).

Each case restates one reference test: inputs from tests/fixtures.py (closed-form, same as
tests/common/mod.rs), expected output = the reference's committed golden, copied as data into
tests/golden/ref/.  Reference tests: tests/visual_blend.rs, visual_filters.rs,
visual_adjustments.rs, scripting.rs, transform_ops.rs, tool_strokes.rs.
"""
import numpy as np
import pytest

import fixtures as fx

BLEND_IDS = {  # src/canvas/layers.rs:125-153
    "normal": 0, "multiply": 1, "screen": 2, "additive": 3, "reflect": 4, "glow": 5, "color_burn": 6,
    "color_dodge": 7, "overlay": 8, "difference": 9, "negation": 10, "lighten": 11, "darken": 12, "xor": 13,
    "overwrite": 14, "hard_light": 15, "soft_light": 16, "exclusion": 17, "subtract": 18, "divide": 19,
    "linear_burn": 20, "vivid_light": 21, "linear_light": 22, "pin_light": 23, "hard_mix": 24,
}


def assert_exact(actual, category, name):
    exp = fx.golden(category, name)
    assert actual.shape == exp.shape, (actual.shape, exp.shape)
    bad, mx = fx.diff_stats(actual, exp)
    assert bad == 0, f"{category}/{name}: {bad} mismatched pixels, max channel diff {mx}"


# ---- blend (tests/visual_blend.rs:19-106) ---------------------------------------------
@pytest.mark.parametrize("name", sorted(BLEND_IDS))
def test_blend_golden(oracle, name):
    layers = [oracle.make_layer(fx.checkerboard(64, 64)),
              oracle.make_layer(fx.blend_foreground(), blend=BLEND_IDS[name])]
    assert_exact(oracle.flatten(layers, 64, 64), "blend", name)


def test_blend_half_opacity(oracle):
    layers = [oracle.make_layer(fx.checkerboard(64, 64)), oracle.make_layer(fx.gradient(64, 64), opacity=0.5)]
    assert_exact(oracle.flatten(layers, 64, 64), "blend", "normal_half_opacity")


# ---- filters (tests/visual_filters.rs:30-176) -----------------------------------------
FILTERS = {
    "gaussian_blur_s2": lambda o, im: o.gaussian_blur(im, 2.0),
    "gaussian_blur_s5": lambda o, im: o.gaussian_blur(im, 5.0),
    "box_blur_r3": lambda o, im: o.box_blur(im, 3.0),
    "motion_blur_45_10": lambda o, im: o.motion_blur(im, 45.0, 10.0),
    "median_r2": lambda o, im: o.median(im, 2),
    "sharpen_a1_r1": lambda o, im: o.sharpen(im, 1.0, 1.0),
    "vignette_08_05": lambda o, im: o.vignette(im, 0.8, 0.5),
    # widened scope, tests/visual_filters.rs:88-160
    "glow_r3_i05": lambda o, im: o.glow(im, 3.0, 0.5),
    "pixelate_8": lambda o, im: o.pixelate(im, 8),
    "bulge_05": lambda o, im: o.bulge(im, 0.5),
    "twist_45": lambda o, im: o.twist(im, 45.0),
    "add_noise_uniform": lambda o, im: o.add_noise(im, 30.0, o.NOISE_UNIFORM, False, 42, 1.0, 1),
    "add_noise_gaussian_mono": lambda o, im: o.add_noise(im, 30.0, o.NOISE_GAUSSIAN, True, 42, 1.0, 1),
    "add_noise_perlin": lambda o, im: o.add_noise(im, 50.0, o.NOISE_PERLIN, False, 42, 5.0, 3),
    "reduce_noise": lambda o, im: o.reduce_noise(im, 0.5, 2),
    # widened scope, part 2: the rest of src/ops/effects/ (tests/visual_filters.rs:44-283)
    "bokeh_blur_r5": lambda o, im: o.bokeh_blur(im, 5.0),
    "zoom_blur": lambda o, im: o.zoom_blur(im, 0.5, 0.5, 0.3, 8),
    "crystallize_s16": lambda o, im: o.crystallize(im, 16.0, 42),
    "dents": lambda o, im: o.dents(im, 20.0, 10.0, 42, 2, 0.5, False, False),
    "halftone_circle": lambda o, im: o.halftone(im, 4.0, 45.0, o.HT_CIRCLE),
    "grid_lines_16": lambda o, im: o.grid(im, 16, 16, 1, (0, 0, 0, 255), o.GRID_LINES, 1.0),
    "contours": lambda o, im: o.contours(im, 10.0, 5.0, 1.0, (0, 0, 0, 255), 42, 2, 0.5),
    "pixel_drag": lambda o, im: o.pixel_drag(im, 42, 50.0, 20, 0.0),
    "rgb_displace": lambda o, im: o.rgb_displace(im, (5, 0), (0, 0), (-5, 0)),
    "ink": lambda o, im: o.ink(im, 1.0, 0.5),
    "oil_painting": lambda o, im: o.oil_painting(im, 3, 20),
    "color_filter_multiply": lambda o, im: o.color_filter(im, (255, 128, 0, 255), 0.5, o.CF_MULTIPLY),
}


@pytest.mark.parametrize("name", sorted(FILTERS))
def test_filter_golden(oracle, name):
    assert_exact(FILTERS[name](oracle, fx.gradient(64, 64)), "filters", name)


def _square(fill):
    """tests/visual_filters.rs:193-214: transparent 64x64 with a solid 32x32 square in the centre."""
    img = fx.solid(64, 64, (0, 0, 0, 0))
    img[16:48, 16:48] = fill
    return img


def test_drop_shadow_golden(oracle):
    out = oracle.drop_shadow(_square((255, 255, 255, 255)), 5, 5, 3.0, False, (0, 0, 0, 255), 0.8)
    assert_exact(out, "filters", "drop_shadow")


def test_outline_outside_golden(oracle):
    out = oracle.outline(_square((255, 0, 0, 255)), 2, (0, 0, 255, 255), oracle.OUTLINE_OUTSIDE, True)
    assert_exact(out, "filters", "outline_outside")


def test_effect_known_answers(oracle):
    """tests/visual_filters.rs:226-236 (canvas border), :338-352 (colour filter identity)."""
    img = fx.solid(8, 8, (10, 20, 30, 255))
    out = oracle.canvas_border(img, 2, (200, 100, 50, 255))
    assert tuple(out[0, 0]) == (200, 100, 50, 255) and tuple(out[3, 3]) == (10, 20, 30, 255)
    g = fx.gradient(64, 64)
    assert np.array_equal(oracle.color_filter(g, (255, 255, 255, 255), 0.0, oracle.CF_MULTIPLY), g)


# ---- adjustments (tests/visual_adjustments.rs:50-216) ---------------------------------
def _auto_levels(o, im):
    sel = im[..., 3] != 0
    luts = np.empty((4, 256), np.uint8)
    for c in range(3):
        luts[c] = o.stretch_lut(int(im[..., c][sel].min()), int(im[..., c][sel].max()))
    luts[3] = np.arange(256, dtype=np.uint8)
    return o.adjust(im, o.LUT_RGBA, luts=luts)


ADJUST = {
    "invert_colors": lambda o, im: o.adjust(im, o.INVERT),
    "invert_alpha": lambda o, im: o.adjust(im, o.INVERT_ALPHA),
    "invert_alpha_double": lambda o, im: o.adjust(im, o.INVERT_ALPHA),
    "sepia": lambda o, im: o.adjust(im, o.SEPIA),
    "auto_levels": _auto_levels,
    "desaturate": lambda o, im: o.adjust(im, o.DESATURATE),
    "brightness_30_contrast_20": lambda o, im: o.adjust(im, o.BRIGHTNESS_CONTRAST, (30.0, 20.0)),
    "hsl_h30_s-20_l10": lambda o, im: o.adjust(im, o.HSL, (30.0, -20.0, 10.0)),
    "exposure_1ev": lambda o, im: o.adjust(im, o.EXPOSURE, (2.0,)),
    "highlights_shadows": lambda o, im: o.adjust(im, o.HIGHLIGHTS_SHADOWS, (30.0, -20.0)),
    "levels": lambda o, im: o.adjust(im, o.LUT_RGB, luts=o.levels_lut(20.0, 235.0, 1.2, 0.0, 255.0)),
    "temperature_tint": lambda o, im: o.adjust(im, o.TEMPERATURE_TINT, (30.0, 10.0)),
    # tests/visual_adjustments.rs:249-333
    "threshold_128": lambda o, im: o.adjust(im, o.THRESHOLD, (128.0,)),
    "posterize_4": lambda o, im: o.adjust(im, o.POSTERIZE, (4.0,)),
    "color_balance": lambda o, im: o.adjust(im, o.COLOR_BALANCE, (10.0, 0.0, -10.0, 0.0, 0.0, 0.0, -10.0, 0.0, 10.0)),
    "gradient_map": lambda o, im: o.adjust(im, o.GRADIENT_MAP, luts=warm_gradient_lut()),
    "black_and_white": lambda o, im: o.adjust(fx.color_bands(64, 64), o.BLACK_AND_WHITE, (0.3, 0.59, 0.11)),
    "vibrance_50": lambda o, im: o.adjust(im, o.VIBRANCE, (np.float32(50.0) / np.float32(100.0),)),
}


def warm_gradient_lut():
    """tests/visual_adjustments.rs:302-311, f32 arithmetic with truncating casts."""
    t = np.arange(256, dtype=np.float32) / np.float32(255.0)
    lut = np.empty((256, 4), np.uint8)
    lut[:, 0] = (t * np.float32(255.0)).astype(np.uint8)
    lut[:, 1] = (t * t * np.float32(200.0)).astype(np.uint8)
    lut[:, 2] = (t * t * t * np.float32(150.0)).astype(np.uint8)
    lut[:, 3] = 255
    return lut


@pytest.mark.parametrize("name", sorted(ADJUST))
def test_adjust_golden(oracle, name):
    out = ADJUST[name](oracle, fx.gradient(64, 64))
    # results go back through TiledImage::from_rgba_image + to_rgba_image (extract_layer)
    out, _ = oracle.tiled_roundtrip(out)
    assert_exact(out, "adjustments", name)


# ---- scripting inline variants (tests/scripting.rs:118-146) ---------------------------
SCRIPT = {
    "apply_blur": lambda o, im: o.gaussian_blur(im, 2.0),
    "apply_invert": lambda o, im: o.adjust(im, o.S_INVERT),
    "apply_sepia": lambda o, im: o.adjust(im, o.S_SEPIA),
    "apply_desaturate": lambda o, im: o.adjust(im, o.S_DESATURATE),
    "apply_brightness_contrast": lambda o, im: o.adjust(im, o.S_BRIGHTNESS_CONTRAST, (20.0, 10.0)),
    "apply_pixelate": lambda o, im: o.pixelate(im, 4),
}


@pytest.mark.parametrize("name", sorted(SCRIPT))
def test_scripting_golden(oracle, name):
    assert_exact(SCRIPT[name](oracle, fx.gradient(64, 64)), "scripting", name)


def test_adjust_identities(oracle):
    """tests/visual_adjustments.rs:283-296, :337-343"""
    g = fx.gradient(64, 64)
    assert np.array_equal(oracle.adjust(g, oracle.COLOR_BALANCE, (0.0,) * 9), g)
    assert np.array_equal(oracle.adjust(g, oracle.VIBRANCE, (0.0,)), g)


def test_scripting_bc_is_truncating(oracle):
    """the rounding variant must NOT match the scripting golden."""
    out = oracle.adjust(fx.gradient(64, 64), oracle.BRIGHTNESS_CONTRAST, (20.0, 10.0))
    bad, _ = fx.diff_stats(out, fx.golden("scripting", "apply_brightness_contrast"))
    assert bad > 0


# ---- warps (tests/transform_ops.rs:162-208, 345-359) ----------------------------------
def test_displacement_radial_push(oracle):
    field = np.zeros((32, 32, 2), np.float32)
    bbox = oracle.liquify(field, oracle.PUSH, 16.0, 16.0, 10.0, 0.8, 3.0, 0.0)
    assert bbox == (6, 6, 26, 26)
    assert_exact(oracle.warp_displacement(fx.gradient_32(), field), "transform", "displacement_radial_push")


def test_displacement_swirl(oracle):
    assert_exact(oracle.warp_displacement(fx.gradient_32(), fx.swirl_field()), "transform", "displacement_swirl")


def test_mesh_warp_deformed(oracle):
    orig = fx.uniform_grid(2, 2, 32.0, 32.0)
    deformed = orig.copy()
    deformed[4] = (20.0, 20.0)
    assert_exact(oracle.mesh_warp(fx.gradient_32(), orig, deformed, 2, 2, 32, 32), "transform", "mesh_warp_deformed")


def test_catmull_rom_known_answers(oracle):
    """tests/transform_ops.rs:51-118"""
    w0 = oracle.catmull_rom_weights(0.0)
    assert np.allclose(w0, [0, 1, 0, 0], atol=1e-6)
    w1 = oracle.catmull_rom_weights(1.0)
    assert np.allclose(w1, [0, 0, 1, 0], atol=1e-6)
    for i in range(11):
        assert abs(float(oracle.catmull_rom_weights(i / 10.0).sum()) - 1.0) < 1e-5
    grid = fx.uniform_grid(2, 2, 32.0, 32.0)
    p = oracle.catmull_rom_surface(grid, 2, 2, 1.0, 1.0)
    assert abs(p[0] - 16.0) < 0.5 and abs(p[1] - 16.0) < 0.5
    p0 = oracle.catmull_rom_surface(grid, 2, 2, 0.0, 0.0)
    assert abs(p0[0]) < 0.5 and abs(p0[1]) < 0.5


def test_warp_identity_and_translate(oracle):
    """tests/transform_ops.rs:125-160"""
    src = fx.gradient_32()
    zero = np.zeros((32, 32, 2), np.float32)
    assert np.array_equal(oracle.warp_displacement(src, zero), src)
    shift = zero.copy()
    shift[..., 0] = 5.0
    out = oracle.warp_displacement(src, shift)
    assert tuple(out[16, 10]) == tuple(src[16, 5])


def test_mesh_warp_identity(oracle):
    """tests/transform_ops.rs:170-195: identity mesh within 2 levels."""
    src = fx.gradient_32()
    grid = fx.uniform_grid(2, 2, 32.0, 32.0)
    out = oracle.mesh_warp(src, grid, grid, 2, 2, 32, 32)
    assert np.abs(out.astype(int) - src.astype(int)).max() <= 2


# ---- brush stamps (tests/tool_strokes.rs) ---------------------------------------------
BLACK, WHITE, RED, BLUE_SEMI = (0, 0, 0, 1), (1, 1, 1, 1), (1, 0, 0, 1), (0, 0, 1, 0.5)
# name -> (size, hardness, aa, colour, eraser, background, kind, args)
STROKES = {
    "brush_circle_center": (20.0, 1.0, True, BLACK, False, "blank", "stamp", [(32.0, 32.0)]),
    "brush_circle_soft": (30.0, 0.0, True, BLACK, False, "blank", "stamp", [(32.0, 32.0)]),
    "brush_circle_hard": (20.0, 1.0, False, BLACK, False, "blank", "stamp", [(32.0, 32.0)]),
    "brush_circle_tiny": (3.0, 1.0, True, RED, False, "blank", "stamp", [(32.0, 32.0)]),
    "brush_circle_large": (60.0, 0.5, True, BLACK, False, "blank", "stamp", [(32.0, 32.0)]),
    "brush_semi_transparent": (20.0, 1.0, True, BLUE_SEMI, False, "blank", "stamp", [(32.0, 32.0)]),
    "brush_secondary_color": (20.0, 1.0, True, RED, False, "blank", "stamp", [(32.0, 32.0)]),
    "eraser_circle": (20.0, 1.0, True, BLACK, True, "white", "stamp", [(32.0, 32.0)]),
    "eraser_soft": (30.0, 0.0, True, BLACK, True, "white", "stamp", [(32.0, 32.0)]),
    "line_horizontal": (8.0, 1.0, True, BLACK, False, "blank", "line", (4.0, 32.0, 60.0, 32.0)),
    "line_vertical": (8.0, 1.0, True, BLACK, False, "blank", "line", (32.0, 4.0, 32.0, 60.0)),
    "line_diagonal": (6.0, 0.8, True, BLACK, False, "blank", "line", (4.0, 4.0, 60.0, 60.0)),
    "line_soft_thick": (16.0, 0.3, True, RED, False, "blank", "line", (10.0, 50.0, 54.0, 10.0)),
    "line_eraser": (10.0, 1.0, True, BLACK, True, "white", "line", (4.0, 32.0, 60.0, 32.0)),
    "stroke_multiple_stamps": (10.0, 0.8, True, BLACK, False, "blank", "stamp",
                               [(8.0 + i * 7.0, 32.0) for i in range(8)]),
    "brush_at_origin": (10.0, 1.0, True, BLACK, False, "blank", "stamp", [(0.0, 0.0)]),
    "brush_at_corner": (20.0, 1.0, True, BLACK, False, "blank", "stamp", [(63.0, 63.0)]),
    "line_zero_length": (12.0, 1.0, True, BLACK, False, "blank", "line", (32.0, 32.0, 32.0, 32.0)),
    # tests/tool_strokes.rs:521-571: pencil = hard brush without anti-aliasing
    "pencil_circle": (12.0, 1.0, False, BLACK, False, "blank", "stamp", [(32.0, 32.0)]),
    "pencil_line": (4.0, 1.0, False, RED, False, "blank", "line", (4.0, 4.0, 60.0, 60.0)),
}


@pytest.mark.parametrize("name", sorted(STROKES))
def test_stroke_golden(oracle, name):
    size, hard, aa, color, eraser, bg, kind, args = STROKES[name]
    img = fx.solid(64, 64, (255, 255, 255, 255) if bg == "white" else (0, 0, 0, 0))
    b = oracle.make_brush(size, hard, aa, color, is_eraser=eraser)
    if kind == "stamp":
        for (x, y) in args:
            oracle.brush_stamp(img, b, x, y)
    else:
        oracle.brush_line(img, b, *args)
    assert_exact(img, "tools", name)


@pytest.mark.parametrize("name,mode", [("brush_dodge_mode", 1), ("brush_burn_mode", 2)])
def test_stroke_mode_golden(oracle, name, mode):
    """tests/tool_strokes.rs:472-515: dodge / burn edit the existing pixels in HSL space."""
    img = fx.gradient(64, 64)
    oracle.brush_stamp(img, oracle.make_brush(24.0, 1.0, True, BLACK, mode=mode), 32.0, 32.0)
    assert_exact(img, "tools", name)


def test_stroke_selection_mask(oracle):
    img = fx.solid(64, 64, (0, 0, 0, 0))
    mask = np.zeros((64, 64), np.uint8)
    mask[:, :32] = 255
    oracle.brush_stamp(img, oracle.make_brush(40.0, 1.0, True, BLACK), 32.0, 32.0, sel_mask=mask)
    assert_exact(img, "tools", "brush_with_selection_mask")


# ---- geometry (tests/visual_transforms.rs, tests/transform_ops.rs:280-303, tests/scripting.rs) -----
def _img_64x48():
    return fx.gradient(64, 48)


GEOMETRY = {
    "flip_canvas_h": lambda o, im: o.orient(im, o.FLIP_H),
    "flip_canvas_v": lambda o, im: o.orient(im, o.FLIP_V),
    "flip_layer_h": lambda o, im: o.orient(im, o.FLIP_H),
    "flip_layer_v": lambda o, im: o.orient(im, o.FLIP_V),
    "rotate_90cw": lambda o, im: o.orient(im, o.ROT90CW),
    "rotate_90ccw": lambda o, im: o.orient(im, o.ROT90CCW),
    "rotate_180": lambda o, im: o.orient(im, o.ROT180),
    "resize_canvas_center": lambda o, im: o.resize_canvas(im, 96, 80, (1, 1), (0, 0, 0, 0)),
    "resize_canvas_topleft": lambda o, im: o.resize_canvas(im, 80, 64, (0, 0), (255, 0, 0, 255)),
    "resize_2x_nearest": lambda o, im: o.resize(im, 128, 96, o.RS_NEAREST),
    "resize_half_bilinear": lambda o, im: o.resize(im, 32, 24, o.RS_TRIANGLE),
    "resize_half_lanczos": lambda o, im: o.resize(im, 32, 24, o.RS_LANCZOS3),
    # the reference test passes 45 degrees *already converted to radians* as the degree argument
    "affine_rotate_45": lambda o, im: o.affine(im, 64, 48, float(np.float32(45.0) * (np.float32(np.pi) / np.float32(180.0)))),
    "flatten_single": lambda o, im: o.flatten([o.make_layer(im)], 64, 48),
}


@pytest.mark.parametrize("name", sorted(GEOMETRY))
def test_geometry_golden(oracle, name):
    out = GEOMETRY[name](oracle, _img_64x48())
    if name != "flatten_single":
        out, _ = oracle.tiled_roundtrip(out)  # results are stored as tiles and read back (extract_layer)
    assert_exact(out, "transforms", name)


def test_affine_goldens(oracle):
    """tests/transform_ops.rs:280-303: affine then composite over a transparent canvas."""
    src = fx.gradient(32, 32)
    for name, rz, sc in (("affine_rotate_90", float(np.float32(np.pi / 2)), 1.0), ("affine_scale_half", 0.0, 0.5)):
        out = oracle.affine(src, 32, 32, rz, scale=sc)
        assert_exact(oracle.flatten([oracle.make_layer(out)], 32, 32), "transform", name)


def test_scripting_flip_goldens(oracle):
    """tests/scripting.rs: flip_horizontal() / flip_vertical() on the script's pixel buffer."""
    g = fx.gradient(64, 64)
    assert_exact(oracle.orient(g, oracle.FLIP_H), "scripting", "flip_horizontal")
    assert_exact(oracle.orient(g, oracle.FLIP_V), "scripting", "flip_vertical")


def test_geometry_identities(oracle):
    """tests/visual_transforms.rs:43-130, :250-262; transform_ops.rs:254-277."""
    im = _img_64x48()
    for op in (oracle.FLIP_H, oracle.FLIP_V, oracle.ROT180):
        assert np.array_equal(oracle.orient(oracle.orient(im, op), op), im)
    r = im
    for _ in range(4):
        r = oracle.orient(r, oracle.ROT90CW)
    assert np.array_equal(r, im)
    assert np.array_equal(oracle.orient(oracle.orient(im, oracle.ROT90CW), oracle.ROT90CCW), im)
    assert np.abs(oracle.affine(im, 64, 48, 0.0).astype(int) - im.astype(int)).max() <= 1
    dot = np.zeros((32, 32, 4), np.uint8)
    dot[16, 16] = (255, 0, 0, 255)
    assert np.array_equal(oracle.affine(dot, 32, 32, 0.0), dot)
    assert np.array_equal(oracle.resize(im, 64, 48, oracle.RS_LANCZOS3), im)
