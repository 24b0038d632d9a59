"""ctypes binding for the CPU oracle (oracle/pfe_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  Nothing under paintfe_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpfe_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pfe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class LayerDesc(C.Structure):
    _fields_ = [
        ("rgba", C.c_void_p),
        ("mask", C.c_void_p),
        ("opacity", C.c_float),
        ("blend", C.c_uint8),
        ("visible", C.c_uint8),
        ("kind", C.c_uint8),
        ("_pad", C.c_uint8),
        ("adj", C.c_float * 16),
    ]


class Brush(C.Structure):
    _fields_ = [
        ("size", C.c_float),
        ("hardness", C.c_float),
        ("flow", C.c_float),
        ("anti_aliased", C.c_int),
        ("color", C.c_float * 4),
        ("is_eraser", C.c_int),
        ("mode", C.c_int),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.pfo_brush_alpha.restype = C.c_float
        _lib.pfo_to_radians.restype = C.c_float
        _lib.pfo_build_gaussian_kernel.restype = C.c_int
        _lib.pfo_gaussian_radius.restype = C.c_int
        _lib.pfo_brush_line_centres.restype = C.c_int
        _lib.pfo_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().pfo_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads for every later call (overrides an inherited OMP_NUM_THREADS). Returns the count in effect."""
    lib().pfo_set_num_threads(int(n))
    return num_threads()


# -- flatten ----------------------------------------------------------------------------
def make_layer(rgba=None, opacity=1.0, blend=0, visible=True, mask=None, kind=0, adj=()):
    return dict(rgba=_u8(rgba), opacity=float(opacity), blend=int(blend), visible=bool(visible),
                mask=_u8(mask), kind=int(kind), adj=tuple(float(v) for v in adj))


def flatten(layers, w, h, active=None):
    arr = (LayerDesc * len(layers))()
    keep = []
    for i, L in enumerate(layers):
        if L["rgba"] is not None:
            assert L["rgba"].shape == (h, w, 4)
            keep.append(L["rgba"])
            arr[i].rgba = L["rgba"].ctypes.data
        if L["mask"] is not None:
            assert L["mask"].shape == (h, w)
            keep.append(L["mask"])
            arr[i].mask = L["mask"].ctypes.data
        arr[i].opacity = L["opacity"]
        arr[i].blend = L["blend"]
        arr[i].visible = 1 if L["visible"] else 0
        arr[i].kind = L["kind"]
        for j, v in enumerate(L["adj"]):
            arr[i].adj[j] = v
    dst = np.empty((h, w, 4), np.uint8)
    active = _u8(active)
    lib().pfo_flatten(arr, C.c_uint32(len(layers)), C.c_uint32(w), C.c_uint32(h), _p(active), _p(dst))
    return dst


def blend_pixel(base, top, mode, opacity):
    b = (C.c_uint8 * 4)(*base)
    t = (C.c_uint8 * 4)(*top)
    o = (C.c_uint8 * 4)()
    lib().pfo_blend_pixel(b, t, C.c_int(mode), C.c_float(opacity), o)
    return tuple(o)


# -- filters ----------------------------------------------------------------------------
def _img_call(fn, src, *args, mask=None):
    src = _u8(src)
    h, w = src.shape[:2]
    dst = np.empty_like(src)
    mask = _u8(mask)
    fn(_p(src), C.c_uint32(w), C.c_uint32(h), *args, _p(mask), _p(dst))
    return dst


def gaussian_kernel(sigma):
    r = lib().pfo_gaussian_radius(C.c_float(sigma))
    k = np.empty(2 * r + 1, np.float32)
    lib().pfo_build_gaussian_kernel(C.c_float(sigma), _p(k))
    return k


def gaussian_blur(src, sigma, mask=None):
    return _img_call(lib().pfo_blur_with_selection, src, C.c_float(sigma), mask=mask)


def box_blur(src, radius, mask=None):
    return _img_call(lib().pfo_box_blur, src, C.c_float(radius), mask=mask)


def motion_blur(src, angle_deg, distance, mask=None):
    return _img_call(lib().pfo_motion_blur, src, C.c_float(angle_deg), C.c_float(distance), mask=mask)


def median(src, radius, mask=None):
    return _img_call(lib().pfo_median, src, C.c_uint32(radius), mask=mask)


def sharpen(src, amount, radius, mask=None):
    return _img_call(lib().pfo_sharpen, src, C.c_float(amount), C.c_float(radius), mask=mask)


def vignette(src, amount, softness, mask=None):
    return _img_call(lib().pfo_vignette, src, C.c_float(amount), C.c_float(softness), mask=mask)


# -- adjustments ------------------------------------------------------------------------
INVERT, INVERT_ALPHA, SEPIA, DESATURATE, BRIGHTNESS_CONTRAST, HSL, EXPOSURE, LUT_RGB, LUT_RGBA, \
    TEMPERATURE_TINT, HIGHLIGHTS_SHADOWS, THRESHOLD, POSTERIZE, COLOR_BALANCE, GRADIENT_MAP, BLACK_AND_WHITE, \
    VIBRANCE = range(17)
S_INVERT, S_DESATURATE, S_SEPIA, S_SEPIA_STRENGTH, S_BRIGHTNESS_CONTRAST, S_HSL, S_EXPOSURE, \
    S_LUT_RGB = range(32, 40)


def adjust(src, op, params=(), luts=None, mask=None, occupancy=None):
    src = _u8(src)
    h, w = src.shape[:2]
    dst = np.empty_like(src)
    p = np.zeros(12, np.float32)
    p[: len(params)] = params
    luts = _u8(luts)
    mask = _u8(mask)
    occupancy = _u8(occupancy)
    lib().pfo_adjust(_p(src), C.c_uint32(w), C.c_uint32(h), C.c_int(op), _p(p), _p(luts), _p(mask),
                     _p(occupancy), _p(dst))
    return dst


def levels_lut(in_black, in_white, gamma, out_black=0.0, out_white=255.0):
    lut = np.empty(256, np.uint8)
    lib().pfo_build_levels_lut(C.c_float(in_black), C.c_float(in_white), C.c_float(gamma),
                               C.c_float(out_black), C.c_float(out_white), _p(lut))
    return lut


def levels_lut_script(in_black, in_white, gamma):
    lut = np.empty(256, np.uint8)
    lib().pfo_build_levels_lut_script(C.c_float(in_black), C.c_float(in_white), C.c_float(gamma), _p(lut))
    return lut


def stretch_lut(mn, mx):
    lut = np.empty(256, np.uint8)
    lib().pfo_build_stretch_lut(C.c_uint8(mn), C.c_uint8(mx), _p(lut))
    return lut


def curves_lut(points):
    pts = _f32(np.asarray(points, np.float32).reshape(-1, 2))
    lut = np.empty(256, np.uint8)
    lib().pfo_build_curves_lut(_p(pts), C.c_int(len(pts)), _p(lut))
    return lut


def compose_curve_luts(five):
    five = _u8(np.asarray(five, np.uint8).reshape(5, 256))
    out = np.empty((4, 256), np.uint8)
    lib().pfo_compose_curve_luts(_p(five), _p(out))
    return out


# -- warps ------------------------------------------------------------------------------
def catmull_rom_weights(t):
    w = np.empty(4, np.float32)
    lib().pfo_catmull_rom_weights(C.c_float(t), _p(w))
    return w


def catmull_rom_surface(points, cols, rows, u, v):
    pts = _f32(points)
    out = np.empty(2, np.float32)
    lib().pfo_catmull_rom_surface(_p(pts), C.c_int(cols), C.c_int(rows), C.c_float(u), C.c_float(v), _p(out))
    return out


def mesh_displacement(orig, deformed, cols, rows, w, h):
    orig = _f32(orig)
    deformed = _f32(deformed)
    out = np.empty((h, w, 2), np.float32)
    lib().pfo_mesh_displacement(_p(orig), _p(deformed), C.c_int(cols), C.c_int(rows), C.c_uint32(w),
                                C.c_uint32(h), _p(out))
    return out


def warp_displacement(src, disp):
    src = _u8(src)
    disp = _f32(disp)
    sh, sw = src.shape[:2]
    h, w = disp.shape[:2]
    dst = np.empty((h, w, 4), np.uint8)
    lib().pfo_warp_displacement(_p(src), C.c_uint32(sw), C.c_uint32(sh), _p(disp), C.c_uint32(w),
                                C.c_uint32(h), _p(dst))
    return dst


def warp_displacement_region(src, disp, prev, rect):
    src, disp, prev = _u8(src), _f32(disp), _u8(prev)
    sh, sw = src.shape[:2]
    h, w = disp.shape[:2]
    dst = np.empty((h, w, 4), np.uint8)
    r = (C.c_int * 4)(*[int(v) for v in rect])
    lib().pfo_warp_displacement_region(_p(src), C.c_uint32(sw), C.c_uint32(sh), _p(disp), _p(prev), r, C.c_uint32(w),
                                       C.c_uint32(h), _p(dst))
    return dst


def mesh_warp(src, orig, deformed, cols, rows, w, h):
    src = _u8(src)
    sh, sw = src.shape[:2]
    orig = _f32(orig)
    deformed = _f32(deformed)
    dst = np.empty((h, w, 4), np.uint8)
    lib().pfo_mesh_warp(_p(src), C.c_uint32(sw), C.c_uint32(sh), _p(orig), _p(deformed), C.c_int(cols),
                        C.c_int(rows), C.c_uint32(w), C.c_uint32(h), _p(dst))
    return dst


PUSH, EXPAND, CONTRACT, TWIRL = range(4)


def liquify(field, kind, cx, cy, radius, strength, a0=0.0, a1=0.0):
    """In place on `field` (h, w, 2) float32; returns the (x0, y0, x1, y1) bbox."""
    assert field.dtype == np.float32 and field.flags.c_contiguous
    h, w = field.shape[:2]
    bbox = (C.c_int * 4)()
    lib().pfo_liquify(_p(field), C.c_uint32(w), C.c_uint32(h), C.c_int(kind), C.c_float(cx), C.c_float(cy),
                      C.c_float(radius), C.c_float(strength), C.c_float(a0), C.c_float(a1), bbox)
    return tuple(bbox)


# -- brush ------------------------------------------------------------------------------
BRUSH_NORMAL, BRUSH_DODGE, BRUSH_BURN, BRUSH_SPONGE = range(4)


def make_brush(size, hardness, anti_aliased, color, flow=1.0, is_eraser=False, mode=0):
    b = Brush()
    b.size, b.hardness, b.flow = size, hardness, flow
    b.anti_aliased = 1 if anti_aliased else 0
    for i in range(4):
        b.color[i] = color[i]
    b.is_eraser = 1 if is_eraser else 0
    b.mode = int(mode)
    return b


def brush_lut(brush):
    lut = np.empty(256, np.uint8)
    lib().pfo_brush_lut(C.byref(brush), _p(lut))
    return lut


def brush_stamp(img, brush, cx, cy, sel_mask=None):
    assert img.dtype == np.uint8 and img.flags.c_contiguous
    h, w = img.shape[:2]
    sel_mask = _u8(sel_mask)
    lib().pfo_brush_stamp(_p(img), C.c_uint32(w), C.c_uint32(h), C.byref(brush), C.c_float(cx), C.c_float(cy),
                          _p(sel_mask))


def brush_line_centres(w, h, x0, y0, x1, y1):
    cap = int(np.ceil(np.hypot(x1 - x0, y1 - y0))) + 4
    c = np.empty((cap, 2), np.float32)
    n = lib().pfo_brush_line_centres(C.c_uint32(w), C.c_uint32(h), C.c_float(x0), C.c_float(y0),
                                     C.c_float(x1), C.c_float(y1), _p(c), C.c_int(cap))
    return c[:n].copy()


def brush_line(img, brush, x0, y0, x1, y1, sel_mask=None):
    assert img.dtype == np.uint8 and img.flags.c_contiguous
    h, w = img.shape[:2]
    sel_mask = _u8(sel_mask)
    lib().pfo_brush_line(_p(img), C.c_uint32(w), C.c_uint32(h), C.byref(brush), C.c_float(x0), C.c_float(y0),
                         C.c_float(x1), C.c_float(y1), _p(sel_mask))


# -- tiles ------------------------------------------------------------------------------
def tiled_roundtrip(src):
    src = _u8(src)
    h, w = src.shape[:2]
    occ = np.empty(((h + 63) // 64, (w + 63) // 64), np.uint8)
    dst = np.empty_like(src)
    lib().pfo_tiled_roundtrip(_p(src), C.c_uint32(w), C.c_uint32(h), _p(occ), _p(dst))
    return dst, occ


# -- widened scope: remaining Rhai Effect-API kernels ------------------------------------------
def glow(src, radius, intensity, mask=None):
    return _img_call(lib().pfo_glow, src, C.c_float(radius), C.c_float(intensity), mask=mask)


def pixelate(src, block_size, mask=None):
    return _img_call(lib().pfo_pixelate, src, C.c_uint32(block_size), mask=mask)


def bulge(src, amount, origin=(0.5, 0.5), mask=None):
    return _img_call(lib().pfo_bulge, src, C.c_float(amount), C.c_float(origin[0]), C.c_float(origin[1]), mask=mask)


def twist(src, angle_deg, origin=(0.5, 0.5), mask=None):
    return _img_call(lib().pfo_twist, src, C.c_float(angle_deg), C.c_float(origin[0]), C.c_float(origin[1]), mask=mask)


NOISE_UNIFORM, NOISE_GAUSSIAN, NOISE_PERLIN = range(3)


def add_noise(src, amount, noise_type, monochrome, seed, scale, octaves, mask=None):
    return _img_call(lib().pfo_add_noise, src, C.c_float(amount), C.c_int(noise_type), C.c_int(1 if monochrome else 0),
                     C.c_uint32(seed), C.c_float(scale), C.c_uint32(octaves), mask=mask)


def reduce_noise(src, strength, radius, mask=None):
    return _img_call(lib().pfo_reduce_noise, src, C.c_float(strength), C.c_uint32(radius), mask=mask)


# -- widened scope, part 2: the rest of src/ops/effects/ -------------------------------------------
def _rgba4(color):
    return (C.c_uint8 * 4)(*[int(v) for v in color])


def ink(src, edge_strength, threshold, mask=None):
    return _img_call(lib().pfo_ink, src, C.c_float(edge_strength), C.c_float(threshold), mask=mask)


def oil_painting(src, radius, levels, mask=None):
    return _img_call(lib().pfo_oil_painting, src, C.c_uint32(radius), C.c_uint32(levels), mask=mask)


CF_MULTIPLY, CF_SCREEN, CF_OVERLAY, CF_SOFT_LIGHT = range(4)


def color_filter(src, color, intensity, mode, mask=None):
    return _img_call(lib().pfo_color_filter, src, _rgba4(color), C.c_float(intensity), C.c_int(mode), mask=mask)


def contours(src, scale, frequency, line_width, color, seed, octaves, blend, mask=None):
    return _img_call(lib().pfo_contours, src, C.c_float(scale), C.c_float(frequency), C.c_float(line_width),
                     _rgba4(color), C.c_uint32(seed), C.c_uint32(octaves), C.c_float(blend), mask=mask)


def crystallize(src, cell_size, seed, mask=None):
    return _img_call(lib().pfo_crystallize, src, C.c_float(cell_size), C.c_uint32(seed), mask=mask)


def dents(src, scale, amount, seed, octaves, roughness, pinch, wrap, mask=None):
    return _img_call(lib().pfo_dents, src, C.c_float(scale), C.c_float(amount), C.c_uint32(seed), C.c_uint32(octaves),
                     C.c_float(roughness), C.c_int(1 if pinch else 0), C.c_int(1 if wrap else 0), mask=mask)


HT_CIRCLE, HT_SQUARE, HT_DIAMOND, HT_LINE = range(4)


def halftone(src, dot_size, angle_deg, shape, mask=None):
    return _img_call(lib().pfo_halftone, src, C.c_float(dot_size), C.c_float(angle_deg), C.c_int(shape), mask=mask)


def bokeh_blur(src, radius, mask=None):
    return _img_call(lib().pfo_bokeh_blur, src, C.c_float(radius), mask=mask)


def zoom_blur(src, center_x, center_y, strength, samples, tint=(0.0, 0.0, 0.0, 0.0), tint_strength=0.0, mask=None):
    t = (C.c_float * 4)(*tint)
    return _img_call(lib().pfo_zoom_blur, src, C.c_float(center_x), C.c_float(center_y), C.c_float(strength),
                     C.c_uint32(samples), t, C.c_float(tint_strength), mask=mask)


GRID_LINES, GRID_CHECKERBOARD = range(2)


def grid(src, cell_w, cell_h, line_width, color, style, opacity, mask=None):
    return _img_call(lib().pfo_grid, src, C.c_uint32(cell_w), C.c_uint32(cell_h), C.c_uint32(line_width),
                     _rgba4(color), C.c_int(style), C.c_float(opacity), mask=mask)


def canvas_border(src, width, color, mask=None):
    return _img_call(lib().pfo_canvas_border, src, C.c_uint32(width), _rgba4(color), mask=mask)


def drop_shadow(src, offset_x, offset_y, blur_radius, widen_radius, color, opacity, mask=None):
    return _img_call(lib().pfo_drop_shadow, src, C.c_int32(offset_x), C.c_int32(offset_y), C.c_float(blur_radius),
                     C.c_int(1 if widen_radius else 0), _rgba4(color), C.c_float(opacity), mask=mask)


OUTLINE_OUTSIDE, OUTLINE_INSIDE, OUTLINE_CENTER = range(3)


def outline(src, width, color, mode, anti_alias, mask=None):
    return _img_call(lib().pfo_outline, src, C.c_uint32(width), _rgba4(color), C.c_int(mode),
                     C.c_int(1 if anti_alias else 0), mask=mask)


def pixel_drag(src, seed, amount, distance, direction, mask=None):
    return _img_call(lib().pfo_pixel_drag, src, C.c_uint32(seed), C.c_float(amount), C.c_uint32(distance),
                     C.c_float(direction), mask=mask)


def rgb_displace(src, r_off, g_off, b_off, mask=None):
    off = (C.c_int32 * 6)(r_off[0], r_off[1], g_off[0], g_off[1], b_off[0], b_off[1])
    return _img_call(lib().pfo_rgb_displace, src, off, mask=mask)


# -- geometry: flips / quarter turns, resize_canvas, apply_affine, imageops::resize ----------------
FLIP_H, FLIP_V, ROT90CW, ROT90CCW, ROT180 = range(5)


def orient(src, op):
    src = _u8(src)
    h, w = src.shape[:2]
    dst = np.empty((w, h, 4) if op in (ROT90CW, ROT90CCW) else (h, w, 4), np.uint8)
    lib().pfo_orient(_p(src), C.c_uint32(w), C.c_uint32(h), C.c_int(op), _p(dst))
    return dst


def resize_canvas(src, new_w, new_h, anchor, fill):
    src = _u8(src)
    h, w = src.shape[:2]
    dst = np.empty((new_h, new_w, 4), np.uint8)
    lib().pfo_resize_canvas(_p(src), C.c_uint32(w), C.c_uint32(h), C.c_uint32(new_w), C.c_uint32(new_h),
                            C.c_uint32(anchor[0]), C.c_uint32(anchor[1]), _rgba4(fill), _p(dst))
    return dst


def affine(src, canvas_w, canvas_h, rotation_z, rotation_x=0.0, rotation_y=0.0, scale=1.0, offset=(0.0, 0.0), nearest=False):
    src = _u8(src)
    h, w = src.shape[:2]
    dst = np.empty((canvas_h, canvas_w, 4), np.uint8)
    lib().pfo_affine(_p(src), C.c_uint32(w), C.c_uint32(h), C.c_uint32(canvas_w), C.c_uint32(canvas_h),
                     C.c_float(rotation_z), C.c_float(rotation_x), C.c_float(rotation_y), C.c_float(scale),
                     C.c_float(offset[0]), C.c_float(offset[1]), C.c_int(1 if nearest else 0), _p(dst))
    return dst


RS_NEAREST, RS_TRIANGLE, RS_CATMULL_ROM, RS_LANCZOS3 = range(4)


def resize(src, new_w, new_h, filter):
    src = _u8(src)
    h, w = src.shape[:2]
    dst = np.empty((new_h, new_w, 4), np.uint8)
    lib().pfo_resize(_p(src), C.c_uint32(w), C.c_uint32(h), C.c_uint32(new_w), C.c_uint32(new_h), C.c_int(filter), _p(dst))
    return dst
