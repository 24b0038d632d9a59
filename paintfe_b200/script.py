"""The Rhai Effect API surface that the CLI batch mode drives (src/ops/scripting.rs:822-1165),
as a line-oriented runner for straight-line effect scripts such as

    apply_blur(4.0); apply_hsl(10.0, 15.0, 0.0); apply_vignette(0.5, 0.3);

PaintFE's Rhai host itself is kept as the caller; this runner exists so the batch
benchmark (config 5) and the parity tests can execute the same effect sequences on the device
without a Rust toolchain.  Each binding maps to the entry point that replaces the reference function
it calls, with the reference's fixed arguments (e.g. apply_sharpen uses radius 1.0, scripting.rs:847).
Covered: every `apply_*` effect, the flips / rotations, `resize_image`, `resize_canvas` and the selection API
(`select_rect`, `select_ellipse`, `clear_selection`, `invert_selection`, `fill_selected`, `delete_selected`).
`execute_script_sync` runs a script through paintfe_b200.rhai_host, the interpreter for the host language around
these calls (variables, control flow, closures, the pixel / utility API); `parse` remains as the splitter for
straight-line scripts.
"""
from __future__ import annotations

import math
import re
from typing import Callable, Dict, List, Tuple

import numpy as np

from . import _lib as L

_CALL = re.compile(r"\s*([A-Za-z_][A-Za-z0-9_]*)\s*\(([^()]*)\)\s*$")

# parse_script_filter / parse_anchor, scripting.rs:60-82
_FILTERS = {"nearest": 0, "nn": 0, "bicubic": 2, "catmull": 2, "catmullrom": 2, "lanczos": 3, "lanczos3": 3}
_ANCHORS = {"top-left": (0, 0), "tl": (0, 0), "nw": (0, 0), "top-center": (1, 0), "tc": (1, 0), "n": (1, 0), "top": (1, 0),
            "top-right": (2, 0), "tr": (2, 0), "ne": (2, 0), "center-left": (0, 1), "cl": (0, 1), "w": (0, 1), "left": (0, 1),
            "center": (1, 1), "c": (1, 1), "middle": (1, 1), "center-right": (2, 1), "cr": (2, 1), "e": (2, 1), "right": (2, 1),
            "bottom-left": (0, 2), "bl": (0, 2), "sw": (0, 2), "bottom-center": (1, 2), "bc": (1, 2), "s": (1, 2), "bottom": (1, 2),
            "bottom-right": (2, 2), "br": (2, 2), "se": (2, 2)}


def _dim(v) -> int:
    return min(max(int(v), 1), 32768)  # `(new_w.max(1) as u32).min(32768)`, scripting.rs:750

# pfe_adjust_op ids (include/pfe_b200.h)
S_INVERT, S_DESATURATE, S_SEPIA, S_SEPIA_STRENGTH, S_BRIGHTNESS_CONTRAST, S_HSL, S_EXPOSURE, S_LUT_RGB = range(32, 40)


def _f32(v) -> float:
    return float(np.float32(v))


def bindings(eng, exact: bool = False) -> Dict[str, Callable]:
    """name -> f(img, mask, *args) -> img, mirroring register_effect_api (scripting.rs:822)."""

    def exposure_gain(ev):
        return _f32(np.power(np.float32(2.0), np.float32(ev)))  # 2.0f32.powf(ev as f32)

    return {
        # blur family goes through apply_effect_to_context: selection mask honoured (scripting.rs:617)
        "apply_blur": lambda im, m, sigma: eng.gaussian_blur(im, _f32(sigma), mask=m, exact=exact),
        "apply_box_blur": lambda im, m, radius: eng.box_blur(im, _f32(int(radius)), mask=m),
        "apply_motion_blur": lambda im, m, angle, distance: eng.motion_blur(im, _f32(angle), _f32(distance), mask=m),
        "apply_sharpen": lambda im, m, amount: eng.sharpen(im, _f32(amount), 1.0, mask=m, exact=exact),
        "apply_median": lambda im, m, radius: eng.median(im, max(int(radius), 1), mask=m),
        "apply_vignette": lambda im, m, strength, softness: eng.vignette(im, _f32(strength), _f32(softness), mask=m),
        # scripting.rs:1077-1131: fixed arguments of the bindings (noise: Gaussian, seed 42, scale 1, 1 octave;
        # reduce_noise radius 2)
        "apply_glow": lambda im, m, radius, intensity: eng.glow(im, _f32(radius), _f32(intensity), mask=m, exact=exact),
        "apply_pixelate": lambda im, m, size: eng.pixelate(im, max(int(size), 1), mask=m),
        "apply_bulge": lambda im, m, amount: eng.bulge(im, _f32(amount), mask=m),
        "apply_twist": lambda im, m, angle: eng.twist(im, _f32(angle), mask=m),
        "apply_noise": lambda im, m, amount, mono: eng.add_noise(im, _f32(amount), 1, bool(mono), 42, 1.0, 1, mask=m),
        "apply_reduce_noise": lambda im, m, strength: eng.reduce_noise(im, _f32(strength), 2, mask=m),
        # scripting.rs:1103-1164: fixed seed 42, 45 degree circle screen, 20 levels
        "apply_crystallize": lambda im, m, size: eng.crystallize(im, _f32(max(int(size), 1)), 42, mask=m),
        "apply_halftone": lambda im, m, dot_size: eng.halftone(im, _f32(dot_size), 45.0, 0, mask=m),
        "apply_ink": lambda im, m, strength, threshold: eng.ink(im, _f32(strength), _f32(threshold), mask=m),
        "apply_oil_painting": lambda im, m, radius: eng.oil_painting(im, max(int(radius), 1), 20, mask=m),
        # inline variants: truncating casts, no mask, alpha untouched (scripting.rs:869-1075)
        "apply_invert": lambda im, m: eng.adjust(im, S_INVERT),
        "apply_desaturate": lambda im, m: eng.adjust(im, S_DESATURATE),
        # scripting.rs:645-745: whole-buffer flips / turns (the selection mask does not apply); the
        # canvas-wide forms transform the script's buffer the same way and leave the replay on the other
        # layers to the caller (CanvasOpRequest)
        "flip_horizontal": lambda im, m: eng.orient(im, 0),
        "flip_vertical": lambda im, m: eng.orient(im, 1),
        "rotate_180": lambda im, m: eng.orient(im, 4),
        "flip_canvas_horizontal": lambda im, m: eng.orient(im, 0),
        "flip_canvas_vertical": lambda im, m: eng.orient(im, 1),
        "rotate_canvas_90cw": lambda im, m: eng.orient(im, 2),
        "rotate_canvas_90ccw": lambda im, m: eng.orient(im, 3),
        "rotate_canvas_180": lambda im, m: eng.orient(im, 4),
        # scripting.rs:744-815: imageops::resize of the buffer / re-anchored copy on a transparent canvas
        "resize_image": lambda im, m, w, h, method="bilinear": (
            im if (_dim(w), _dim(h)) == (im.shape[1], im.shape[0])
            else eng.resize(im, _dim(w), _dim(h), _FILTERS.get(str(method).lower(), 1))),
        "resize_canvas": lambda im, m, w, h, anchor="top-left": eng.resize_canvas(
            im, _dim(w), _dim(h), _ANCHORS.get(str(anchor).lower(), (0, 0)), (0, 0, 0, 0)),
        "apply_sepia": lambda im, m, *s: (eng.adjust(im, S_SEPIA) if not s else
                                          eng.adjust(im, S_SEPIA_STRENGTH, (min(max(float(s[0]), 0.0), 1.0),))),
        "apply_brightness_contrast": lambda im, m, b, c: eng.adjust(im, S_BRIGHTNESS_CONTRAST, (_f32(b), _f32(c))),
        "apply_hsl": lambda im, m, h, s, l: eng.adjust(im, S_HSL, (_f32(h), _f32(s), _f32(l))),
        "apply_exposure": lambda im, m, ev: eng.adjust(im, S_EXPOSURE, (exposure_gain(ev),)),
        "apply_levels": lambda im, m, black, white, gamma: eng.adjust(
            im, S_LUT_RGB, luts=eng.levels_lut_script(_f32(black), _f32(white), _f32(gamma))),
    }


def parse(source: str) -> List[Tuple[str, Tuple[float, ...]]]:
    """Split a straight-line script into (function, args). Comments (`// ...`) are ignored."""
    calls = []
    src = re.sub(r"//[^\n]*", "", source)
    for stmt in src.split(";"):
        if not stmt.strip():
            continue
        m = _CALL.match(stmt)
        if not m:
            raise ValueError(f"unsupported script statement (only apply_*(numbers) calls): {stmt.strip()!r}")
        lit = {"true": 1.0, "false": 0.0}

        def arg(a):
            a = a.strip()
            if len(a) >= 2 and a[0] == a[-1] == '"':
                return a[1:-1]  # string literal (resize_image's method, resize_canvas' anchor)
            return lit[a] if a in lit else float(a)

        args = tuple(arg(a) for a in m.group(2).split(",") if a.strip())
        calls.append((m.group(1), args))
    return calls


def execute_script_sync(eng, source: str, pixels, mask=None, exact: bool = False):
    """Shape of scripting::execute_script_sync (scripting.rs:1733): flat RGBA in, flat RGBA out.
    `pixels` may be a numpy array (host tier) or a CUDA tensor (device tier: the whole script runs
    without leaving the device). `exact` selects the bit-exact Gaussian (PFE_GAUSS_EXACT) over the
    default FMA path (<= 1 level)."""
    from .rhai_host import run_script
    return run_script(eng, source, pixels, mask, exact)[0]


# Calls that log a CanvasOpRequest (scripting.rs:687-813): the caller replays them on every other layer.
CANVAS_OPS = {"flip_canvas_horizontal", "flip_canvas_vertical", "rotate_canvas_90cw", "rotate_canvas_90ccw", "rotate_canvas_180",
              "resize_image", "resize_canvas"}


def apply_canvas_ops(eng, layers, active_index: int, canvas_ops):
    """scripting::apply_canvas_ops (scripting.rs:1640): replay the logged canvas-wide calls, in order, on every layer
    image except the active one (which the script already transformed).  `layers` are flat RGBA images (host arrays or
    device tensors) of the pre-script canvas size; returns the new list."""
    table = bindings(eng)
    out = list(layers)
    for name, args in canvas_ops:
        for i, img in enumerate(out):
            if i != active_index and img is not None:
                out[i] = table[name](img, None, *args)
    return out


# Selection API of the script host (scripting.rs:1359-1481): the mask is host state of the script context;
# fill / delete run on the device as a masked constant fill.
_SELECTION_API = {"select_rect", "select_ellipse", "clear_selection", "invert_selection", "fill_selected", "delete_selected"}


def _selection_call(eng, name, args, img, mask):
    h, w = int(img.shape[0]), int(img.shape[1])
    on_dev = not isinstance(img, np.ndarray)

    def to_side(m):
        if m is None or not on_dev:
            return m
        import torch as _t
        return _t.from_numpy(np.ascontiguousarray(m)).to(img.device)

    def host(m):
        return m if (m is None or isinstance(m, np.ndarray)) else m.cpu().numpy()

    if name == "select_rect":  # :1359-1375, half-open [x1, x2) x [y1, y2)
        x1, y1, x2, y2 = (int(a) for a in args)
        cl = lambda v, hi: min(max(v, 0), hi)
        m = np.zeros((h, w), np.uint8)
        m[cl(y1, h):cl(y2, h), cl(x1, w):cl(x2, w)] = 255
        return img, to_side(m)
    if name == "select_ellipse":  # :1380-1399, f64 arithmetic
        cx, cy, rx, ry = (float(a) for a in args)
        rx2, ry2 = max(rx * rx, 0.001), max(ry * ry, 0.001)
        dx = np.arange(w, dtype=np.float64)[None, :] - cx
        dy = np.arange(h, dtype=np.float64)[:, None] - cy
        m = np.where((dx * dx) / rx2 + (dy * dy) / ry2 <= 1.0, 255, 0).astype(np.uint8)
        return img, to_side(m)
    if name == "clear_selection":
        return img, None
    if name == "invert_selection":  # :1418-1431: no selection -> nothing selected
        m = host(mask)
        m = np.zeros((h, w), np.uint8) if m is None else (255 - m).astype(np.uint8)
        return img, to_side(m)
    # fill_selected(r, g, b, a) / delete_selected(): every selected pixel becomes the colour (:1435-1481).
    # canvas_border with a border as wide as the canvas is exactly that masked constant fill.
    color = (0, 0, 0, 0) if name == "delete_selected" else tuple(min(max(int(a), 0), 255) for a in args)
    return eng.canvas_border(img, 1 << 30, color, mask=mask), mask
