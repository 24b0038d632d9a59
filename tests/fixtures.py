"""Deterministic test images, restating the reference's generators.

Reference: tests/common/mod.rs:272-340 (gradient, checkerboard, solid, colour bands),
tests/visual_blend.rs:19-36 (blend foreground), tests/transform_ops.rs:25-45 (gradient_32,
uniform_grid).  No RNG; closed-form so the reference's golden PNGs can be reproduced here.
"""
from __future__ import annotations

import os

import numpy as np
from PIL import Image

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref")


def golden(category: str, name: str) -> np.ndarray:
    """Load tests/golden/ref/<category>/<name>.png as RGBA8 (tests/common/mod.rs:190-193)."""
    return np.array(Image.open(os.path.join(GOLDEN_DIR, category, name + ".png")).convert("RGBA"))


def gradient(w: int, h: int) -> np.ndarray:
    x = np.arange(w, dtype=np.uint32)
    y = np.arange(h, dtype=np.uint32)
    r = (x * 255 // (w - 1)).astype(np.uint8) if w > 1 else np.full(w, 128, np.uint8)
    b = (y * 255 // (h - 1)).astype(np.uint8) if h > 1 else np.full(h, 128, np.uint8)
    img = np.empty((h, w, 4), np.uint8)
    img[..., 0] = r[None, :]
    img[..., 1] = 255 - r[None, :]
    img[..., 2] = b[:, None]
    img[..., 3] = 255
    return img


def checkerboard(w: int, h: int) -> np.ndarray:
    x = np.arange(w)[None, :] // 8
    y = np.arange(h)[:, None] // 8
    v = np.where((x + y) % 2 == 0, 255, 0).astype(np.uint8)
    img = np.empty((h, w, 4), np.uint8)
    img[..., 0] = img[..., 1] = img[..., 2] = v
    img[..., 3] = 255
    return img


def solid(w: int, h: int, color) -> np.ndarray:
    img = np.empty((h, w, 4), np.uint8)
    img[...] = np.asarray(color, np.uint8)
    return img


def color_bands(w: int, h: int) -> np.ndarray:
    colors = np.array([[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255], [0, 255, 255, 255],
                       [255, 0, 255, 255], [255, 255, 0, 255], [255, 255, 255, 255], [0, 0, 0, 255]], np.uint8)
    band = np.minimum(np.arange(w) * 8 // w, 7)
    return np.broadcast_to(colors[band][None, :, :], (h, w, 4)).copy()


def blend_foreground(w: int = 64, h: int = 64) -> np.ndarray:
    """tests/visual_blend.rs:26-36 — f32 arithmetic, truncating casts."""
    x = np.arange(w, dtype=np.float32)[None, :]
    y = np.arange(h, dtype=np.float32)[:, None]
    img = np.empty((h, w, 4), np.uint8)
    img[..., 0] = ((x / np.float32(w)) * np.float32(255.0)).astype(np.uint8) + np.zeros((h, 1), np.uint8)
    img[..., 1] = ((y / np.float32(h)) * np.float32(255.0)).astype(np.uint8) + np.zeros((1, w), np.uint8)
    img[..., 2] = 128
    s = (np.arange(w, dtype=np.uint32)[None, :] + np.arange(h, dtype=np.uint32)[:, None]).astype(np.float32)
    img[..., 3] = ((s / np.float32(w + h - 2)) * np.float32(200.0) + np.float32(55.0)).astype(np.uint8)
    return img


def gradient_32() -> np.ndarray:
    img = np.empty((32, 32, 4), np.uint8)
    img[..., 0] = (np.arange(32, dtype=np.uint8) * 8)[None, :]
    img[..., 1] = (np.arange(32, dtype=np.uint8) * 8)[:, None]
    img[..., 2] = 128
    img[..., 3] = 255
    return img


def uniform_grid(cols: int, rows: int, w: float, h: float) -> np.ndarray:
    pts = np.empty(((rows + 1) * (cols + 1), 2), np.float32)
    i = 0
    for r in range(rows + 1):
        for c in range(cols + 1):
            pts[i, 0] = np.float32(c) / np.float32(cols) * np.float32(w)
            pts[i, 1] = np.float32(r) / np.float32(rows) * np.float32(h)
            i += 1
    return pts


def swirl_field() -> np.ndarray:
    """tests/transform_ops.rs:345-357, strict f32."""
    f = np.zeros((32, 32, 2), np.float32)
    f32 = np.float32
    for y in range(32):
        for x in range(32):
            dx = f32(x) - f32(16.0)
            dy = f32(y) - f32(16.0)
            r = max(np.sqrt(f32(dx * dx) + f32(dy * dy), dtype=np.float32), f32(0.001))
            strength = max(f32(1.0) - f32(r / f32(16.0)), f32(0.0))
            f[y, x, 0] = f32(f32(-dy * strength) * f32(0.5))
            f[y, x, 1] = f32(f32(dx * strength) * f32(0.5))
    return f


def random_rgba(rng: np.random.Generator, w: int, h: int, alpha: str = "uniform") -> np.ndarray:
    img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    if alpha == "binary":
        img[..., 3] = np.where(img[..., 3] > 127, 255, 0)
    elif alpha == "opaque":
        img[..., 3] = 255
    return img


def diff_stats(a: np.ndarray, b: np.ndarray):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return int((d.max(axis=-1) > 0).sum()) if d.ndim == 3 else int((d > 0).sum()), int(d.max()) if d.size else 0
