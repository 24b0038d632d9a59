"""`.pfe` project files: reader for v0/v1/v2 raster layers and writer for v1.

Reference: src/io.rs:85-208 (structures), :296-340 (v1 writer), :477-497 (magic dispatch),
:1110-1146 (v1 loader and its validation).  The files are bincode 1.3.3 with the default options:
little-endian fixed-width integers, `u64` length prefixes for String / Vec, `usize` as u64, `bool`
as one byte, `Option` as a one-byte tag.  The magic string "PFE<n>" therefore sits at bytes 8..12.

This is harness plumbing on either side of the hot path: it lets the CLI run
`--flatten` on a multi-layer project (BASELINE config 1).  v3 (adjustment layers, HDR metadata) is
rejected with a clear error rather than half-parsed.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

CHUNK = 64
CHUNK_BYTES = CHUNK * CHUNK * 4
MAX_CANVAS_DIM = 25_000  # io.rs:500
MAX_LAYERS = 256         # io.rs:503


class PfeError(ValueError):
    pass


@dataclass
class PfeLayer:
    name: str
    visible: bool
    opacity: float
    blend_mode: int
    chunks: Dict[Tuple[int, int], np.ndarray] = field(default_factory=dict)  # (cx, cy) -> (64, 64, 4) u8

    def to_flat(self, w: int, h: int) -> np.ndarray:
        """TiledImage::to_rgba_image (tiled_image.rs:271-293)."""
        out = np.zeros((h, w, 4), np.uint8)
        for (cx, cy), t in self.chunks.items():
            x0, y0 = cx * CHUNK, cy * CHUNK
            if x0 >= w or y0 >= h:
                continue
            cw, ch = min(CHUNK, w - x0), min(CHUNK, h - y0)
            out[y0:y0 + ch, x0:x0 + cw] = t[:ch, :cw]
        return out

    def occupancy(self, w: int, h: int) -> np.ndarray:
        occ = np.zeros(((h + CHUNK - 1) // CHUNK, (w + CHUNK - 1) // CHUNK), np.uint8)
        for (cx, cy) in self.chunks:
            if cy < occ.shape[0] and cx < occ.shape[1]:
                occ[cy, cx] = 1
        return occ


@dataclass
class PfeProject:
    width: int
    height: int
    active_layer_index: int
    layers: List[PfeLayer]


class _Reader:
    def __init__(self, raw: bytes):
        self.b, self.o = raw, 0

    def take(self, n: int) -> bytes:
        if n < 0 or self.o + n > len(self.b):
            raise PfeError("unexpected end of file")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def u8(self): return self.take(1)[0]
    def u32(self): return struct.unpack("<I", self.take(4))[0]
    def u64(self): return struct.unpack("<Q", self.take(8))[0]
    def f32(self): return struct.unpack("<f", self.take(4))[0]
    def string(self): return self.take(self.u64()).decode("utf-8")
    def blob(self): return self.take(self.u64())


def load_pfe_from_bytes(raw: bytes) -> PfeProject:
    """io.rs:477-497 + the per-version loaders."""
    if len(raw) < 12:
        raise PfeError("File too small")
    magic = raw[8:12].decode("utf-8", "replace")
    if magic not in ("PFE0", "PFE1", "PFE2"):
        if magic == "PFE3":
            raise PfeError("PFE3 projects (adjustment layers / HDR metadata) are not supported by this reader")
        raise PfeError(f"Unknown magic '{magic}'")
    r = _Reader(raw)
    r.string()
    w, h = r.u32(), r.u32()
    if w == 0 or h == 0:
        raise PfeError("Image dimensions cannot be zero")
    if w > MAX_CANVAS_DIM or h > MAX_CANVAS_DIM:
        raise PfeError(f"Image size {w}x{h} exceeds maximum allowed {MAX_CANVAS_DIM}x{MAX_CANVAS_DIM}")
    active = r.u64()
    n_layers = r.u64()
    if n_layers > MAX_LAYERS:
        raise PfeError(f"Project contains {n_layers} layers, which exceeds the maximum of {MAX_LAYERS}")
    layers = []
    for _ in range(n_layers):
        name, visible, opacity, blend = r.string(), r.u8() != 0, r.f32(), r.u8()
        L = PfeLayer(name, visible, opacity, blend)
        if magic == "PFE0":  # flat w*h*4 buffer (io.rs:106-113)
            px = r.blob()
            if len(px) != w * h * 4:
                raise PfeError(f"Layer '{name}' has {len(px)} bytes, expected {w * h * 4}")
            flat = np.frombuffer(px, np.uint8).reshape(h, w, 4)
            for cy in range((h + CHUNK - 1) // CHUNK):
                for cx in range((w + CHUNK - 1) // CHUNK):
                    blk = flat[cy * CHUNK:(cy + 1) * CHUNK, cx * CHUNK:(cx + 1) * CHUNK]
                    if blk[..., 3].any():  # from_rgba_image keeps chunks with any alpha (tiled_image.rs:82-97)
                        t = np.zeros((CHUNK, CHUNK, 4), np.uint8)
                        t[:blk.shape[0], :blk.shape[1]] = blk
                        L.chunks[(cx, cy)] = t
        else:
            if magic == "PFE2":
                layer_type = r.u8()
            for _ in range(r.u64()):
                cx, cy = r.u32(), r.u32()
                px = r.blob()
                if len(px) != CHUNK_BYTES:
                    raise PfeError(f"Chunk ({cx},{cy}) in layer '{name}' has {len(px)} bytes, expected {CHUNK_BYTES}")
                L.chunks[(cx, cy)] = np.frombuffer(px, np.uint8).reshape(CHUNK, CHUNK, 4).copy()
            if magic == "PFE2":
                if r.u8():      # Option<Vec<u8>> text_data: the rasterised chunks above are what composites
                    r.blob()
                del layer_type
        layers.append(L)
    return PfeProject(w, h, active, layers)


def load_pfe(path: str) -> PfeProject:
    with open(path, "rb") as f:
        return load_pfe_from_bytes(f.read())


def save_pfe_v1(project: PfeProject) -> bytes:
    """build_pfe_v1 + write_pfe_v1 (io.rs:296-340); chunks in chunk_keys order (row-major)."""
    out = bytearray()

    def string(s: str):
        b = s.encode("utf-8")
        out.extend(struct.pack("<Q", len(b)))
        out.extend(b)

    string("PFE1")
    out.extend(struct.pack("<IIQ", project.width, project.height, project.active_layer_index))
    out.extend(struct.pack("<Q", len(project.layers)))
    for L in project.layers:
        string(L.name)
        out.extend(struct.pack("<BfB", 1 if L.visible else 0, L.opacity, L.blend_mode & 0xFF))
        keys = sorted(L.chunks, key=lambda k: (k[1], k[0]))
        out.extend(struct.pack("<Q", len(keys)))
        for (cx, cy) in keys:
            out.extend(struct.pack("<IIQ", cx, cy, CHUNK_BYTES))
            out.extend(np.ascontiguousarray(L.chunks[(cx, cy)], np.uint8).tobytes())
    return bytes(out)


def layer_from_flat(name: str, flat: np.ndarray, visible=True, opacity=1.0, blend_mode=0) -> PfeLayer:
    """TiledImage::from_rgba_image (tiled_image.rs:50-104): only chunks with some alpha are kept."""
    h, w = flat.shape[:2]
    L = PfeLayer(name, visible, float(opacity), int(blend_mode))
    for cy in range((h + CHUNK - 1) // CHUNK):
        for cx in range((w + CHUNK - 1) // CHUNK):
            blk = flat[cy * CHUNK:(cy + 1) * CHUNK, cx * CHUNK:(cx + 1) * CHUNK]
            if blk[..., 3].any():
                t = np.zeros((CHUNK, CHUNK, 4), np.uint8)
                t[:blk.shape[0], :blk.shape[1]] = blk
                L.chunks[(cx, cy)] = t
    return L
