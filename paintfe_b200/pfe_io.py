"""`.pfe` project files: reader for v0/v1/v2/v3 and writers for v1 and v3.

Reference: src/io.rs:85-208 (structures), :296-340 (v1 writer), :477-497 (magic dispatch),
:1110-1146 (v1 loader and its validation).  The files are bincode 1.3.3 with the default options:
little-endian fixed-width integers, `u64` length prefixes for String / Vec, `usize` as u64, `bool`
as one byte, `Option` as a one-byte tag.  The magic string "PFE<n>" therefore sits at bytes 8..12.

This is harness plumbing on either side of the hot path: it lets the CLI run
`--flatten` on a multi-layer project (BASELINE config 1).  v3 (io.rs:171-208, :405-460, :812-900) adds
layer folders, adjustment layers (`content_data` = bincode of AdjustmentLayerData, layers.rs:244-273) and
per-layer format / HDR / source metadata, which is parsed and carried but does not affect compositing.
The reference has no .pfe fixtures (its tests round-trip through its own writer), so the v3 layout here
follows the serde derive order of the structs and is checked against this module's own writer only.
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
    # v3 only
    folder_id: object = None          # Option<u64>
    layer_type: int = 0               # 0 Raster, 1 Text, 2 Adjustment
    adjustment: object = None         # (kind, params): kind 1 Exposure, 2 BrightnessContrast, 3 Invert, 4 ChannelMixer
    extra: object = None              # raw (pixel_format, hdr, source metadata, webp, deep pixels) for a faithful rewrite

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
    folders: list = field(default_factory=list)  # v3: dicts {id, name, visible, collapsed, insert_above_layer, color_index}
    next_layer_folder_id: int = 1

    def layer_effectively_visible(self, i: int) -> bool:
        """canvas_state.rs:216-227: a layer in a hidden folder does not composite."""
        L = self.layers[i]
        if not L.visible:
            return False
        if L.folder_id is None:
            return True
        for f in self.folders:
            if f["id"] == L.folder_id:
                return bool(f["visible"])
        return True


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
    def opt(self, fn): return fn() if self.u8() else None


def load_pfe_from_bytes(raw: bytes) -> PfeProject:
    """io.rs:477-497 + the per-version loaders."""
    if len(raw) < 12:
        raise PfeError("File too small")
    magic = raw[8:12].decode("utf-8", "replace")
    if magic == "PFE3":
        return _load_v3(raw)
    if magic not in ("PFE0", "PFE1", "PFE2"):
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


# AdjustmentKind (layers.rs:244-262): bincode writes the variant index as u32, then the fields in order
_ADJ_FIELDS = {0: 1, 1: 2, 2: 0, 3: 16}  # Exposure{ev}, BrightnessContrast{b, c}, Invert, ChannelMixer{4 x [f32; 4]}


def _read_adjustment(blob: bytes):
    r = _Reader(blob)
    variant = r.u32()
    if variant not in _ADJ_FIELDS:
        raise PfeError(f"unknown AdjustmentKind variant {variant}")
    return variant + 1, tuple(r.f32() for _ in range(_ADJ_FIELDS[variant]))  # kind ids of pfe_layer_kind


def _write_adjustment(adj) -> bytes:
    kind, params = adj
    return struct.pack("<I", kind - 1) + b"".join(struct.pack("<f", float(v)) for v in params)


def _load_v3(raw: bytes) -> PfeProject:
    """load_pfe_v3, io.rs:812-900."""
    r = _Reader(raw)
    r.string()
    w, h = r.u32(), r.u32()
    if w == 0 or h == 0:
        raise PfeError("Image dimensions cannot be zero")
    if w > MAX_CANVAS_DIM or h > MAX_CANVAS_DIM:
        raise PfeError(f"Image size {w}x{h} exceeds maximum allowed {MAX_CANVAS_DIM}x{MAX_CANVAS_DIM}")
    active = r.u64()
    folders = [dict(id=r.u64(), name=r.string(), visible=r.u8() != 0, collapsed=r.u8() != 0,
                    insert_above_layer=r.opt(r.u64), color_index=r.opt(r.u8)) for _ in range(r.u64())]
    next_folder = r.u64()
    n_layers = r.u64()
    if n_layers > MAX_LAYERS:
        raise PfeError(f"Project contains {n_layers} layers, which exceeds the maximum of {MAX_LAYERS}")
    layers = []
    for _ in range(n_layers):
        name, visible = r.string(), r.u8() != 0
        folder_id = r.opt(r.u64)
        opacity, blend, layer_type = r.f32(), r.u8(), r.u8()
        L = PfeLayer(name, visible, opacity, blend, folder_id=folder_id, layer_type=layer_type)
        for _ in range(r.u64()):
            cx, cy = r.u32(), r.u32()
            px = r.blob()
            if len(px) != CHUNK_BYTES:
                raise PfeError(f"Chunk ({cx},{cy}) in layer '{name}' has {len(px)} bytes, expected {CHUNK_BYTES}")
            L.chunks[(cx, cy)] = np.frombuffer(px, np.uint8).reshape(CHUNK, CHUNK, 4).copy()
        content = r.opt(r.blob)
        if layer_type == 2 and content is not None:
            try:
                L.adjustment = _read_adjustment(content)
            except PfeError:
                L.adjustment = None  # `.ok()` -> falls back to LayerContent::Raster (io.rs:868-874)
        start = r.o
        r.u32()                                                        # pixel_format
        r.u8(); r.opt(r.f32); r.opt(r.f32); r.opt(r.string)            # hdr_metadata
        r.opt(r.string); r.opt(r.string); r.opt(r.string)              # source_metadata
        for _ in range(r.u64()):
            r.string(); r.string()
        for _ in range(r.u64()):
            r.blob()
        r.u32()                                                        # webp_frame_compression
        if r.u8():                                                     # deep_pixels: Option<DeepRgbaBuffer>
            variant = r.u32()
            r.take(r.u64() * {0: 1, 1: 2, 2: 2, 3: 4}.get(variant, 1))
        L.extra = (content if layer_type != 2 else None, raw[start:r.o])
        layers.append(L)
    if not layers:
        raise PfeError("Project contains no layers")
    return PfeProject(w, h, min(active, len(layers) - 1), layers, folders, next_folder)


def save_pfe_v3(project: PfeProject) -> bytes:
    """build_pfe_v3 + write_pfe_v3 (io.rs:405-475)."""
    out = bytearray()

    def string(s: str):
        b = s.encode("utf-8")
        out.extend(struct.pack("<Q", len(b)))
        out.extend(b)

    def opt(v, fmt):
        out.extend(b"\x00" if v is None else b"\x01" + struct.pack(fmt, v))

    string("PFE3")
    out.extend(struct.pack("<IIQ", project.width, project.height, project.active_layer_index))
    out.extend(struct.pack("<Q", len(project.folders)))
    for f in project.folders:
        out.extend(struct.pack("<Q", f["id"]))
        string(f["name"])
        out.extend(struct.pack("<BB", 1 if f["visible"] else 0, 1 if f.get("collapsed") else 0))
        opt(f.get("insert_above_layer"), "<Q")
        opt(f.get("color_index"), "<B")
    out.extend(struct.pack("<Q", project.next_layer_folder_id))
    out.extend(struct.pack("<Q", len(project.layers)))
    # default tail: RgbaU8, HdrMetadata::default, ImageMetadata::default, Lossless, no deep pixels
    default_tail = struct.pack("<I", 0) + b"\x00" * 4 + b"\x00" * 3 + struct.pack("<QQ", 0, 0) + struct.pack("<I", 1) + b"\x00"
    for L in project.layers:
        string(L.name)
        out.extend(b"\x01" if L.visible else b"\x00")
        opt(L.folder_id, "<Q")
        layer_type = 2 if L.adjustment is not None else L.layer_type
        out.extend(struct.pack("<fBB", L.opacity, L.blend_mode & 0xFF, layer_type))
        keys = sorted(L.chunks, key=lambda k: (k[1], k[0]))
        out.extend(struct.pack("<Q", len(keys)))
        for (cx, cy) in keys:
            out.extend(struct.pack("<IIQ", cx, cy, CHUNK_BYTES))
            out.extend(np.ascontiguousarray(L.chunks[(cx, cy)], np.uint8).tobytes())
        content = _write_adjustment(L.adjustment) if L.adjustment is not None else (L.extra[0] if L.extra else None)
        if content is None:
            out.extend(b"\x00")
        else:
            out.extend(b"\x01" + struct.pack("<Q", len(content)) + content)
        out.extend(L.extra[1] if L.extra else default_tail)
    return bytes(out)


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
