"""Thin Python handle on a pfe_ctx: numpy arrays go through the host tier (pfe_*), torch CUDA
tensors through the device tier (pfe_dev_*) on torch's current stream.  No compute happens in
Python and nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import _lib as L

try:  # torch is plumbing only (device memory + streams for the device tier)
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_tensor(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _np_u8(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


def _ptr(a):
    if a is None:
        return None
    if _is_tensor(a):
        if not a.is_contiguous():
            raise ValueError("device tensors must be contiguous")
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


class Layer(dict):
    """Marshalled form of one `Layer` for flatten(): keys rgba, mask, opacity, blend, visible, kind, adj."""


def make_layer(rgba=None, opacity=1.0, blend=0, visible=True, mask=None, kind=0, adj=()):
    return Layer(rgba=rgba, mask=mask, opacity=float(opacity), blend=int(blend), visible=bool(visible),
                 kind=int(kind), adj=tuple(float(v) for v in adj))


class DeviceTiled:
    """A TiledImage resident on the device (`pfe_tiled`): chunk pool, pointer table, occupancy."""

    def __init__(self, eng, w, h, _handle=None):
        self.eng, self.w, self.h = eng, int(w), int(h)
        hnd = _handle
        if hnd is None:
            hnd = C.c_void_p()
            eng._ck(eng.lib.pfe_tiled_create(eng.h, self.w, self.h, C.byref(hnd)))
        self.hnd = hnd
        self.n_chunks = ((self.w + 63) // 64) * ((self.h + 63) // 64)

    def clone(self):
        """A snapshot sharing every chunk (copy on write): pfe_tiled_clone."""
        self.eng.use_torch_stream()
        hnd = C.c_void_p()
        self.eng._ck(self.eng.lib.pfe_tiled_clone(self.eng.h, self.hnd, C.byref(hnd)))
        return DeviceTiled(self.eng, self.w, self.h, _handle=hnd)

    def make_mut(self, chunk_indices):
        """ensure_chunk_mut for the listed chunks: populated, and private to this image afterwards."""
        idx = np.ascontiguousarray(np.asarray(chunk_indices, np.uint32).reshape(-1))
        self.eng.use_torch_stream()
        self.eng._ck(self.eng.lib.pfe_tiled_make_mut(self.eng.h, self.hnd, _ptr(idx), len(idx)))
        return self

    def chunk_ids(self):
        ids = np.empty(self.n_chunks, np.int32)
        self.eng._ck(self.eng.lib.pfe_tiled_chunk_ids(self.eng.h, self.hnd, _ptr(ids)))
        return ids

    @property
    def table(self):
        return self.eng.lib.pfe_tiled_table(self.hnd)

    def upload(self, table):
        """table: row-major chunk grid of (64,64,4) uint8 arrays or None."""
        arr, keep = Engine._chunk_table(table)
        self.eng._ck(self.eng.lib.pfe_tiled_upload(self.eng.h, self.hnd, arr))
        return self

    def from_flat(self, flat_dev):
        self.eng.use_torch_stream()
        self.eng._ck(self.eng.lib.pfe_tiled_from_flat(self.eng.h, self.hnd, _ptr(flat_dev)))
        return self

    def to_flat(self, out=None):
        self.eng.use_torch_stream()
        out = out if out is not None else torch.empty((self.h, self.w, 4), dtype=torch.uint8, device=f"cuda:{self.eng.device}")
        self.eng._ck(self.eng.lib.pfe_tiled_to_flat(self.eng.h, self.hnd, _ptr(out)))
        return out

    def download(self, want_tiles=True):
        occ = np.empty(self.n_chunks, np.uint8)
        tiles = np.zeros((self.n_chunks, 64, 64, 4), np.uint8) if want_tiles else None
        self.eng.use_torch_stream()
        self.eng._ck(self.eng.lib.pfe_tiled_download(self.eng.h, self.hnd, _ptr(occ), _ptr(tiles)))
        return occ, tiles

    def close(self):
        if getattr(self, "hnd", None) and getattr(self.eng, "h", None):
            self.eng.lib.pfe_tiled_destroy(self.eng.h, self.hnd)
        self.hnd = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One CUDA device, one stream, scratch buffers (`pfe_ctx`)."""

    def __init__(self, device: Optional[int] = None):
        self.lib = L.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = device
        h = C.c_void_p()
        rc = self.lib.pfe_ctx_create(device, C.byref(h))
        if rc != L.PFE_OK:
            raise L.PfeError(rc, "pfe_ctx_create (the pixel engine needs a CUDA device; there is no CPU path)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.pfe_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != L.PFE_OK:
            raise L.PfeError(rc, (self.lib.pfe_last_error(self.h) or b"").decode())

    def sync(self):
        self._ck(self.lib.pfe_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.lib.pfe_ctx_launch_count(self.h))

    def profile(self, enable: bool):
        """Bracket every kernel launch with CUDA events (see pfe_ctx_profile)."""
        self._ck(self.lib.pfe_ctx_profile(self.h, 1 if enable else 0))

    def profile_read(self) -> dict:
        import json
        buf = C.create_string_buffer(1 << 16)
        self._ck(self.lib.pfe_ctx_profile_read(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def use_torch_stream(self):
        """Enqueue device-tier work on torch's current stream for this device."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._ck(self.lib.pfe_ctx_set_stream(self.h, C.c_void_p(s)))

    def _dev(self, *xs) -> bool:
        """True when the call should use the device tier (any torch tensor among the images)."""
        d = any(_is_tensor(x) for x in xs if x is not None)
        if d:
            for x in xs:
                if x is not None and not (_is_tensor(x) and x.is_cuda):
                    raise ValueError("device-tier call: every image must be a CUDA tensor")
            self.use_torch_stream()
        return d

    def _out_like(self, src, shape=None, dtype=None):
        if _is_tensor(src):
            return torch.empty(shape or tuple(src.shape), dtype=dtype or src.dtype, device=src.device)
        return np.empty(shape or src.shape, dtype or src.dtype)

    @staticmethod
    def _hw(img):
        if img.ndim != 3 or img.shape[2] != 4:
            raise ValueError("images are (h, w, 4) uint8")
        return int(img.shape[0]), int(img.shape[1])

    def _prep(self, img):
        return img if _is_tensor(img) else _np_u8(img)

    def _prep_mask(self, mask, h, w):
        if mask is None:
            return None
        if _is_tensor(mask):
            if tuple(mask.shape) != (h, w):
                raise ValueError("mask must be (h, w)")
            return mask
        return _np_u8(mask, (h, w))

    # -- flatten ------------------------------------------------------------------------------
    def _layer_array(self, layers: Sequence[Layer], h, w):
        arr = (L.LayerDesc * max(len(layers), 1))()
        keep = []
        dev = None
        for i, Ly in enumerate(layers):
            rgba, mask = Ly.get("rgba"), Ly.get("mask")
            if rgba is not None:
                rgba = self._prep(rgba)
                if tuple(rgba.shape) != (h, w, 4):
                    raise ValueError(f"layer {i}: expected {(h, w, 4)}, got {tuple(rgba.shape)}")
                dev = _is_tensor(rgba) if dev is None else dev
                if _is_tensor(rgba) != dev:
                    raise ValueError("all layers must be on the same side (numpy or CUDA tensors)")
                arr[i].rgba = _ptr(rgba).value
                keep.append(rgba)
            if mask is not None:
                mask = self._prep_mask(mask, h, w)
                arr[i].mask = _ptr(mask).value
                keep.append(mask)
            arr[i].opacity = Ly.get("opacity", 1.0)
            arr[i].blend = Ly.get("blend", 0) & 0xFF
            arr[i].visible = 1 if Ly.get("visible", True) else 0
            arr[i].kind = Ly.get("kind", 0)
            for j, v in enumerate(Ly.get("adj", ())):
                arr[i].adj[j] = v
        return arr, keep, bool(dev)

    def flatten(self, layers: Sequence[Layer], w: int, h: int, active=None, out=None):
        """CanvasState::composite over marshalled layers (canvas_state.rs:482)."""
        arr, keep, dev = self._layer_array(layers, h, w)
        if dev:
            self.use_torch_stream()
            dst = out if out is not None else torch.empty((h, w, 4), dtype=torch.uint8, device=keep[0].device)
            self._ck(self.lib.pfe_dev_flatten(self.h, arr, len(layers), w, h, _ptr(active), _ptr(dst)))
        else:
            dst = out if out is not None else np.empty((h, w, 4), np.uint8)
            active = None if active is None else _np_u8(active)
            self._ck(self.lib.pfe_flatten(self.h, arr, len(layers), w, h, _ptr(active), _ptr(dst)))
        return dst

    def prepare_layers(self, layers: Sequence[Layer], w: int, h: int):
        """Marshal a layer list once (the pfe_layer_desc array) for repeated `flatten_prepared` calls: a caller that
        flattens the same device-resident stack every step then pays no per-call Python marshalling."""
        arr, keep, dev = self._layer_array(layers, h, w)
        if not dev and any(Ly.get("rgba") is not None for Ly in layers):
            raise ValueError("prepare_layers is for device-resident layers")
        return (arr, keep, len(layers), int(w), int(h))

    def flatten_prepared(self, prepared, out, active=None):
        arr, _keep, n, w, h = prepared
        self.use_torch_stream()
        self._ck(self.lib.pfe_dev_flatten(self.h, arr, n, w, h, _ptr(active), _ptr(out)))
        return out

    def flatten_prepared_peer(self, prepared, out, peer_dst: int, peer_flag: int = 0, flag_value: int = 0, active=None):
        """pfe_dev_flatten_peer: flatten into `out` and, with the same stores, into the raw device address `peer_dst`
        (normally a neighbour GPU's halo rows, see `peer_open`); then release `flag_value` at address `peer_flag`."""
        arr, _keep, n, w, h = prepared
        self.use_torch_stream()
        self._ck(self.lib.pfe_dev_flatten_peer(self.h, arr, n, w, h, _ptr(active), _ptr(out), C.c_void_p(peer_dst),
                                               C.c_void_p(peer_flag) if peer_flag else None, flag_value & 0xFFFFFFFF))
        return out

    def peer_signal(self, flag_addr: int, value: int):
        """Release `value` at device address `flag_addr` (system scope) after everything enqueued on the stream so far."""
        self.use_torch_stream()
        self._ck(self.lib.pfe_dev_peer_signal(self.h, C.c_void_p(flag_addr), value & 0xFFFFFFFF))

    def peer_wait(self, flags_addr: int, n: int, value: int, timeout_ms: int = 2000):
        """Stream-ordered wait until the `n` u32 flags at device address `flags_addr` have all reached `value`."""
        self.use_torch_stream()
        self._ck(self.lib.pfe_dev_peer_wait(self.h, C.c_void_p(flags_addr), n, value & 0xFFFFFFFF, timeout_ms))

    def peer_alloc(self, nbytes: int):
        """(device address, 64-byte handle) of zero-filled device memory other processes of this node can map."""
        p = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        self._ck(self.lib.pfe_peer_alloc(self.h, nbytes, C.byref(p), handle))
        return int(p.value), bytes(handle)

    def peer_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        self._ck(self.lib.pfe_peer_open(self.h, buf, C.byref(p)))
        return int(p.value)

    def peer_close(self, addr: int):
        self._ck(self.lib.pfe_peer_close(self.h, C.c_void_p(addr)))

    def peer_free(self, addr: int):
        self._ck(self.lib.pfe_peer_free(self.h, C.c_void_p(addr)))

    def flatten_gaussian(self, layers, w, h, sigma, active=None, exact=False, out=None):
        """Host tier only: composite() then parallel_gaussian_blur without leaving the device."""
        arr, keep, dev = self._layer_array(layers, h, w)
        if dev:
            raise ValueError("flatten_gaussian is the host-tier fused call; chain flatten + gaussian_blur on device")
        dst = out if out is not None else np.empty((h, w, 4), np.uint8)
        active = None if active is None else _np_u8(active)
        self._ck(self.lib.pfe_flatten_gaussian(self.h, arr, len(layers), w, h, _ptr(active), C.c_float(sigma),
                                               _ptr(dst), L.GAUSS_EXACT if exact else 0))
        return dst

    # -- tile-native flatten / device-resident TiledImage (tiles.cu) --------------------------------
    @staticmethod
    def _chunk_table(table):
        """list (row-major chunk grid) of (64,64,4) uint8 arrays or None -> ctypes pointer array."""
        arr = (C.c_void_p * max(len(table), 1))()
        keep = []
        for i, t in enumerate(table):
            if t is not None:
                t = _np_u8(t, (64, 64, 4))
                keep.append(t)
                arr[i] = t.ctypes.data
        return arr, keep

    def flatten_tiles(self, layers, w, h, out=None):
        """CanvasState::composite on chunk tables. Each layer is a dict like make_layer()'s with `tiles`
        (and optionally `mask_tiles`) = a chunk table (host tier: list of arrays / None) or a DeviceTiled."""
        arr = (L.TileLayerDesc * max(len(layers), 1))()
        keep = []
        dev = None
        for i, Ly in enumerate(layers):
            for key, field in (("tiles", "chunks"), ("mask_tiles", "mask_chunks")):
                t = Ly.get(key)
                if t is None:
                    continue
                is_dev = isinstance(t, DeviceTiled)
                dev = is_dev if dev is None else dev
                if is_dev != dev:
                    raise ValueError("all layers must be on the same side (host chunk tables or DeviceTiled)")
                if is_dev:
                    setattr(arr[i], field, t.table)
                else:
                    a, k = self._chunk_table(t)
                    keep += [a, k]
                    setattr(arr[i], field, C.cast(a, C.c_void_p).value)
            arr[i].opacity = Ly.get("opacity", 1.0)
            arr[i].blend = Ly.get("blend", 0) & 0xFF
            arr[i].visible = 1 if Ly.get("visible", True) else 0
            arr[i].kind = Ly.get("kind", 0)
            for j, v in enumerate(Ly.get("adj", ())):
                arr[i].adj[j] = v
        if dev:
            self.use_torch_stream()
            dst = out if out is not None else torch.empty((h, w, 4), dtype=torch.uint8, device=f"cuda:{self.device}")
            self._ck(self.lib.pfe_dev_flatten_tiles(self.h, arr, len(layers), w, h, _ptr(dst)))
        else:
            dst = out if out is not None else np.empty((h, w, 4), np.uint8)
            self._ck(self.lib.pfe_flatten_tiles(self.h, arr, len(layers), w, h, _ptr(dst)))
        return dst

    def tiled(self, w, h):
        return DeviceTiled(self, w, h)

    # -- single-image ops -----------------------------------------------------------------------
    def _img_op(self, host_fn, dev_fn, src, mask, out, *mid, tail=()):
        src = self._prep(src)
        h, w = self._hw(src)
        mask = self._prep_mask(mask, h, w)
        dst = out if out is not None else self._out_like(src)
        fn = dev_fn if self._dev(src, mask, dst) else host_fn
        self._ck(fn(self.h, _ptr(src), w, h, *mid, _ptr(mask), _ptr(dst), *tail))
        return dst

    def gaussian_blur(self, src, sigma, mask=None, exact=False, out=None):
        return self._img_op(self.lib.pfe_gaussian_blur, self.lib.pfe_dev_gaussian_blur, src, mask, out,
                            C.c_float(sigma), tail=(L.GAUSS_EXACT if exact else 0,))

    def gaussian_band_h(self, ext, y0, rows, sigma, exact=False):
        """Device tier: H pass of rows [y0, y0+rows) of an extended band (pfe_dev_gaussian_band_h)."""
        self._dev(ext)
        self._ck(self.lib.pfe_dev_gaussian_band_h(self.h, _ptr(ext), int(ext.shape[1]), int(ext.shape[0]), int(y0), int(rows),
                                                  C.c_float(sigma), L.GAUSS_EXACT if exact else 0))

    def gaussian_band_v(self, ext, y0, rows, sigma, exact=False, out=None):
        """Device tier: V pass producing rows [y0, y0+rows) of the extended band `ext` (only its shape is used)."""
        w, ext_rows = int(ext.shape[1]), int(ext.shape[0])
        dst = out if out is not None else torch.empty((rows, w, 4), dtype=torch.uint8, device=ext.device)
        self._dev(ext, dst)
        self._ck(self.lib.pfe_dev_gaussian_band_v(self.h, w, ext_rows, int(y0), int(rows), C.c_float(sigma), _ptr(dst),
                                                  L.GAUSS_EXACT if exact else 0))
        return dst

    def check_async(self):
        """Synchronise and raise if a stream-asynchronous call recorded an error on the device (pfe_ctx_check_async)."""
        self._ck(self.lib.pfe_ctx_check_async(self.h))

    def box_blur(self, src, radius, mask=None, out=None):
        return self._img_op(self.lib.pfe_box_blur, self.lib.pfe_dev_box_blur, src, mask, out, C.c_float(radius))

    def motion_blur(self, src, angle_deg, distance, mask=None, out=None):
        return self._img_op(self.lib.pfe_motion_blur, self.lib.pfe_dev_motion_blur, src, mask, out,
                            C.c_float(angle_deg), C.c_float(distance))

    def median(self, src, radius, mask=None, out=None):
        return self._img_op(self.lib.pfe_median, self.lib.pfe_dev_median, src, mask, out, C.c_uint32(radius))

    def sharpen(self, src, amount, radius, mask=None, exact=False, out=None):
        return self._img_op(self.lib.pfe_sharpen, self.lib.pfe_dev_sharpen, src, mask, out, C.c_float(amount),
                            C.c_float(radius), tail=(L.GAUSS_EXACT if exact else 0,))

    def vignette(self, src, amount, softness, mask=None, out=None):
        return self._img_op(self.lib.pfe_vignette, self.lib.pfe_dev_vignette, src, mask, out, C.c_float(amount),
                            C.c_float(softness))

    def glow(self, src, radius, intensity, mask=None, exact=False, out=None):
        return self._img_op(self.lib.pfe_glow, self.lib.pfe_dev_glow, src, mask, out, C.c_float(radius),
                            C.c_float(intensity), tail=(L.GAUSS_EXACT if exact else 0,))

    def pixelate(self, src, block_size, mask=None, out=None):
        return self._img_op(self.lib.pfe_pixelate, self.lib.pfe_dev_pixelate, src, mask, out, C.c_uint32(block_size))

    def bulge(self, src, amount, origin=(0.5, 0.5), mask=None, out=None):
        return self._img_op(self.lib.pfe_bulge, self.lib.pfe_dev_bulge, src, mask, out, C.c_float(amount),
                            C.c_float(origin[0]), C.c_float(origin[1]))

    def twist(self, src, angle_deg, origin=(0.5, 0.5), mask=None, out=None):
        return self._img_op(self.lib.pfe_twist, self.lib.pfe_dev_twist, src, mask, out, C.c_float(angle_deg),
                            C.c_float(origin[0]), C.c_float(origin[1]))

    def add_noise(self, src, amount, noise_type, monochrome, seed, scale, octaves, mask=None, out=None):
        return self._img_op(self.lib.pfe_add_noise, self.lib.pfe_dev_add_noise, src, mask, out, C.c_float(amount),
                            int(noise_type), 1 if monochrome else 0, C.c_uint32(seed), C.c_float(scale), C.c_uint32(octaves))

    def reduce_noise(self, src, strength, radius, mask=None, out=None):
        return self._img_op(self.lib.pfe_reduce_noise, self.lib.pfe_dev_reduce_noise, src, mask, out,
                            C.c_float(strength), C.c_uint32(radius))

    # -- the rest of src/ops/effects/ (effects3.cu) ---------------------------------------------------
    @staticmethod
    def _rgba(color):
        return (C.c_uint8 * 4)(*[int(v) for v in color])

    def ink(self, src, edge_strength, threshold, mask=None, out=None):
        return self._img_op(self.lib.pfe_ink, self.lib.pfe_dev_ink, src, mask, out, C.c_float(edge_strength), C.c_float(threshold))

    def oil_painting(self, src, radius, levels, mask=None, out=None):
        return self._img_op(self.lib.pfe_oil_painting, self.lib.pfe_dev_oil_painting, src, mask, out,
                            C.c_uint32(radius), C.c_uint32(levels))

    def color_filter(self, src, color, intensity, mode, mask=None, out=None):
        return self._img_op(self.lib.pfe_color_filter, self.lib.pfe_dev_color_filter, src, mask, out, self._rgba(color),
                            C.c_float(intensity), int(mode))

    def contours(self, src, scale, frequency, line_width, color, seed, octaves, blend, mask=None, out=None):
        return self._img_op(self.lib.pfe_contours, self.lib.pfe_dev_contours, src, mask, out, C.c_float(scale),
                            C.c_float(frequency), C.c_float(line_width), self._rgba(color), C.c_uint32(seed),
                            C.c_uint32(octaves), C.c_float(blend))

    def crystallize(self, src, cell_size, seed, mask=None, out=None):
        return self._img_op(self.lib.pfe_crystallize, self.lib.pfe_dev_crystallize, src, mask, out, C.c_float(cell_size),
                            C.c_uint32(seed))

    def dents(self, src, scale, amount, seed, octaves, roughness, pinch, wrap, mask=None, out=None):
        return self._img_op(self.lib.pfe_dents, self.lib.pfe_dev_dents, src, mask, out, C.c_float(scale), C.c_float(amount),
                            C.c_uint32(seed), C.c_uint32(octaves), C.c_float(roughness), 1 if pinch else 0, 1 if wrap else 0)

    def halftone(self, src, dot_size, angle_deg, shape, mask=None, out=None):
        return self._img_op(self.lib.pfe_halftone, self.lib.pfe_dev_halftone, src, mask, out, C.c_float(dot_size),
                            C.c_float(angle_deg), int(shape))

    def bokeh_blur(self, src, radius, mask=None, out=None):
        return self._img_op(self.lib.pfe_bokeh_blur, self.lib.pfe_dev_bokeh_blur, src, mask, out, C.c_float(radius))

    def zoom_blur(self, src, center_x, center_y, strength, samples, tint=(0.0, 0.0, 0.0, 0.0), tint_strength=0.0,
                  mask=None, out=None):
        t = (C.c_float * 4)(*tint)
        return self._img_op(self.lib.pfe_zoom_blur, self.lib.pfe_dev_zoom_blur, src, mask, out, C.c_float(center_x),
                            C.c_float(center_y), C.c_float(strength), C.c_uint32(samples), t, C.c_float(tint_strength))

    def grid(self, src, cell_w, cell_h, line_width, color, style, opacity, mask=None, out=None):
        return self._img_op(self.lib.pfe_grid, self.lib.pfe_dev_grid, src, mask, out, C.c_uint32(cell_w), C.c_uint32(cell_h),
                            C.c_uint32(line_width), self._rgba(color), int(style), C.c_float(opacity))

    def canvas_border(self, src, width, color, mask=None, out=None):
        return self._img_op(self.lib.pfe_canvas_border, self.lib.pfe_dev_canvas_border, src, mask, out, C.c_uint32(width),
                            self._rgba(color))

    def drop_shadow(self, src, offset_x, offset_y, blur_radius, widen_radius, color, opacity, mask=None, exact=False, out=None):
        return self._img_op(self.lib.pfe_drop_shadow, self.lib.pfe_dev_drop_shadow, src, mask, out, C.c_int32(offset_x),
                            C.c_int32(offset_y), C.c_float(blur_radius), 1 if widen_radius else 0, self._rgba(color),
                            C.c_float(opacity), tail=(L.GAUSS_EXACT if exact else 0,))

    def outline(self, src, width, color, mode, anti_alias, mask=None, out=None):
        return self._img_op(self.lib.pfe_outline, self.lib.pfe_dev_outline, src, mask, out, C.c_uint32(width),
                            self._rgba(color), int(mode), 1 if anti_alias else 0)

    def pixel_drag(self, src, seed, amount, distance, direction, mask=None, out=None):
        return self._img_op(self.lib.pfe_pixel_drag, self.lib.pfe_dev_pixel_drag, src, mask, out, C.c_uint32(seed),
                            C.c_float(amount), C.c_uint32(distance), C.c_float(direction))

    def rgb_displace(self, src, r_off, g_off, b_off, mask=None, out=None):
        off = (C.c_int32 * 6)(r_off[0], r_off[1], g_off[0], g_off[1], b_off[0], b_off[1])
        return self._img_op(self.lib.pfe_rgb_displace, self.lib.pfe_dev_rgb_displace, src, mask, out, off)

    # -- geometry (geometry.cu): source and result shapes may differ --------------------------------
    def _reshape_op(self, host_fn, dev_fn, src, out_hw, out, *args):
        src = self._prep(src)
        dev = _is_tensor(src)
        if out is None:
            out = torch.empty((out_hw[0], out_hw[1], 4), dtype=torch.uint8, device=src.device) if dev \
                else np.empty((out_hw[0], out_hw[1], 4), np.uint8)
        self._ck((dev_fn if dev else host_fn)(self.h, _ptr(src), *args, _ptr(out)))
        return out

    def orient(self, src, op, out=None):
        h, w = self._hw(src)
        return self._reshape_op(self.lib.pfe_orient, self.lib.pfe_dev_orient, src, (w, h) if op in (2, 3) else (h, w), out,
                                w, h, int(op))

    def resize_canvas(self, src, new_w, new_h, anchor, fill, out=None):
        h, w = self._hw(src)
        return self._reshape_op(self.lib.pfe_resize_canvas, self.lib.pfe_dev_resize_canvas, src, (new_h, new_w), out,
                                w, h, new_w, new_h, anchor[0], anchor[1], self._rgba(fill))

    def affine(self, src, canvas_w, canvas_h, rotation_z, rotation_x=0.0, rotation_y=0.0, scale=1.0, offset=(0.0, 0.0),
               nearest=False, out=None):
        h, w = self._hw(src)
        return self._reshape_op(self.lib.pfe_affine, self.lib.pfe_dev_affine, src, (canvas_h, canvas_w), out, w, h,
                                canvas_w, canvas_h, C.c_float(rotation_z), C.c_float(rotation_x), C.c_float(rotation_y),
                                C.c_float(scale), C.c_float(offset[0]), C.c_float(offset[1]), 1 if nearest else 0)

    def resize(self, src, new_w, new_h, filter, out=None):
        h, w = self._hw(src)
        return self._reshape_op(self.lib.pfe_resize, self.lib.pfe_dev_resize, src, (new_h, new_w), out, w, h, new_w, new_h,
                                int(filter))

    def adjust(self, src, op, params=(), luts=None, mask=None, occupancy=None, out=None):
        src = self._prep(src)
        h, w = self._hw(src)
        mask = self._prep_mask(mask, h, w)
        d = L.AdjustDesc()
        d.op = int(op)
        for i, v in enumerate(params):
            d.params[i] = float(v)
        luts_np = None
        if luts is not None:
            luts_np = _np_u8(np.asarray(luts).reshape(-1))
            d.luts = luts_np.ctypes.data
        dst = out if out is not None else self._out_like(src)
        if self._dev(src, mask, dst):
            if occupancy is not None and not _is_tensor(occupancy):
                occupancy = torch.as_tensor(np.ascontiguousarray(occupancy, dtype=np.uint8), device=src.device)
            self._ck(self.lib.pfe_dev_adjust(self.h, _ptr(src), w, h, C.byref(d), _ptr(mask), _ptr(occupancy), _ptr(dst)))
        else:
            occupancy = None if occupancy is None else _np_u8(occupancy)
            self._ck(self.lib.pfe_adjust(self.h, _ptr(src), w, h, C.byref(d), _ptr(mask), _ptr(occupancy), _ptr(dst)))
        return dst

    def channel_minmax(self, src, mask=None):
        src = self._prep(src)
        h, w = self._hw(src)
        mask = self._prep_mask(mask, h, w)
        out = np.zeros(6, np.uint8)
        fn = self.lib.pfe_dev_channel_minmax if self._dev(src, mask) else self.lib.pfe_channel_minmax
        self._ck(fn(self.h, _ptr(src), w, h, _ptr(mask), _ptr(out)))
        return out

    # -- LUT builders (host arithmetic inside the library) -----------------------------------
    def levels_lut(self, in_black, in_white, gamma, out_black=0.0, out_white=255.0):
        lut = np.empty(256, np.uint8)
        self.lib.pfe_build_levels_lut(in_black, in_white, gamma, out_black, out_white, _ptr(lut))
        return lut

    def levels_lut_script(self, in_black, in_white, gamma):
        lut = np.empty(256, np.uint8)
        self.lib.pfe_build_levels_lut_script(in_black, in_white, gamma, _ptr(lut))
        return lut

    def stretch_lut(self, mn, mx):
        lut = np.empty(256, np.uint8)
        self.lib.pfe_build_stretch_lut(int(mn), int(mx), _ptr(lut))
        return lut

    def curves_lut(self, points):
        pts = np.ascontiguousarray(np.asarray(points, np.float32).reshape(-1, 2))
        lut = np.empty(256, np.uint8)
        self.lib.pfe_build_curves_lut(_ptr(pts), len(pts), _ptr(lut))
        return lut

    def compose_curve_luts(self, five):
        five = _np_u8(np.asarray(five).reshape(5 * 256))
        out = np.empty(4 * 256, np.uint8)
        self.lib.pfe_compose_curve_luts(_ptr(five), _ptr(out))
        return out.reshape(4, 256)

    # -- warps --------------------------------------------------------------------------------
    def warp_displacement(self, src, disp, out=None):
        src = self._prep(src)
        sh, sw = self._hw(src)
        if not _is_tensor(disp):
            disp = np.ascontiguousarray(disp, dtype=np.float32)
        h, w = int(disp.shape[0]), int(disp.shape[1])
        dst = out if out is not None else self._out_like(src, (h, w, 4))
        fn = self.lib.pfe_dev_warp_displacement if self._dev(src, disp, dst) else self.lib.pfe_warp_displacement
        self._ck(fn(self.h, _ptr(src), sw, sh, _ptr(disp), w, h, _ptr(dst)))
        return dst

    def warp_displacement_region(self, src, disp, prev, dirty_rect, out=None):
        """warp_displacement_region (transform.rs:1206-1285): `prev` outside dirty_rect = (x0, y0, x1, y1), the warp inside."""
        src, prev = self._prep(src), self._prep(prev)
        sh, sw = self._hw(src)
        if not _is_tensor(disp):
            disp = np.ascontiguousarray(disp, dtype=np.float32)
        h, w = int(disp.shape[0]), int(disp.shape[1])
        rect = (C.c_int32 * 4)(*[int(v) for v in dirty_rect])
        dst = out if out is not None else self._out_like(prev)
        fn = self.lib.pfe_dev_warp_displacement_region if self._dev(src, disp, prev, dst) else self.lib.pfe_warp_displacement_region
        self._ck(fn(self.h, _ptr(src), sw, sh, _ptr(disp), _ptr(prev), rect, w, h, _ptr(dst)))
        return dst

    @staticmethod
    def _pts(p):
        return None if p is None else np.ascontiguousarray(np.asarray(p, np.float32).reshape(-1, 2))

    def mesh_displacement(self, original, deformed, cols, rows, w, h, device_out=False):
        original, deformed = self._pts(original), self._pts(deformed)
        if device_out:
            self.use_torch_stream()
            out = torch.empty((h, w, 2), dtype=torch.float32, device=f"cuda:{self.device}")
            self._ck(self.lib.pfe_dev_mesh_displacement(self.h, _ptr(original), _ptr(deformed), cols, rows, w, h, _ptr(out)))
        else:
            out = np.empty((h, w, 2), np.float32)
            self._ck(self.lib.pfe_mesh_displacement(self.h, _ptr(original), _ptr(deformed), cols, rows, w, h, _ptr(out)))
        return out

    def mesh_warp(self, src, original, deformed, cols, rows, w, h, y0=0, rows_out=None, out=None):
        src = self._prep(src)
        sh, sw = self._hw(src)
        original, deformed = self._pts(original), self._pts(deformed)
        rows_out = h if rows_out is None else rows_out
        dst = out if out is not None else self._out_like(src, (rows_out, w, 4))
        if self._dev(src, dst):
            self._ck(self.lib.pfe_dev_mesh_warp(self.h, _ptr(src), sw, sh, _ptr(original), _ptr(deformed), cols, rows,
                                                w, h, y0, rows_out, _ptr(dst)))
        else:
            if y0 != 0 or rows_out != h:
                raise ValueError("band form is device-tier only")
            self._ck(self.lib.pfe_mesh_warp(self.h, _ptr(src), sw, sh, _ptr(original), _ptr(deformed), cols, rows, w, h,
                                            _ptr(dst)))
        return dst

    def warp_band(self, src_rows, src_h, src_y0, w, h, y0, rows_out, disp_band=None, original=None, deformed=None,
                  cols=0, rows=0, out=None):
        """Device tier only: rows [y0, y0+rows_out) of a warp whose source is the row window src_rows."""
        sw = int(src_rows.shape[1])
        self._dev(src_rows, disp_band)
        original, deformed = self._pts(original), self._pts(deformed)
        dst = out if out is not None else torch.empty((rows_out, w, 4), dtype=torch.uint8, device=src_rows.device)
        self._ck(self.lib.pfe_dev_warp_band(self.h, _ptr(src_rows), sw, src_h, src_y0, int(src_rows.shape[0]), _ptr(disp_band),
                                            _ptr(original), _ptr(deformed), cols, rows, w, h, y0, rows_out, _ptr(dst)))
        return dst

    def liquify(self, field, kind, cx, cy, radius, strength, a0=0.0, a1=0.0):
        """In place on `field` ((h, w, 2) float32, numpy or CUDA tensor). Returns the bbox."""
        h, w = int(field.shape[0]), int(field.shape[1])
        bbox = (C.c_int32 * 4)()
        if _is_tensor(field):
            self._dev(field)
            fn = self.lib.pfe_dev_liquify
        else:
            if field.dtype != np.float32 or not field.flags.c_contiguous:
                raise ValueError("field must be contiguous float32")
            fn = self.lib.pfe_liquify
        self._ck(fn(self.h, _ptr(field), w, h, int(kind), cx, cy, radius, strength, a0, a1, bbox))
        return tuple(bbox)

    def disp_reach(self, disp_band, y0, h_total, out=None):
        """Device-side reach of a band's field (pfe_dev_disp_reach): int32 tensor [min, max] of floor(y - dy)."""
        rows, w = int(disp_band.shape[0]), int(disp_band.shape[1])
        out = out if out is not None else torch.empty(2, dtype=torch.int32, device=disp_band.device)
        self.use_torch_stream()
        self._ck(self.lib.pfe_dev_disp_reach(self.h, _ptr(disp_band), w, rows, int(y0), int(h_total), _ptr(out)))
        return out

    # -- brush --------------------------------------------------------------------------------
    @staticmethod
    def brush_desc(size, hardness, anti_aliased, color, flow=1.0, is_eraser=False, mode=0):
        b = L.BrushDesc()
        b.size, b.hardness, b.flow = size, hardness, flow
        b.anti_aliased = 1 if anti_aliased else 0
        for i in range(4):
            b.color[i] = color[i]
        b.is_eraser = 1 if is_eraser else 0
        b.mode = int(mode)  # BrushMode: 0 Normal, 1 Dodge, 2 Burn, 3 Sponge
        return b

    def brush_lut(self, brush):
        lut = np.empty(256, np.uint8)
        self.lib.pfe_brush_lut(C.byref(brush), _ptr(lut))
        return lut

    def brush_line_centres(self, w, h, x0, y0, x1, y1):
        cap = int(np.ceil(np.hypot(x1 - x0, y1 - y0))) + 4
        c = np.empty((cap, 2), np.float32)
        n = self.lib.pfe_brush_line_centres(w, h, x0, y0, x1, y1, _ptr(c), cap)
        return c[:n].copy()

    def brush_stamps(self, image, brush, centres, selection_mask=None):
        """In place on `image`; centres is (n, 2) float32 host data."""
        h, w = self._hw(image)
        centres = np.ascontiguousarray(np.asarray(centres, np.float32).reshape(-1, 2))
        selection_mask = self._prep_mask(selection_mask, h, w)
        if _is_tensor(image):
            self._dev(image, selection_mask)
            fn = self.lib.pfe_dev_brush_stamps
        else:
            if image.dtype != np.uint8 or not image.flags.c_contiguous:
                raise ValueError("image must be contiguous uint8")
            fn = self.lib.pfe_brush_stamps
        self._ck(fn(self.h, _ptr(image), w, h, C.byref(brush), _ptr(centres), len(centres), _ptr(selection_mask)))
        return image

    # -- tiles (host only) --------------------------------------------------------------------
    def flat_to_tiles(self, flat, want_tiles=True):
        flat = _np_u8(flat)
        h, w = self._hw(flat)
        cyn, cxn = (h + 63) // 64, (w + 63) // 64
        occ = np.empty((cyn, cxn), np.uint8)
        tiles = np.zeros((cyn * cxn, 64, 64, 4), np.uint8) if want_tiles else None
        rc = self.lib.pfe_flat_to_tiles(_ptr(flat), w, h, _ptr(occ), _ptr(tiles))
        if rc != L.PFE_OK:
            raise L.PfeError(rc, "pfe_flat_to_tiles")
        return occ, tiles

    def tiles_to_flat(self, table, w, h):
        """table: list (row-major chunk grid) of (64,64,4) uint8 arrays or None."""
        n = len(table)
        arr = (C.c_void_p * n)()
        keep = []
        for i, t in enumerate(table):
            if t is not None:
                t = _np_u8(t, (64, 64, 4))
                keep.append(t)
                arr[i] = t.ctypes.data
        flat = np.empty((h, w, 4), np.uint8)
        rc = self.lib.pfe_tiles_to_flat(arr, w, h, _ptr(flat))
        if rc != L.PFE_OK:
            raise L.PfeError(rc, "pfe_tiles_to_flat")
        return flat


_default = None


def default_engine() -> Engine:
    """Process-wide engine on cuda:LOCAL_RANK (one process per GPU)."""
    global _default
    if _default is None:
        _default = Engine()
    return _default
