"""ctypes binding of libpfe_b200.so (the C ABI in include/pfe_b200.h).

Loading never falls back to anything: if the shared library is missing the import of the symbol
table raises, and every compute call needs a CUDA device (PFE_ERR_NO_DEVICE otherwise).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpfe_b200.so")

PFE_OK = 0
STATUS = {0: "PFE_OK", -1: "PFE_ERR_INVALID_ARG", -2: "PFE_ERR_UNSUPPORTED", -3: "PFE_ERR_CUDA",
          -4: "PFE_ERR_NO_DEVICE", -5: "PFE_ERR_OOM"}
GAUSS_EXACT = 1
MESH_MAX_POINTS = 256
CHUNK = 64


class PfeError(RuntimeError):
    def __init__(self, code: int, msg: str = ""):
        self.code = code
        super().__init__(f"{STATUS.get(code, code)}: {msg}")


class LayerDesc(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("mask", C.c_void_p), ("opacity", C.c_float), ("blend", C.c_uint8),
                ("visible", C.c_uint8), ("kind", C.c_uint8), ("_pad", C.c_uint8), ("adj", C.c_float * 16)]


class AdjustDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("params", C.c_float * 12), ("luts", C.c_void_p)]


class TileLayerDesc(C.Structure):
    _fields_ = [("chunks", C.c_void_p), ("mask_chunks", C.c_void_p), ("opacity", C.c_float), ("blend", C.c_uint8),
                ("visible", C.c_uint8), ("kind", C.c_uint8), ("_pad", C.c_uint8), ("adj", C.c_float * 16)]


class BrushDesc(C.Structure):
    _fields_ = [("size", C.c_float), ("hardness", C.c_float), ("flow", C.c_float), ("anti_aliased", C.c_int32),
                ("color", C.c_float * 4), ("is_eraser", C.c_int32), ("mode", C.c_int32)]


_u32, _f32, _vp, _i32 = C.c_uint32, C.c_float, C.c_void_p, C.c_int32
_ctx = C.c_void_p

# name -> (restype, argtypes).  Must list every symbol include/pfe_b200.h declares
# (tests/test_abi.py cross-checks the header against this table and against the .so).
SIGNATURES = {
    "pfe_abi_version": (C.c_int, []),
    "pfe_ctx_create": (C.c_int, [C.c_int, C.POINTER(_ctx)]),
    "pfe_ctx_destroy": (C.c_int, [_ctx]),
    "pfe_ctx_set_stream": (C.c_int, [_ctx, _vp]),
    "pfe_ctx_use_own_stream": (C.c_int, [_ctx]),
    "pfe_ctx_sync": (C.c_int, [_ctx]),
    "pfe_ctx_check_async": (C.c_int, [_ctx]),
    "pfe_last_error": (C.c_char_p, [_ctx]),
    "pfe_ctx_launch_count": (C.c_uint64, [_ctx]),
    "pfe_ctx_profile": (C.c_int, [_ctx, C.c_int]),
    "pfe_ctx_profile_read": (C.c_int, [_ctx, C.c_char_p, C.c_size_t]),
    "pfe_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "pfe_host_free": (C.c_int, [_vp]),
    "pfe_dev_alloc": (C.c_int, [_ctx, C.c_size_t, C.POINTER(_vp)]),
    "pfe_dev_free": (C.c_int, [_ctx, _vp]),
    "pfe_dev_upload": (C.c_int, [_ctx, _vp, _vp, C.c_size_t]),
    "pfe_dev_download": (C.c_int, [_ctx, _vp, _vp, C.c_size_t]),
    "pfe_flatten": (C.c_int, [_ctx, C.POINTER(LayerDesc), _u32, _u32, _u32, _vp, _vp]),
    "pfe_dev_flatten": (C.c_int, [_ctx, C.POINTER(LayerDesc), _u32, _u32, _u32, _vp, _vp]),
    "pfe_flatten_tiles": (C.c_int, [_ctx, C.POINTER(TileLayerDesc), _u32, _u32, _u32, _vp]),
    "pfe_dev_flatten_tiles": (C.c_int, [_ctx, C.POINTER(TileLayerDesc), _u32, _u32, _u32, _vp]),
    "pfe_tiled_create": (C.c_int, [_ctx, _u32, _u32, C.POINTER(_vp)]),
    "pfe_tiled_destroy": (C.c_int, [_ctx, _vp]),
    "pfe_tiled_clone": (C.c_int, [_ctx, _vp, C.POINTER(_vp)]),
    "pfe_tiled_make_mut": (C.c_int, [_ctx, _vp, _vp, _u32]),
    "pfe_tiled_chunk_ids": (C.c_int, [_ctx, _vp, _vp]),
    "pfe_tiled_upload": (C.c_int, [_ctx, _vp, _vp]),
    "pfe_tiled_from_flat": (C.c_int, [_ctx, _vp, _vp]),
    "pfe_tiled_to_flat": (C.c_int, [_ctx, _vp, _vp]),
    "pfe_tiled_download": (C.c_int, [_ctx, _vp, _vp, _vp]),
    "pfe_tiled_table": (_vp, [_vp]),
    "pfe_tiled_occupancy": (_vp, [_vp]),
    "pfe_gaussian_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _vp, _vp, _u32]),
    "pfe_dev_gaussian_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _vp, _vp, _u32]),
    "pfe_box_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _vp, _vp]),
    "pfe_dev_box_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _vp, _vp]),
    "pfe_motion_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp]),
    "pfe_dev_motion_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp]),
    "pfe_median": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _vp, _vp]),
    "pfe_dev_median": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _vp, _vp]),
    "pfe_sharpen": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp, _u32]),
    "pfe_dev_sharpen": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp, _u32]),
    "pfe_vignette": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp]),
    "pfe_dev_vignette": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp]),
    "pfe_glow": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp, _u32]),
    "pfe_dev_glow": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp, _u32]),
    "pfe_pixelate": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _vp, _vp]),
    "pfe_dev_pixelate": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _vp, _vp]),
    "pfe_bulge": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _f32, _vp, _vp]),
    "pfe_dev_bulge": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _f32, _vp, _vp]),
    "pfe_twist": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _f32, _vp, _vp]),
    "pfe_dev_twist": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _f32, _vp, _vp]),
    "pfe_add_noise": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, C.c_int, C.c_int, _u32, _f32, _u32, _vp, _vp]),
    "pfe_dev_add_noise": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, C.c_int, C.c_int, _u32, _f32, _u32, _vp, _vp]),
    "pfe_reduce_noise": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _u32, _vp, _vp]),
    "pfe_dev_reduce_noise": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _u32, _vp, _vp]),
    "pfe_ink": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp]),
    "pfe_dev_ink": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _vp, _vp]),
    "pfe_oil_painting": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _vp, _vp]),
    "pfe_dev_oil_painting": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _vp, _vp]),
    "pfe_color_filter": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _f32, C.c_int, _vp, _vp]),
    "pfe_dev_color_filter": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _f32, C.c_int, _vp, _vp]),
    "pfe_contours": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _f32, _vp, _u32, _u32, _f32, _vp, _vp]),
    "pfe_dev_contours": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _f32, _vp, _u32, _u32, _f32, _vp, _vp]),
    "pfe_crystallize": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _u32, _vp, _vp]),
    "pfe_dev_crystallize": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _u32, _vp, _vp]),
    "pfe_dents": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _u32, _u32, _f32, C.c_int, C.c_int, _vp, _vp]),
    "pfe_dev_dents": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _u32, _u32, _f32, C.c_int, C.c_int, _vp, _vp]),
    "pfe_halftone": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, C.c_int, _vp, _vp]),
    "pfe_dev_halftone": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, C.c_int, _vp, _vp]),
    "pfe_bokeh_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _vp, _vp]),
    "pfe_dev_bokeh_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _vp, _vp]),
    "pfe_zoom_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _f32, _u32, _vp, _f32, _vp, _vp]),
    "pfe_dev_zoom_blur": (C.c_int, [_ctx, _vp, _u32, _u32, _f32, _f32, _f32, _u32, _vp, _f32, _vp, _vp]),
    "pfe_grid": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _u32, _vp, C.c_int, _f32, _vp, _vp]),
    "pfe_dev_grid": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _u32, _vp, C.c_int, _f32, _vp, _vp]),
    "pfe_canvas_border": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "pfe_dev_canvas_border": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "pfe_drop_shadow": (C.c_int, [_ctx, _vp, _u32, _u32, C.c_int32, C.c_int32, _f32, C.c_int, _vp, _f32, _vp, _vp, _u32]),
    "pfe_dev_drop_shadow": (C.c_int, [_ctx, _vp, _u32, _u32, C.c_int32, C.c_int32, _f32, C.c_int, _vp, _f32, _vp, _vp, _u32]),
    "pfe_outline": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _vp, C.c_int, C.c_int, _vp, _vp]),
    "pfe_dev_outline": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _vp, C.c_int, C.c_int, _vp, _vp]),
    "pfe_pixel_drag": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _f32, _u32, _f32, _vp, _vp]),
    "pfe_dev_pixel_drag": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _f32, _u32, _f32, _vp, _vp]),
    "pfe_rgb_displace": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _vp, _vp]),
    "pfe_dev_rgb_displace": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _vp, _vp]),
    "pfe_adjust": (C.c_int, [_ctx, _vp, _u32, _u32, C.POINTER(AdjustDesc), _vp, _vp, _vp]),
    "pfe_dev_adjust": (C.c_int, [_ctx, _vp, _u32, _u32, C.POINTER(AdjustDesc), _vp, _vp, _vp]),
    "pfe_build_levels_lut": (None, [_f32, _f32, _f32, _f32, _f32, _vp]),
    "pfe_build_levels_lut_script": (None, [_f32, _f32, _f32, _vp]),
    "pfe_build_stretch_lut": (None, [C.c_uint8, C.c_uint8, _vp]),
    "pfe_build_curves_lut": (None, [_vp, C.c_int, _vp]),
    "pfe_compose_curve_luts": (None, [_vp, _vp]),
    "pfe_channel_minmax": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _vp]),
    "pfe_dev_channel_minmax": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _vp]),
    "pfe_orient": (C.c_int, [_ctx, _vp, _u32, _u32, C.c_int, _vp]),
    "pfe_dev_orient": (C.c_int, [_ctx, _vp, _u32, _u32, C.c_int, _vp]),
    "pfe_resize_canvas": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp]),
    "pfe_dev_resize_canvas": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp, _vp]),
    "pfe_affine": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _f32, _f32, _f32, _f32, _f32, _f32, C.c_int, _vp]),
    "pfe_dev_affine": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _f32, _f32, _f32, _f32, _f32, _f32, C.c_int, _vp]),
    "pfe_resize": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, C.c_int, _vp]),
    "pfe_dev_resize": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, C.c_int, _vp]),
    "pfe_warp_displacement": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _u32, _u32, _vp]),
    "pfe_dev_warp_displacement": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _u32, _u32, _vp]),
    "pfe_warp_displacement_region": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _vp, _vp, _u32, _u32, _vp]),
    "pfe_dev_warp_displacement_region": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _vp, _vp, _u32, _u32, _vp]),
    "pfe_mesh_displacement": (C.c_int, [_ctx, _vp, _vp, _u32, _u32, _u32, _u32, _vp]),
    "pfe_dev_mesh_displacement": (C.c_int, [_ctx, _vp, _vp, _u32, _u32, _u32, _u32, _vp]),
    "pfe_mesh_warp": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _vp, _u32, _u32, _u32, _u32, _vp]),
    "pfe_dev_mesh_warp": (C.c_int, [_ctx, _vp, _u32, _u32, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp]),
    "pfe_dev_warp_band": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _u32, _u32, _u32, _u32, _u32, _u32, _vp]),
    "pfe_dev_gaussian_band_h": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _f32, _u32]),
    "pfe_dev_gaussian_band_v": (C.c_int, [_ctx, _u32, _u32, _u32, _u32, _f32, _vp, _u32]),
    "pfe_peer_alloc": (C.c_int, [_ctx, C.c_size_t, C.POINTER(_vp), _vp]),
    "pfe_peer_open": (C.c_int, [_ctx, _vp, C.POINTER(_vp)]),
    "pfe_peer_close": (C.c_int, [_ctx, _vp]),
    "pfe_peer_free": (C.c_int, [_ctx, _vp]),
    "pfe_dev_flatten_peer": (C.c_int, [_ctx, C.POINTER(LayerDesc), _u32, _u32, _u32, _vp, _vp, _vp, _vp, _u32]),
    "pfe_dev_peer_signal": (C.c_int, [_ctx, _vp, _u32]),
    "pfe_dev_peer_wait": (C.c_int, [_ctx, _vp, _u32, _u32, _u32]),
    "pfe_dev_disp_reach": (C.c_int, [_ctx, _vp, _u32, _u32, _u32, _u32, _vp]),
    "pfe_liquify": (C.c_int, [_ctx, _vp, _u32, _u32, C.c_int, _f32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "pfe_dev_liquify": (C.c_int, [_ctx, _vp, _u32, _u32, C.c_int, _f32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "pfe_brush_stamps": (C.c_int, [_ctx, _vp, _u32, _u32, C.POINTER(BrushDesc), _vp, _u32, _vp]),
    "pfe_dev_brush_stamps": (C.c_int, [_ctx, _vp, _u32, _u32, C.POINTER(BrushDesc), _vp, _u32, _vp]),
    "pfe_brush_line_centres": (C.c_int, [_u32, _u32, _f32, _f32, _f32, _f32, _vp, C.c_int]),
    "pfe_brush_lut": (None, [C.POINTER(BrushDesc), _vp]),
    "pfe_tiles_to_flat": (C.c_int, [_vp, _u32, _u32, _vp]),
    "pfe_flat_to_tiles": (C.c_int, [_vp, _u32, _u32, _vp, _vp]),
    "pfe_flatten_gaussian": (C.c_int, [_ctx, C.POINTER(LayerDesc), _u32, _u32, _u32, _vp, _f32, _vp, _u32]),
}

_lib = None


def load():
    """dlopen libpfe_b200.so and bind every symbol. Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m paintfe_b200.build` "
                              "(there is no CPU fallback for the pixel engine)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.pfe_abi_version() != 1:
            raise ImportError("libpfe_b200.so ABI version mismatch")
        _lib = lib
    return _lib
