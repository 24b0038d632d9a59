#!/usr/bin/env python
"""Small pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are tiny on purpose (the sanitizer slows kernels 10-100x); results are still compared to the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import pfo
from paintfe_b200.engine import Engine, make_layer

eng = Engine(0)
rng = np.random.default_rng(5)
w, h = 200, 150
img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
mask = (rng.random((h, w)) < 0.6).astype(np.uint8) * 255
layers = [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for _ in range(6)]
meta = [dict(blend=(5 * i + 1) % 25, opacity=0.3 + 0.1 * i) for i in range(6)]
ok = True


def check(name, got, exp, tol=0):
    global ok
    d = int(np.abs(np.asarray(got).astype(int) - np.asarray(exp).astype(int)).max())
    print(f"{name:28s} max diff {d}")
    ok = ok and d <= tol


check("flatten", eng.flatten([make_layer(im, **m) for im, m in zip(layers, meta)], w, h),
      pfo.flatten([pfo.make_layer(im, **m) for im, m in zip(layers, meta)], w, h))
for s in (2.0, 20.0):
    check(f"gaussian exact s={s}", eng.gaussian_blur(img, s, exact=True), pfo.gaussian_blur(img, s))
    check(f"gaussian fast s={s}", eng.gaussian_blur(img, s), pfo.gaussian_blur(img, s), 1)
os.environ["PFE_GAUSS_FUSED"] = "1"  # the fused small-radius kernel, which the dispatcher keeps for large images
for s in (0.3, 1.0, 5.3):
    check(f"gaussian fused exact s={s}", eng.gaussian_blur(img, s, exact=True), pfo.gaussian_blur(img, s))
    check(f"gaussian fused fast s={s}", eng.gaussian_blur(img, s), pfo.gaussian_blur(img, s), 1)
check("sharpen fused", eng.sharpen(img, 1.0, 2.0, mask=mask, exact=True), pfo.sharpen(img, 1.0, 2.0, mask=mask))
del os.environ["PFE_GAUSS_FUSED"]
check("gaussian masked", eng.gaussian_blur(img, 3.0, mask=mask, exact=True), pfo.gaussian_blur(img, 3.0, mask=mask))
check("sharpen", eng.sharpen(img, 1.0, 2.0, exact=True), pfo.sharpen(img, 1.0, 2.0))
check("glow", eng.glow(img, 3.0, 0.5, exact=True), pfo.glow(img, 3.0, 0.5))
check("box", eng.box_blur(img, 5.0), pfo.box_blur(img, 5.0))
check("motion", eng.motion_blur(img, 30.0, 8.0), pfo.motion_blur(img, 30.0, 8.0))
check("median", eng.median(img, 2), pfo.median(img, 2))
check("vignette", eng.vignette(img, 0.8, 0.5), pfo.vignette(img, 0.8, 0.5))
check("hsl", eng.adjust(img, pfo.HSL, (30.0, -20.0, 10.0)), pfo.adjust(img, pfo.HSL, (30.0, -20.0, 10.0)))
disp = rng.normal(0, 5, (h, w, 2)).astype(np.float32)
check("warp", eng.warp_displacement(img, disp), pfo.warp_displacement(img, disp))
orig = np.array([[c / 3 * w, r / 3 * h] for r in range(4) for c in range(4)], np.float32)
deformed = (orig + rng.normal(0, 3, orig.shape)).astype(np.float32)
check("mesh warp", eng.mesh_warp(img, orig, deformed, 3, 3, w, h), pfo.mesh_warp(img, orig, deformed, 3, 3, w, h))
check("pixelate", eng.pixelate(img, 7), pfo.pixelate(img, 7))
check("bulge", eng.bulge(img, 0.5), pfo.bulge(img, 0.5))
check("twist", eng.twist(img, 45.0), pfo.twist(img, 45.0), 1)
check("noise perlin", eng.add_noise(img, 50.0, 2, False, 42, 5.0, 3), pfo.add_noise(img, 50.0, 2, False, 42, 5.0, 3))
check("bilateral", eng.reduce_noise(img, 10.0, 2), pfo.reduce_noise(img, 10.0, 2), 1)
# effects3.cu
col = (200, 40, 90, 180)
sparse = img.copy()
sparse[rng.random((h, w)) < 0.6] = 0
for name, args, src in (("ink", (2.0, 0.5), img), ("oil_painting", (10, 64), img), ("color_filter", (col, 0.6, 3), img),
                        ("contours", (10.0, 5.0, 1.0, col, 42, 2, 0.5), img), ("crystallize", (16.0, 42), img),
                        ("dents", (20.0, 10.0, 42, 2, 0.5, True, True), img), ("halftone", (4.0, 45.0, 0), img),
                        ("bokeh_blur", (6.5,), img), ("zoom_blur", (0.5, 0.5, 0.3, 8), img), ("grid", (16, 16, 1, col, 0, 0.7), img),
                        ("canvas_border", (3, col), img), ("outline", (3, col, 2, True), sparse), ("pixel_drag", (42, 50.0, 20, 30.0), img),
                        ("rgb_displace", ((5, 0), (0, 0), (-5, 3)), img)):
    check(name, getattr(eng, name)(src, *args, mask=mask), getattr(pfo, name)(src, *args, mask=mask))
check("drop_shadow", eng.drop_shadow(sparse, 5, -3, 3.0, True, col, 0.8, exact=True), pfo.drop_shadow(sparse, 5, -3, 3.0, True, col, 0.8))
# geometry.cu
for op in range(5):
    check(f"orient {op}", eng.orient(img, op), pfo.orient(img, op))
check("resize_canvas", eng.resize_canvas(img, w + 13, h - 7, (1, 2), col), pfo.resize_canvas(img, w + 13, h - 7, (1, 2), col))
check("affine", eng.affine(img, w, h, 33.0, 10.0, -5.0, 0.8, (3.0, 2.0)), pfo.affine(img, w, h, 33.0, 10.0, -5.0, 0.8, (3.0, 2.0)))
check("resize lanczos", eng.resize(img, 77, 201, 3), pfo.resize(img, 77, 201, 3))
# adjust ops 11-16
for op, prm, luts in ((pfo.THRESHOLD, (128.0,), None), (pfo.POSTERIZE, (4.0,), None), (pfo.COLOR_BALANCE, (10.0, 0.0, -10.0, 0.0, 0.0, 0.0, -10.0, 0.0, 10.0), None),
                      (pfo.GRADIENT_MAP, (), rng.integers(0, 256, (256, 4), dtype=np.uint8)), (pfo.BLACK_AND_WHITE, (30.0, 59.0, 11.0), None), (pfo.VIBRANCE, (0.5,), None)):
    check(f"adjust op {op}", eng.adjust(img, op, prm, luts=luts), pfo.adjust(img, op, prm, luts=luts))
# tiles.cu: host chunk tables and device-resident tiles
occ, tiles = eng.flat_to_tiles(sparse)
table = [tiles[k] if occ.reshape(-1)[k] else None for k in range(occ.size)]
occ2, tiles2 = eng.flat_to_tiles(layers[0])
table2 = [tiles2[k] if occ2.reshape(-1)[k] else None for k in range(occ2.size)]
tl = [dict(tiles=table2, blend=0, opacity=1.0), dict(tiles=table, mask_tiles=table2, blend=8, opacity=0.7), dict(kind=3, opacity=0.5)]
want = pfo.flatten([pfo.make_layer(layers[0]), pfo.make_layer(sparse, blend=8, opacity=0.7, mask=np.ascontiguousarray(layers[0][..., 3])),
                    pfo.make_layer(None, kind=3, opacity=0.5)], w, h, active=occ | occ2)
check("flatten_tiles host", eng.flatten_tiles(tl, w, h), want)
dt = [eng.tiled(w, h).upload(table2), eng.tiled(w, h).upload(table)]
check("flatten_tiles device", eng.flatten_tiles([dict(tiles=dt[0], blend=0, opacity=1.0), dict(tiles=dt[1], mask_tiles=dt[0], blend=8, opacity=0.7),
                                                 dict(kind=3, opacity=0.5)], w, h).cpu().numpy(), want)
for t in dt:
    t.close()
# dodge / burn / sponge
for mode in (1, 2, 3):
    a, b = img.copy(), img.copy()
    pfo.brush_line(a, pfo.make_brush(12.0, 0.8, True, (1, 0, 0, 1), mode=mode), 10.0, 10.0, 150.0, 120.0)
    eng.brush_stamps(b, eng.brush_desc(12.0, 0.8, True, (1, 0, 0, 1), mode=mode), eng.brush_line_centres(w, h, 10.0, 10.0, 150.0, 120.0))
    check(f"brush mode {mode}", b, a)
a, b = img.copy(), img.copy()
br = pfo.make_brush(12.0, 0.8, True, (1, 0, 0, 1))
pfo.brush_line(a, br, 10.0, 10.0, 150.0, 120.0)
eng.brush_stamps(b, eng.brush_desc(12.0, 0.8, True, (1, 0, 0, 1)), eng.brush_line_centres(w, h, 10.0, 10.0, 150.0, 120.0))
check("brush", b, a)
# ---- round 2 kernels --------------------------------------------------------------------------------------------
import torch

for r in (5, 20):  # column histograms with 8-bit counters
    check(f"median hist r={r}", eng.median(img, r), pfo.median(img, r))
os.environ["PFE_MEDIAN_KERNEL"] = "hist16"
check("median hist16 r=5", eng.median(img, 5), pfo.median(img, 5))
del os.environ["PFE_MEDIAN_KERNEL"]
dimg = torch.from_numpy(img).cuda()
# band-form Gaussian (work queues in both passes): rows [40, 110) of the image from an extended band with 60-row halos
for exact in (True, False):
    want = pfo.gaussian_blur(img, 20.0)
    ext = dimg[0:150].contiguous()  # the whole image as "extended band": top halo 40 rows, band 70 rows, bottom halo 40 rows
    eng.gaussian_band_h(ext, 40, 70, 20.0, exact=exact)
    eng.gaussian_band_h(ext, 0, 40, 20.0, exact=exact)
    eng.gaussian_band_h(ext, 110, 40, 20.0, exact=exact)
    got = eng.gaussian_band_v(ext, 40, 70, 20.0, exact=exact)
    check(f"gaussian band exact={exact}", got.cpu().numpy(), want[40:110], 0 if exact else 1)
check("gaussian masked s=20", eng.gaussian_blur(img, 20.0, mask=mask, exact=True), pfo.gaussian_blur(img, 20.0, mask=mask))
# the flatten that also stores into a (here local) peer buffer and releases a flag, then the flag wait
dl = [make_layer(torch.from_numpy(im).cuda(), **m) for im, m in zip(layers, meta)]
prep = eng.prepare_layers(dl, w, h)
near = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
far = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
flags = torch.zeros(4, dtype=torch.int32, device="cuda")
eng.flatten_prepared_peer(prep, near, far.data_ptr(), flags.data_ptr() + 8, 3)
eng.peer_wait(flags.data_ptr() + 8, 1, 3, timeout_ms=5000)
want = pfo.flatten([pfo.make_layer(im, **m) for im, m in zip(layers, meta)], w, h)
check("flatten_peer near", near.cpu().numpy(), want)
check("flatten_peer far", far.cpu().numpy(), want)
eng.peer_signal(flags.data_ptr(), 9)
eng.peer_wait(flags.data_ptr(), 1, 9, timeout_ms=5000)
eng.check_async()
check("peer flags", flags.cpu().numpy(), np.array([9, 0, 3, 0], np.int32))
# region warp and a long binned stroke
rect = (20, 30, 150, 120)
check("warp region", eng.warp_displacement_region(img, disp, img, rect), pfo.warp_displacement_region(img, disp, img, rect))
a, b = img.copy(), img.copy()
pts = [(float(10 + (i * 7) % 180), float(10 + (i * 13) % 130)) for i in range(600)]
brd = pfo.make_brush(9.0, 0.7, True, (0, 1, 0, 1))
for x, y in pts:
    pfo.brush_stamp(a, brd, x, y)
eng.brush_stamps(b, eng.brush_desc(9.0, 0.7, True, (0, 1, 0, 1)), np.array(pts, np.float32))
check("brush 600 stamps", b, a)
eng.close()
print("SANITIZE_SMOKE", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
