#!/usr/bin/env python
"""One timing pass over every kernel of the hot path at BASELINE sizes.

  python tools/bench_ops.py [--only REGEX] [--steps K] [--once]      # 1 GPU, device-resident inputs

For each op: CUDA-event time per call (8K = 7680x4320 unless the name says otherwise), the algorithmic bytes
(every input byte read once, every output byte written once) and the fraction of the measured HBM
copy bandwidth (MEASURED_PEAKS.json) that time corresponds to.  One JSON line per op on stdout.
`--once` runs each selected op exactly once after one warm-up call: the form to put under
`ncu --metrics ... -k regex:...` (a number printed under a profiler is never a bench value).
Variants of the Gaussian kernels are selected through the library's tuning knobs, which are
read per call, so one process can A/B them.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from paintfe_b200.engine import Engine, make_layer

W8K, H8K = 7680, 4320


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=".*")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--big", type=int, default=16384, help="edge of the square canvas of the warp ops")
    args = ap.parse_args()
    pat = re.compile(args.only)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    eng = Engine(0)
    eng.use_torch_stream()
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    gen = torch.Generator(device=dev).manual_seed(0x5EED)
    px = W8K * H8K
    img = torch.randint(0, 256, (H8K, W8K, 4), dtype=torch.uint8, device=dev, generator=gen)
    out = torch.empty_like(img)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def timed(fn, env=None):
        old = {}
        for k, v in (env or {}).items():
            old[k] = os.environ.get(k)
            os.environ[k] = v
        try:
            fn()
            if args.once:
                torch.cuda.synchronize()
                torch.cuda.nvtx.range_push("measure")  # ncu --nvtx --nvtx-include "measure/" profiles only this call
                fn()
                torch.cuda.synchronize()
                torch.cuda.nvtx.range_pop()
                return None
            for _ in range(2):
                fn()
            # an 8K image is 133 MB, about the size of L2: flush between iterations and time each call by itself
            tot = 0.0
            for _ in range(args.steps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            return tot / args.steps
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def emit(name, ms, bytes_alg, n_px=px, **extra):
        d = {"op": name, "ms": ms, "mpx_s": None if ms is None else n_px / ms / 1e3, "alg_bytes": bytes_alg}
        if ms is not None:
            d["achieved_gbs"] = bytes_alg / (ms * 1e-3) / 1e9
            d["frac_of_hbm"] = d["achieved_gbs"] / peak
        d.update(extra)
        print(json.dumps(d), flush=True)

    def run(name, fn, bytes_alg, env=None, n_px=px, **extra):
        if not pat.search(name):
            return
        ms = timed(fn, env)
        emit(name, ms, bytes_alg, n_px, **({"env": env} if env else {}), **extra)

    # ---- Gaussian sigma=20 (the headline filter) and its variants ----------------------------------------
    g20 = lambda: eng.gaussian_blur(img, 20.0, out=out)
    run("gaussian s20 fast (default)", g20, 8 * px)
    for tag, env in (("UW=0", {"PFE_GAUSS_UW": "0"}), ("UW=1 NH=16", {"PFE_GAUSS_NH": "16"}),
                     ("UW=1 NV=8", {"PFE_GAUSS_NV": "8"}), ("UW=1 V_WARPS=12", {"PFE_GAUSS_V_WARPS": "12"}),
                     ("UW=1 V_WARPS=12 NV=8", {"PFE_GAUSS_V_WARPS": "12", "PFE_GAUSS_NV": "8"}),
                     ("UW=0 NV=8", {"PFE_GAUSS_UW": "0", "PFE_GAUSS_NV": "8"})):
        run(f"gaussian s20 fast [{tag}]", g20, 8 * px, env=env)
    if torch.cuda.is_available():  # the two passes on their own (band-form entry points; the V pass reuses the H result)
        run("gaussian s20 H pass only", lambda: eng.gaussian_band_h(img, 0, H8K, 20.0), 20 * px)
        for tag, env in (("no staging loads", {"PFE_GAUSS_DBG": "1"}), ("no stores", {"PFE_GAUSS_DBG": "2"}), ("neither", {"PFE_GAUSS_DBG": "3"}),
                         ("in-line staging instead of the register prefetch", {"PFE_GAUSS_DBG": "4"})):
            run(f"gaussian s20 H pass only [diagnosis: {tag}]", lambda: eng.gaussian_band_h(img, 0, H8K, 20.0), 20 * px, env=env)
        eng.gaussian_band_h(img, 0, H8K, 20.0)
        run("gaussian s20 V pass only", lambda: eng.gaussian_band_v(img, 0, H8K, 20.0, out=out), 20 * px)
        for cg in (2, 4, 9, 14):  # ring chunk size in 8-row groups (default: a tile in at most 8 chunks = 4 groups at sigma 20)
            run(f"gaussian s20 V pass only [VCHUNK={cg}]", lambda: eng.gaussian_band_v(img, 0, H8K, 20.0, out=out), 20 * px, env={"PFE_GAUSS_VCHUNK": str(cg)})
    run("gaussian s20 EXACT", lambda: eng.gaussian_blur(img, 20.0, exact=True, out=out), 8 * px)
    run("gaussian s50 fast", lambda: eng.gaussian_blur(img, 50.0, out=out), 8 * px)
    run("gaussian s4 fast (fused H+V)", lambda: eng.gaussian_blur(img, 4.0, out=out), 8 * px)
    run("sharpen a1 r2", lambda: eng.sharpen(img, 1.0, 2.0, out=out), 8 * px)

    # ---- box / motion / median / vignette ------------------------------------------------------------------
    run("box r3", lambda: eng.box_blur(img, 3.0, out=out), 8 * px)
    run("box r40", lambda: eng.box_blur(img, 40.0, out=out), 8 * px)
    run("motion 45deg d10", lambda: eng.motion_blur(img, 45.0, 10.0, out=out), 8 * px)
    run("median r2", lambda: eng.median(img, 2, out=out), 8 * px)
    run("median r7", lambda: eng.median(img, 7, out=out), 8 * px)
    run("median r20", lambda: eng.median(img, 20, out=out), 8 * px)
    for rr in (7, 20):
        run(f"median r{rr} [16-bit column counters]", lambda: eng.median(img, rr, out=out), 8 * px, env={"PFE_MEDIAN_KERNEL": "hist16"})
    run("median r32", lambda: eng.median(img, 32, out=out), 8 * px)
    run("vignette", lambda: eng.vignette(img, 0.8, 0.5, out=out), 8 * px)

    # ---- per-pixel adjustments ---------------------------------------------------------------------------
    lut = eng.levels_lut(20.0, 235.0, 1.2)
    curves = np.stack([eng.curves_lut([(0, 0), (64, 48), (192, 210), (255, 255)])] * 4)
    run("adjust invert", lambda: eng.adjust(img, 0, out=out), 8 * px)
    run("adjust brightness/contrast", lambda: eng.adjust(img, 4, (30.0, 20.0), out=out), 8 * px)
    run("adjust HSL", lambda: eng.adjust(img, 5, (30.0, -20.0, 10.0), out=out), 8 * px)
    run("adjust exposure", lambda: eng.adjust(img, 6, (2.0,), out=out), 8 * px)
    run("adjust levels LUT", lambda: eng.adjust(img, 7, luts=lut, out=out), 8 * px)
    run("adjust curves LUT (4 ch)", lambda: eng.adjust(img, 8, luts=curves, out=out), 8 * px)
    run("adjust vibrance", lambda: eng.adjust(img, 16, (0.5,), out=out), 8 * px)
    run("adjust script HSL (truncating)", lambda: eng.adjust(img, 37, (30.0, -20.0, 10.0), out=out), 8 * px)

    # ---- flatten: the three config-2 stacks ------------------------------------------------------------------
    stacks = (("flatten 16L modes 0-15", 0, False), ("flatten 16L modes 16-24,0-6", 16, False),
              ("flatten 16L modes 0-15 alpha {0,255}", 0, True))
    mode_names = [f"flatten mode {m:02d} x16" for m in range(25)]
    if any(pat.search(n) for n in [s[0] for s in stacks] + mode_names):
        layers = [torch.randint(0, 256, (H8K, W8K, 4), dtype=torch.uint8, device=dev, generator=gen) for _ in range(16)]
        for name, off, binary in stacks:
            ls = layers
            if binary:
                ls = [t.clone() for t in layers]
                for t in ls:
                    t[..., 3] = torch.where(t[..., 3] > 127, 255, 0).to(torch.uint8)
            dl = [make_layer(t, blend=(i + off) % 25, opacity=0.25 + 0.05 * i) for i, t in enumerate(ls)]
            run(name, lambda: eng.flatten(dl, W8K, H8K, out=out), 68 * px)
            del ls, dl
        # one mode at a time (16 layers of the same mode): which modes cost what
        if any(pat.search(n) for n in mode_names):
            for mode in range(25):
                dl = [make_layer(t, blend=mode, opacity=0.25 + 0.05 * i) for i, t in enumerate(layers)]
                run(f"flatten mode {mode:02d} x16", lambda: eng.flatten(dl, W8K, H8K, out=out), 68 * px)
        del layers

    # ---- warps on a big square canvas --------------------------------------------------------------------
    S = args.big
    if any(pat.search(n) for n in ("liquify push", "mesh warp 6x6 fused", "warp displacement", "mesh displacement field")):
        big = torch.randint(0, 256, (S, S, 4), dtype=torch.uint8, device=dev, generator=gen)
        bout = torch.empty_like(big)
        orig = np.zeros((49, 2), np.float32)
        for r in range(7):
            for c in range(7):
                orig[r * 7 + c] = (np.float32(c) / np.float32(6) * S, np.float32(r) / np.float32(6) * S)
        deformed = orig.copy()
        for i in range(7):
            for j in range(7):
                deformed[i * 7 + j] += np.float32(8.0 * np.sin(i) * np.cos(j))
        field = torch.zeros((S, S, 2), dtype=torch.float32, device=dev)
        prng = np.random.default_rng(0x5EED)
        pushes = [(float(prng.uniform(0, S)), float(prng.uniform(0, S)), float(prng.uniform(-20, 20)), float(prng.uniform(-20, 20))) for _ in range(64)]
        k = [0]

        def one_push():
            cx, cy, dx, dy = pushes[k[0] % 64]
            k[0] += 1
            eng.liquify(field, 0, cx, cy, 200.0, 0.8, dx, dy)

        # a push touches a (2r+1)^2 box of the field: read + write 8 bytes per field element
        run(f"liquify push r200 ({S}^2 field)", one_push, 401 * 401 * 16, n_px=401 * 401)
        run(f"mesh warp 6x6 fused ({S}^2)", lambda: eng.mesh_warp(big, orig, deformed, 6, 6, S, S, out=bout), 8 * S * S, n_px=S * S)
        run(f"warp displacement ({S}^2)", lambda: eng.warp_displacement(big, field, out=bout), 16 * S * S, n_px=S * S)
        run(f"mesh displacement field ({S}^2)", lambda: eng.mesh_displacement(orig, deformed, 6, 6, S, S, device_out=True), 8 * S * S, n_px=S * S)
        del big, bout, field

    # ---- brush: one 2000-stamp stroke across the 8K canvas -----------------------------------------------
    if pat.search("brush stroke 2000 stamps size 40"):
        canvas = torch.zeros((H8K, W8K, 4), dtype=torch.uint8, device=dev)
        centres = np.stack([np.linspace(200, W8K - 200, 2000, dtype=np.float32), np.linspace(300, H8K - 300, 2000, dtype=np.float32)], 1)
        brush = eng.brush_desc(40.0, 0.6, True, (0.9, 0.2, 0.1, 1.0), flow=0.8)
        # pixels under the stroke's bounding box are read and written once
        bb = int((centres[:, 0].max() - centres[:, 0].min() + 42) * (centres[:, 1].max() - centres[:, 1].min() + 42))
        run("brush stroke 2000 stamps size 40", lambda: eng.brush_stamps(canvas, brush, centres), 8 * bb, n_px=bb)

    eng.close()


if __name__ == "__main__":
    main()
