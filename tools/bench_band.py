#!/usr/bin/env python
"""Per-kernel times of ONE rank's band step at the band sizes of 1/2/4/8 GPUs (8K canvas, 16 layers, sigma 20), on one GPU
with no transfer: where a small band loses against rows/4320 of the whole-canvas kernel.  CUDA events, L2 flushed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paintfe_b200.engine import Engine, make_layer  # noqa: E402

W, H, NL, SIGMA, R = 7680, 4320, 16, 20.0, 60


def main():
    eng = Engine(0)
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    layers = [torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device=dev, generator=gen) for _ in range(NL)]
    meta = [dict(blend=i % 25, opacity=0.25 + 0.05 * i) for i in range(NL)]

    def timed(fn, n=20):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(n):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / n

    full = None
    for world in (1, 2, 4, 8):
        rows = H // world
        top = bot = (R if world > 1 else 0)
        ext = torch.empty((top + rows + bot, W, 4), dtype=torch.uint8, device=dev)
        out = torch.empty((rows, W, 4), dtype=torch.uint8, device=dev)
        prep = eng.prepare_layers([make_layer(t[:rows], **m) for t, m in zip(layers, meta)], W, rows)
        f = timed(lambda: eng.flatten_prepared(prep, ext[top:top + rows]))
        h = timed(lambda: eng.gaussian_band_h(ext, 0, ext.shape[0], SIGMA))
        v = timed(lambda: eng.gaussian_band_v(ext, top, rows, SIGMA, out=out))

        def chain():
            eng.flatten_prepared(prep, ext[top:top + rows])
            eng.gaussian_band_h(ext, 0, ext.shape[0], SIGMA)
            eng.gaussian_band_v(ext, top, rows, SIGMA, out=out)

        c = timed(chain)
        # the same work as two half-bands on two streams (each flatten -> H), joined before the V pass: the halves fill
        # each other's partial waves and one half's H pass runs under the other's flatten
        h1 = (rows // 2) // 4 * 4
        prep_a = eng.prepare_layers([make_layer(t[:h1], **m) for t, m in zip(layers, meta)], W, h1)
        prep_b = eng.prepare_layers([make_layer(t[h1:rows], **m) for t, m in zip(layers, meta)], W, rows - h1)
        side = torch.cuda.Stream(device=dev, priority=-1)
        ev0, ev1 = torch.cuda.Event(), torch.cuda.Event()

        def split2():
            main = torch.cuda.current_stream(dev)
            ev0.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev0)
                eng.flatten_prepared(prep_a, ext[top:top + h1])
                eng.gaussian_band_h(ext, 0, top + h1, SIGMA)
                ev1.record(side)
            eng.flatten_prepared(prep_b, ext[top + h1:top + rows])
            eng.gaussian_band_h(ext, top + h1, ext.shape[0] - top - h1, SIGMA)
            main.wait_event(ev1)
            eng.gaussian_band_v(ext, top, rows, SIGMA, out=out)

        c2 = timed(split2)
        rec = {"world": world, "band_rows": rows, "ext_rows": ext.shape[0], "flatten_ms": f, "h_ms": h, "v_ms": v, "sum_ms": f + h + v, "chained_ms": c, "split2_ms": c2}
        if full is None:
            full = rec
        else:
            rec["ideal_ms"] = {"flatten": full["flatten_ms"] / world, "h": full["h_ms"] * ext.shape[0] / H, "v": full["v_ms"] / world}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
