#!/usr/bin/env python
"""Small-radius Gaussian: the fused H+V kernel against the two-pass kernels (PFE_GAUSS_FUSED=0 against =1), same
inputs, device tier.  Prints one JSON line per (size, sigma): both times, whether the outputs are identical."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from paintfe_b200.engine import Engine

eng = Engine(0)
gen = torch.Generator(device="cuda").manual_seed(1)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for (w, h) in ((3840, 2160), (7680, 4320), (1000, 700)):
    img = torch.randint(0, 256, (h, w, 4), dtype=torch.uint8, device="cuda", generator=gen)
    out = torch.empty_like(img)
    for sigma in (0.5, 1.0, 2.0, 4.0, 5.3):
        res = {}
        for exact in (False, True):
            os.environ["PFE_GAUSS_FUSED"] = "1"
            fused = eng.gaussian_blur(img, sigma, exact=exact).clone()
            t_fused = timed(lambda: eng.gaussian_blur(img, sigma, exact=exact, out=out))
            os.environ["PFE_GAUSS_FUSED"] = "0"
            two = eng.gaussian_blur(img, sigma, exact=exact).clone()
            t_two = timed(lambda: eng.gaussian_blur(img, sigma, exact=exact, out=out))
            os.environ["PFE_GAUSS_FUSED"] = "1"
            res["exact" if exact else "fast"] = dict(fused_ms=round(t_fused, 4), two_pass_ms=round(t_two, 4),
                                                     identical=bool(torch.equal(fused, two)))
        sh = timed(lambda: eng.sharpen(img, 1.0, sigma, out=out))
        os.environ["PFE_GAUSS_FUSED"] = "0"
        sh2 = timed(lambda: eng.sharpen(img, 1.0, sigma, out=out))
        os.environ["PFE_GAUSS_FUSED"] = "1"
        print(json.dumps(dict(w=w, h=h, sigma=sigma, **res, sharpen_fused_ms=round(sh, 4), sharpen_two_pass_ms=round(sh2, 4))), flush=True)
eng.close()
