#!/usr/bin/env python
"""compute-sanitizer pass over the fused small-radius Gaussian only (mbarrier ring between H and V warps):
  compute-sanitizer --tool racecheck python tools/sanitize_fused.py
Small images, several radii (ring depths 3..6), two strips wide, several segments tall; checked against the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import pfo
from paintfe_b200.engine import Engine

os.environ["PFE_GAUSS_FUSED"] = "1"
os.environ["PFE_GAUSS_FUSED_SEGS"] = "3"
eng = Engine(0)
rng = np.random.default_rng(11)
ok = True
quick = os.environ.get("PFE_SAN_QUICK") == "1"
for (w, h) in (((200, 150),) if quick else ((200, 150), (129, 41))):
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    mask = (rng.random((h, w)) < 0.6).astype(np.uint8) * 255
    for s in ((1.0, 4.0) if quick else (0.3, 1.0, 2.5, 4.0, 5.3)):
        d = int(np.abs(eng.gaussian_blur(img, s, exact=True).astype(int) - pfo.gaussian_blur(img, s).astype(int)).max())
        d2 = int(np.abs(eng.gaussian_blur(img, s).astype(int) - pfo.gaussian_blur(img, s).astype(int)).max())
        d3 = int(np.abs(eng.sharpen(img, 1.0, s, mask=mask, exact=True).astype(int) - pfo.sharpen(img, 1.0, s, mask=mask).astype(int)).max())
        print(f"{w}x{h} sigma={s}: exact {d}, fma {d2}, sharpen {d3}")
        ok = ok and d == 0 and d2 <= 1 and d3 == 0
eng.close()
print("SANITIZE_FUSED", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
