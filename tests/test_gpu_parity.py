"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference goldens.

Bar: bit-exact for flatten, adjustments, box, motion, median, vignette, sharpen/gaussian in EXACT
mode, warps and brush stamps; <= 1 level per channel for the default (FMA) Gaussian family and for
liquify-derived warps (device exp vs libm expf).
"""
import numpy as np
import pytest

import fixtures as fx
from test_oracle_golden import ADJUST, BLEND_IDS, FILTERS, GEOMETRY, SCRIPT, STROKES, BLACK

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from paintfe_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()


def exact(a, b, what=""):
    bad, mx = fx.diff_stats(np.asarray(a), np.asarray(b))
    assert bad == 0, f"{what}: {bad} mismatched pixels, max diff {mx}"


def within1(a, b, what=""):
    d = np.abs(np.asarray(a).astype(np.int16) - np.asarray(b).astype(np.int16))
    assert d.max() <= 1, f"{what}: max diff {d.max()}"
    return float((d > 0).mean())


# ---------------------------------------------------------------------------------------------
# goldens, straight through the engine
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(BLEND_IDS))
def test_blend_golden(eng, name):
    from paintfe_b200.engine import make_layer

    out = eng.flatten([make_layer(fx.checkerboard(64, 64)), make_layer(fx.blend_foreground(), blend=BLEND_IDS[name])], 64, 64)
    exact(out, fx.golden("blend", name), name)


def test_blend_half_opacity_golden(eng):
    from paintfe_b200.engine import make_layer

    out = eng.flatten([make_layer(fx.checkerboard(64, 64)), make_layer(fx.gradient(64, 64), opacity=0.5)], 64, 64)
    exact(out, fx.golden("blend", "normal_half_opacity"))


class _EngAsOracle:
    """Adapter so the golden tables written for the oracle drive the engine unchanged."""

    def __init__(self, eng, exact_gauss=True):
        self.e, self.x = eng, exact_gauss
        from oracle import pfo
        for k in dir(pfo):
            if k.isupper():
                setattr(self, k, getattr(pfo, k))

    def gaussian_blur(self, im, s, mask=None): return self.e.gaussian_blur(im, s, mask=mask, exact=self.x)
    def box_blur(self, im, r, mask=None): return self.e.box_blur(im, r, mask=mask)
    def motion_blur(self, im, a, d, mask=None): return self.e.motion_blur(im, a, d, mask=mask)
    def median(self, im, r, mask=None): return self.e.median(im, r, mask=mask)
    def sharpen(self, im, a, r, mask=None): return self.e.sharpen(im, a, r, mask=mask, exact=self.x)
    def vignette(self, im, a, s, mask=None): return self.e.vignette(im, a, s, mask=mask)
    def glow(self, im, r, i, mask=None): return self.e.glow(im, r, i, mask=mask, exact=self.x)
    def pixelate(self, im, b, mask=None): return self.e.pixelate(im, b, mask=mask)
    def bulge(self, im, a, origin=(0.5, 0.5), mask=None): return self.e.bulge(im, a, origin, mask=mask)
    def twist(self, im, a, origin=(0.5, 0.5), mask=None): return self.e.twist(im, a, origin, mask=mask)
    def add_noise(self, im, amount, t, mono, seed, scale, octaves, mask=None):
        return self.e.add_noise(im, amount, t, mono, seed, scale, octaves, mask=mask)
    def reduce_noise(self, im, s, r, mask=None): return self.e.reduce_noise(im, s, r, mask=mask)
    def drop_shadow(self, im, ox, oy, br, widen, color, op, mask=None):
        return self.e.drop_shadow(im, ox, oy, br, widen, color, op, mask=mask, exact=self.x)
    def flatten(self, layers, w, h): return self.e.flatten(layers, w, h)
    def make_layer(self, *a, **k):
        from paintfe_b200.engine import make_layer
        return make_layer(*a, **k)
    def __getattr__(self, name):  # the remaining effects take the oracle's arguments unchanged
        if name in ("orient", "resize_canvas", "affine", "resize", "ink", "oil_painting", "color_filter", "contours", "crystallize", "dents", "halftone", "bokeh_blur",
                    "zoom_blur", "grid", "canvas_border", "outline", "pixel_drag", "rgb_displace"):
            return getattr(self.e, name)
        raise AttributeError(name)
    def adjust(self, im, op, params=(), luts=None, mask=None, occupancy=None):
        return self.e.adjust(im, op, params, luts=luts, mask=mask, occupancy=occupancy)
    def levels_lut(self, *a): return self.e.levels_lut(*a)
    def levels_lut_script(self, *a): return self.e.levels_lut_script(*a)
    def stretch_lut(self, mn, mx): return self.e.stretch_lut(mn, mx)


TRANSCENDENTAL = {"twist_45", "add_noise_gaussian_mono", "reduce_noise"}  # sin/cos/ln/exp per pixel on the device


@pytest.mark.parametrize("name", sorted(FILTERS))
def test_filter_golden_exact(eng, name):
    got = FILTERS[name](_EngAsOracle(eng, True), fx.gradient(64, 64))
    if name in TRANSCENDENTAL:
        assert within1(got, fx.golden("filters", name), name) < 0.01  # and almost always identical
    else:
        exact(got, fx.golden("filters", name), name)


@pytest.mark.parametrize("name", sorted(FILTERS))
def test_filter_golden_fast(eng, name):
    within1(FILTERS[name](_EngAsOracle(eng, False), fx.gradient(64, 64)), fx.golden("filters", name), name)


@pytest.mark.parametrize("name", sorted(ADJUST))
def test_adjust_golden(eng, oracle, name):
    if name == "auto_levels":
        im = fx.gradient(64, 64)
        mm = eng.channel_minmax(im)
        luts = np.stack([eng.stretch_lut(mm[0], mm[1]), eng.stretch_lut(mm[2], mm[3]), eng.stretch_lut(mm[4], mm[5]),
                         np.arange(256, dtype=np.uint8)])
        out = eng.adjust(im, oracle.LUT_RGBA, luts=luts)
    else:
        out = ADJUST[name](_EngAsOracle(eng), fx.gradient(64, 64))
    occ, tiles = eng.flat_to_tiles(out)
    table = [tiles[i] if occ.reshape(-1)[i] else None for i in range(occ.size)]
    exact(eng.tiles_to_flat(table, 64, 64), fx.golden("adjustments", name), name)


@pytest.mark.parametrize("name", sorted(SCRIPT))
def test_scripting_golden(eng, name):
    exact(SCRIPT[name](_EngAsOracle(eng), fx.gradient(64, 64)), fx.golden("scripting", name), name)


def test_warp_goldens(eng, oracle):
    exact(eng.warp_displacement(fx.gradient_32(), fx.swirl_field()), fx.golden("transform", "displacement_swirl"))
    orig = fx.uniform_grid(2, 2, 32.0, 32.0)
    deformed = orig.copy()
    deformed[4] = (20.0, 20.0)
    exact(eng.mesh_warp(fx.gradient_32(), orig, deformed, 2, 2, 32, 32), fx.golden("transform", "mesh_warp_deformed"))
    d = eng.mesh_displacement(orig, deformed, 2, 2, 32, 32)
    assert np.array_equal(d, oracle.mesh_displacement(orig, deformed, 2, 2, 32, 32))
    exact(eng.warp_displacement(fx.gradient_32(), d), fx.golden("transform", "mesh_warp_deformed"))
    field = np.zeros((32, 32, 2), np.float32)
    assert eng.liquify(field, 0, 16.0, 16.0, 10.0, 0.8, 3.0, 0.0) == (6, 6, 26, 26)
    within1(eng.warp_displacement(fx.gradient_32(), field), fx.golden("transform", "displacement_radial_push"))


@pytest.mark.parametrize("name", sorted(STROKES))
def test_stroke_golden(eng, name):
    size, hard, aa, color, eraser, bg, kind, args = STROKES[name]
    img = fx.solid(64, 64, (255, 255, 255, 255) if bg == "white" else (0, 0, 0, 0))
    b = eng.brush_desc(size, hard, aa, color, is_eraser=eraser)
    centres = np.asarray(args, np.float32) if kind == "stamp" else eng.brush_line_centres(64, 64, *args)
    eng.brush_stamps(img, b, centres)
    exact(img, fx.golden("tools", name), name)


@pytest.mark.parametrize("name,mode", [("brush_dodge_mode", 1), ("brush_burn_mode", 2)])
def test_stroke_mode_golden(eng, name, mode):
    img = fx.gradient(64, 64)
    eng.brush_stamps(img, eng.brush_desc(24.0, 1.0, True, BLACK, mode=mode), [(32.0, 32.0)])
    exact(img, fx.golden("tools", name), name)


def test_stroke_modes_random(eng, oracle):
    """Dodge / burn / sponge over overlapping stamps (each stamp re-reads what the previous one wrote)."""
    rng = np.random.default_rng(77)
    for mode in (1, 2, 3):
        for aa in (True, False):
            img = fx.random_rgba(rng, 97, 61)
            want = img.copy()
            color = (0.2, 0.4, 0.6, float(rng.uniform(0.3, 1.0)))
            centres = np.stack([np.linspace(5, 90, 40), np.linspace(50, 8, 40)], 1).astype(np.float32)
            ob = oracle.make_brush(17.0, 0.6, aa, color, flow=0.8, mode=mode)
            for cx, cy in centres:
                oracle.brush_stamp(want, ob, float(cx), float(cy))
            eng.brush_stamps(img, eng.brush_desc(17.0, 0.6, aa, color, flow=0.8, mode=mode), centres)
            exact(img, want, f"brush mode {mode} aa {aa}")


def test_stroke_selection_mask_golden(eng):
    img = fx.solid(64, 64, (0, 0, 0, 0))
    mask = np.zeros((64, 64), np.uint8)
    mask[:, :32] = 255
    eng.brush_stamps(img, eng.brush_desc(40.0, 1.0, True, BLACK), [(32.0, 32.0)], selection_mask=mask)
    exact(img, fx.golden("tools", "brush_with_selection_mask"))


# ---------------------------------------------------------------------------------------------
# randomised parity against the oracle
# ---------------------------------------------------------------------------------------------
SIZES = [(64, 64), (67, 45), (1, 1), (3, 130), (256, 192), (129, 1)]


def _stack(rng, oracle, w, h, n, alpha="uniform", masks=True, adj=True):
    from paintfe_b200.engine import make_layer

    o_layers, e_layers = [], []
    for i in range(n):
        kind = 0
        if adj and rng.random() < 0.15:
            kind = int(rng.integers(1, 5))
        if kind:
            params = {1: (float(2.0 ** rng.uniform(-2, 2)),), 2: (float(rng.uniform(-100, 100)), float(rng.uniform(-100, 100))),
                      3: (), 4: tuple(float(v) for v in rng.uniform(-0.5, 1.2, 16))}[kind]
            kw = dict(opacity=float(rng.choice([1.0, 0.0, rng.uniform(0, 1)])), kind=kind, adj=params,
                      visible=bool(rng.random() > 0.1))
            o_layers.append(oracle.make_layer(**kw))
            e_layers.append(make_layer(**kw))
            continue
        img = fx.random_rgba(rng, w, h, alpha)
        mask = None
        if masks and rng.random() < 0.3:
            mask = rng.integers(0, 256, (h, w), dtype=np.uint8)
            mask[rng.random((h, w)) < 0.5] = 0
        kw = dict(opacity=float(rng.choice([1.0, 0.0, 0.5, 1.5, -0.2, rng.uniform(0, 1)])),
                  blend=int(rng.integers(0, 25)) if rng.random() > 0.05 else 77, visible=bool(rng.random() > 0.1), mask=mask)
        o_layers.append(oracle.make_layer(img, **kw))
        e_layers.append(make_layer(img, **kw))
    return o_layers, e_layers


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("alpha", ["uniform", "binary"])
def test_flatten_random_stacks(eng, oracle, w, h, alpha):
    rng = np.random.default_rng(w * 1000 + h + (7 if alpha == "binary" else 0))
    for n in (0, 1, 2, 7, 16, 40):
        ol, el = _stack(rng, oracle, w, h, n, alpha)
        exact(eng.flatten(el, w, h), oracle.flatten(ol, w, h), f"{w}x{h} n={n}")


def test_flatten_every_mode_every_byte_pair(eng, oracle):
    """All 25 modes over a dense sweep of (base, top) channel/alpha values, several opacities."""
    from paintfe_b200.engine import make_layer

    rng = np.random.default_rng(11)
    v = np.arange(256, dtype=np.uint8)
    base = np.empty((256, 256, 4), np.uint8)
    top = np.empty((256, 256, 4), np.uint8)
    base[..., 0] = v[:, None]; top[..., 0] = v[None, :]
    for op in (1.0, 0.5, 0.25, 0.999):
        base[..., 1] = rng.integers(0, 256, (256, 256)); top[..., 1] = rng.integers(0, 256, (256, 256))
        base[..., 2] = v[None, ::-1]; top[..., 2] = v[:, None]
        base[..., 3] = rng.choice([0, 1, 127, 128, 254, 255], (256, 256)); top[..., 3] = rng.choice([0, 1, 64, 200, 254, 255], (256, 256))
        for mode in range(25):
            got = eng.flatten([make_layer(base, blend=14), make_layer(top, blend=mode, opacity=op)], 256, 256)
            exp = oracle.flatten([oracle.make_layer(base, blend=14), oracle.make_layer(top, blend=mode, opacity=op)], 256, 256)
            exact(got, exp, f"mode {mode} opacity {op}")


def test_flatten_active_chunks_with_adjustment(eng, oracle):
    """Adjustment layers only touch active chunks (canvas_state.rs:579-584)."""
    from paintfe_b200.engine import make_layer

    rng = np.random.default_rng(5)
    w, h = 200, 130
    img = fx.random_rgba(rng, w, h)
    img[:64, 64:128] = 0
    _, occ = oracle.tiled_roundtrip(img)
    assert occ[0, 1] == 0
    kw = [dict(rgba=img), dict(kind=3, opacity=1.0)]
    got = eng.flatten([make_layer(**k) for k in kw], w, h, active=occ)
    exp = oracle.flatten([oracle.make_layer(**k) for k in kw], w, h, active=occ)
    exact(got, exp)
    assert tuple(got[10, 70]) == (0, 0, 0, 0)
    dense = eng.flatten([make_layer(**k) for k in kw], w, h)
    assert tuple(dense[10, 70]) == (255, 255, 255, 0)


@pytest.mark.parametrize("w,h", SIZES + [(700, 300)])
@pytest.mark.parametrize("sigma", [0.0, 0.3, 0.7, 2.0, 5.0, 9.5, 20.0])
def test_gaussian_random(eng, oracle, w, h, sigma):
    rng = np.random.default_rng(int(sigma * 10) + w)
    img = fx.random_rgba(rng, w, h)
    exp = oracle.gaussian_blur(img, sigma)
    exact(eng.gaussian_blur(img, sigma, exact=True), exp, f"exact {w}x{h} s={sigma}")
    within1(eng.gaussian_blur(img, sigma), exp, f"fast {w}x{h} s={sigma}")


@pytest.mark.parametrize("w,h", [(300, 200), (129, 77), (1, 50), (50, 1), (640, 481)])
@pytest.mark.parametrize("sigma", [0.2, 1.0, 2.5, 4.0, 5.3])
def test_gaussian_fused_small_radius(eng, oracle, monkeypatch, w, h, sigma):
    """The fused H+V kernel (radius <= 16; normally picked for large images only) forced on small ones: bit-exact
    with the oracle in exact mode, identical to the two-pass kernels in FMA mode, epilogues and selection included."""
    rng = np.random.default_rng(int(sigma * 10) + 7 * w)
    img = fx.random_rgba(rng, w, h)
    mask = (rng.random((h, w)) < 0.5).astype(np.uint8) * 255
    monkeypatch.setenv("PFE_GAUSS_FUSED", "0")
    two_pass = [eng.gaussian_blur(img, sigma), eng.sharpen(img, 1.5, sigma), eng.glow(img, sigma, 0.7, mask=mask)]
    monkeypatch.setenv("PFE_GAUSS_FUSED", "1")
    exact(eng.gaussian_blur(img, sigma, exact=True), oracle.gaussian_blur(img, sigma), f"fused exact {w}x{h} s={sigma}")
    exact(eng.gaussian_blur(img, sigma, mask=mask, exact=True), oracle.gaussian_blur(img, sigma, mask=mask), "fused, selection crop")
    exact(eng.sharpen(img, 1.5, sigma, mask=mask, exact=True), oracle.sharpen(img, 1.5, sigma, mask=mask), "fused sharpen")
    exact(eng.glow(img, sigma, 0.7, exact=True), oracle.glow(img, sigma, 0.7), "fused glow")
    for got, ref, what in zip([eng.gaussian_blur(img, sigma), eng.sharpen(img, 1.5, sigma), eng.glow(img, sigma, 0.7, mask=mask)],
                              two_pass, ("blur", "sharpen", "glow")):
        exact(got, ref, f"fused FMA {what} == two-pass FMA {what}")


def test_gaussian_fused_several_tasks_per_cta(eng, oracle, monkeypatch):
    """More strip segments than resident CTAs: a CTA walks several segments in a row and the ring's barrier phases run
    on across them (what happens by itself on an 8K image)."""
    monkeypatch.setenv("PFE_GAUSS_FUSED", "1")
    monkeypatch.setenv("PFE_GAUSS_FUSED_SEGS", "40")  # 11 strips x 40 segments = 440 tasks > 2 CTAs x 148 SMs
    img = fx.random_rgba(np.random.default_rng(77), 1300, 1900)
    for sigma in (0.8, 3.0):
        exact(eng.gaussian_blur(img, sigma, exact=True), oracle.gaussian_blur(img, sigma), f"fused, many segments, s={sigma}")


def test_gaussian_fused_is_picked_for_large_images(eng, oracle):
    """4K, sigma 2: the dispatcher takes the fused kernel by itself (launch counter: 1 kernel instead of 2)."""
    import torch

    img = torch.randint(0, 256, (2160, 3840, 4), dtype=torch.uint8, device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    n0 = eng.launches
    blurred = eng.gaussian_blur(img, 2.0, exact=True)
    assert eng.launches - n0 == 1
    y0, x0 = 1000, 1800  # interior window against the oracle on a crop with a radius-6 apron
    crop = img[y0 - 6:y0 + 70, x0 - 6:x0 + 134].cpu().numpy()
    exact(blurred[y0:y0 + 64, x0:x0 + 128].cpu().numpy(), oracle.gaussian_blur(crop, 2.0)[6:70, 6:134], "4K fused window")
    exact(blurred[:40, :200].cpu().numpy(), oracle.gaussian_blur(img[:46, :206].cpu().numpy(), 2.0)[:40, :200], "4K fused corner")


def test_gaussian_selection_mask(eng, oracle):
    rng = np.random.default_rng(2)
    w, h = 300, 200
    img = fx.random_rgba(rng, w, h)
    for mask in (np.zeros((h, w), np.uint8), np.full((h, w), 255, np.uint8), None):
        if mask is None:
            mask = np.zeros((h, w), np.uint8)
            mask[40:90, 100:260] = rng.integers(0, 2, (50, 160), dtype=np.uint8) * 200
            mask[199, 0] = 1
        exact(eng.gaussian_blur(img, 4.0, mask=mask, exact=True), oracle.gaussian_blur(img, 4.0, mask=mask))
    small = np.zeros((h, w), np.uint8)
    small[100:110, 150:160] = 255
    exact(eng.gaussian_blur(img, 6.0, mask=small, exact=True), oracle.gaussian_blur(img, 6.0, mask=small))


@pytest.mark.parametrize("w,h", SIZES + [(700, 300)])
def test_box_motion_median_vignette_sharpen_random(eng, oracle, w, h):
    rng = np.random.default_rng(w * 7 + h)
    img = fx.random_rgba(rng, w, h)
    mask = (rng.random((h, w)) < 0.7).astype(np.uint8) * 255
    for r in (0.4, 0.5, 1.0, 3.0, 10.5, 40.0):
        exact(eng.box_blur(img, r), oracle.box_blur(img, r), f"box r={r}")
    exact(eng.box_blur(img, 3.0, mask=mask), oracle.box_blur(img, 3.0, mask=mask), "box mask")
    for ang, dist in ((45.0, 10.0), (0.0, 0.5), (90.0, 1.0), (-30.0, 25.5), (200.0, 3.0)):
        exact(eng.motion_blur(img, ang, dist), oracle.motion_blur(img, ang, dist), f"motion {ang},{dist}")
    exact(eng.motion_blur(img, 30.0, 7.0, mask=mask), oracle.motion_blur(img, 30.0, 7.0, mask=mask), "motion mask")
    for r in (0, 1, 2, 3, 7):
        exact(eng.median(img, r), oracle.median(img, r), f"median r={r}")
    exact(eng.median(img, 2, mask=mask), oracle.median(img, 2, mask=mask), "median mask")
    for a, s in ((0.8, 0.5), (0.0, 0.3), (1.5, 0.0), (0.5, 2.0)):
        exact(eng.vignette(img, a, s), oracle.vignette(img, a, s), f"vignette {a},{s}")
    exact(eng.vignette(img, 0.5, 0.3, mask=mask), oracle.vignette(img, 0.5, 0.3, mask=mask), "vignette mask")
    for a, r in ((1.0, 1.0), (0.0, 2.0), (2.5, 3.0), (1.0, 0.0)):
        exp = oracle.sharpen(img, a, r)
        exact(eng.sharpen(img, a, r, exact=True), exp, f"sharpen {a},{r}")
        d = np.abs(eng.sharpen(img, a, r).astype(int) - exp.astype(int)).max()
        assert d <= int(np.ceil(abs(a))) + 1  # a +-1 blur level moves the result by <= amount (+ rounding)
    exact(eng.sharpen(img, 1.0, 2.0, mask=mask, exact=True), oracle.sharpen(img, 1.0, 2.0, mask=mask), "sharpen mask")


def test_median_large_radius(eng, oracle):
    rng = np.random.default_rng(9)
    img = fx.random_rgba(rng, 90, 70)
    for r in (12, 21, 32, 33, 40):  # 32 = last radius of the column-histogram kernel (65 pixels per column), 33 = first of the bisection kernel
        exact(eng.median(img, r), oracle.median(img, r), f"median r={r}")


def test_median_kernels_agree(eng, oracle):
    """Every radius class of pfe_median (forgetful selection r <= 2, column histograms r <= 32, bisection above) against
    the oracle on images with flat areas, ramps and noise (ties are where rank selection goes wrong), at sizes that
    are not multiples of the 32-column strips or the row segments, with a mask; then the forced bisection kernel."""
    import os

    rng = np.random.default_rng(31)
    w, h = 203, 171
    img = fx.random_rgba(rng, w, h)
    img[20:90, 30:120] = (17, 200, 0, 255)                    # flat block: every window value equal
    img[100:160, :, 0] = np.arange(w, dtype=np.uint8)[None]   # ramp
    img[:, 150:, 3] = rng.integers(0, 3, (h, w - 150), dtype=np.uint8) * 127  # three-valued channel
    mask = (rng.random((h, w)) < 0.5).astype(np.uint8) * 255
    for r in (1, 2, 3, 4, 5, 9, 16):
        exp = oracle.median(img, r)
        exact(eng.median(img, r), exp, f"median r={r}")
        for forced in ("bisect", "hist16"):  # hist16: the column-histogram kernel with 16-bit counters (8-bit is the default)
            os.environ["PFE_MEDIAN_KERNEL"] = forced
            try:
                exact(eng.median(img, r), exp, f"median r={r} ({forced} kernel)")
            finally:
                del os.environ["PFE_MEDIAN_KERNEL"]
    for r in (1, 3, 6):
        exact(eng.median(img, r, mask=mask), oracle.median(img, r, mask=mask), f"median mask r={r}")


def test_median_radius_above_127(eng, oracle):
    """The reference sorts any window (noise.rs:357-410); radii above the 16-bit counters' range take the global
    kernel with 32-bit counts."""
    rng = np.random.default_rng(3)
    img = fx.random_rgba(rng, 37, 29)
    exact(eng.median(img, 130), oracle.median(img, 130), "median r=130")


def test_adjust_all_ops_random(eng, oracle):
    rng = np.random.default_rng(21)
    for (w, h) in [(67, 45), (256, 192), (1, 1)]:
        img = fx.random_rgba(rng, w, h)
        img[0, 0] = (10, 10, 10, 255)  # grey pixel: HSL's achromatic branch
        mask = (rng.random((h, w)) < 0.6).astype(np.uint8) * 255
        lut1 = rng.integers(0, 256, 256, dtype=np.uint8)
        lut4 = rng.integers(0, 256, (4, 256), dtype=np.uint8)
        cases = [(oracle.INVERT, (), None), (oracle.INVERT_ALPHA, (), None), (oracle.SEPIA, (), None),
                 (oracle.DESATURATE, (), None), (oracle.BRIGHTNESS_CONTRAST, (30.0, 20.0), None),
                 (oracle.BRIGHTNESS_CONTRAST, (-100.0, 100.0), None), (oracle.HSL, (30.0, -20.0, 10.0), None),
                 (oracle.HSL, (-170.0, 80.0, -30.0), None), (oracle.HSL, (0.0, 0.0, 0.0), None),
                 (oracle.EXPOSURE, (2.0,), None), (oracle.EXPOSURE, (0.3,), None), (oracle.LUT_RGB, (), lut1),
                 (oracle.LUT_RGBA, (), lut4), (oracle.TEMPERATURE_TINT, (30.0, 10.0), None),
                 (oracle.HIGHLIGHTS_SHADOWS, (30.0, -20.0), None), (oracle.THRESHOLD, (128.0,), None),
                 (oracle.THRESHOLD, (0.0,), None), (oracle.POSTERIZE, (4.0,), None), (oracle.POSTERIZE, (2.0,), None),
                 (oracle.POSTERIZE, (16.0,), None), (oracle.COLOR_BALANCE, (10.0, 0.0, -10.0, 0.0, 0.0, 0.0, -10.0, 0.0, 10.0), None),
                 (oracle.COLOR_BALANCE, (-100.0, 50.0, 100.0, 30.0, -60.0, 90.0, 100.0, -100.0, 5.0), None),
                 (oracle.GRADIENT_MAP, (), lut4.T.copy()), (oracle.BLACK_AND_WHITE, (0.3, 0.59, 0.11), None),
                 (oracle.BLACK_AND_WHITE, (200.0, 150.0, 40.0), None), (oracle.VIBRANCE, (0.5,), None),
                 (oracle.VIBRANCE, (-1.0,), None), (oracle.VIBRANCE, (1.0,), None), (oracle.S_INVERT, (), None),
                 (oracle.S_DESATURATE, (), None), (oracle.S_SEPIA, (), None), (oracle.S_SEPIA_STRENGTH, (0.4,), None),
                 (oracle.S_BRIGHTNESS_CONTRAST, (20.0, 10.0), None), (oracle.S_HSL, (10.0, 15.0, 0.0), None),
                 (oracle.S_HSL, (-200.0, -50.0, 20.0), None), (oracle.S_EXPOSURE, (1.7,), None),
                 (oracle.S_LUT_RGB, (), lut1)]
        for op, params, luts in cases:
            exact(eng.adjust(img, op, params, luts=luts), oracle.adjust(img, op, params, luts=luts), f"op {op} {params}")
            exact(eng.adjust(img, op, params, luts=luts, mask=mask), oracle.adjust(img, op, params, luts=luts, mask=mask),
                  f"op {op} masked")
        _, occ = oracle.tiled_roundtrip(img)
        occ[0, 0] = 0
        exact(eng.adjust(img, oracle.INVERT, occupancy=occ), oracle.adjust(img, oracle.INVERT, occupancy=occ), "occupancy")
    mm = eng.channel_minmax(img, mask=mask)
    sel = (mask > 0) & (img[..., 3] != 0)
    exp = [f(img[..., c][sel]) for c in range(3) for f in (np.min, np.max)] if sel.any() else [255, 0] * 3
    assert list(mm) == [int(v) for v in exp]


def test_adjust_exhaustive_hsl_inputs(eng, oracle):
    """Every (r,g,b) on a 64-step lattice plus random triples, both HSL variants."""
    g = np.arange(0, 256, 5, dtype=np.uint8)
    r_, g_, b_ = np.meshgrid(g, g, g, indexing="ij")
    n = r_.size
    w = 512
    h = (n + w - 1) // w
    img = np.zeros((h * w, 4), np.uint8)
    img[:n, 0], img[:n, 1], img[:n, 2] = r_.ravel(), g_.ravel(), b_.ravel()
    img[:, 3] = 255
    img = img.reshape(h, w, 4)
    for op in (oracle.HSL, oracle.S_HSL):
        for p in ((30.0, -20.0, 10.0), (123.0, 40.0, -5.0), (-359.0, 100.0, 0.0)):
            exact(eng.adjust(img, op, p), oracle.adjust(img, op, p), f"hsl {op} {p}")
    for v in (0.5, -0.5, 0.0, 1.0):
        exact(eng.adjust(img, oracle.VIBRANCE, (v,)), oracle.adjust(img, oracle.VIBRANCE, (v,)), f"vibrance {v}")
    exact(eng.adjust(img, oracle.COLOR_BALANCE, (40.0, -10.0, 5.0, 1.0, 2.0, 3.0, -30.0, 20.0, 60.0)),
          oracle.adjust(img, oracle.COLOR_BALANCE, (40.0, -10.0, 5.0, 1.0, 2.0, 3.0, -30.0, 20.0, 60.0)), "colour balance lattice")


def test_warps_random(eng, oracle):
    rng = np.random.default_rng(4)
    for (w, h) in [(67, 45), (256, 192)]:
        src = fx.random_rgba(rng, w, h)
        disp = rng.normal(0, 6, (h, w, 2)).astype(np.float32)
        disp[0, 0] = (1e9, -1e9); disp[1, 1] = (np.nan, 0.0); disp[2, 2] = (0.5, w + 5.0)
        exact(eng.warp_displacement(src, disp), oracle.warp_displacement(src, disp), "warp")
        # source of a different size than the field
        src2 = fx.random_rgba(rng, w // 2 + 1, h + 3)
        exact(eng.warp_displacement(src2, disp), oracle.warp_displacement(src2, disp), "warp src!=dst size")
        for cols, rows in ((2, 2), (6, 6), (1, 3), (15, 15)):
            orig = fx.uniform_grid(cols, rows, float(w), float(h))
            deformed = (orig + rng.normal(0, 4, orig.shape)).astype(np.float32)
            d_full = oracle.mesh_displacement(orig, deformed, cols, rows, w, h)
            assert np.array_equal(eng.mesh_displacement(orig, deformed, cols, rows, w, h), d_full)
            d_fast = oracle.mesh_displacement(None, deformed, cols, rows, w, h)
            assert np.array_equal(eng.mesh_displacement(None, deformed, cols, rows, w, h), d_fast)
            exact(eng.mesh_warp(src, orig, deformed, cols, rows, w, h), oracle.mesh_warp(src, orig, deformed, cols, rows, w, h),
                  f"mesh {cols}x{rows}")


def test_liquify_random(eng, oracle):
    rng = np.random.default_rng(6)
    w, h = 200, 150
    fo = np.zeros((h, w, 2), np.float32)
    fe = np.zeros((h, w, 2), np.float32)
    for i in range(24):
        kind = i % 4
        cx, cy = float(rng.uniform(-10, w + 10)), float(rng.uniform(-10, h + 10))
        r, s = float(rng.uniform(0.5, 60)), float(rng.uniform(0.1, 1.0))
        a0, a1 = (float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5))) if kind == 0 else (float(i % 8 < 4), 0.0)
        assert eng.liquify(fe, kind, cx, cy, r, s, a0, a1) == oracle.liquify(fo, kind, cx, cy, r, s, a0, a1)
    # device exp is the correctly rounded value; libm expf may differ from it by one ulp on rare inputs
    assert np.allclose(fe, fo, rtol=3e-7, atol=1e-7)
    assert (fe != fo).mean() < 0.02
    src = fx.random_rgba(rng, w, h)
    within1(eng.warp_displacement(src, fe), oracle.warp_displacement(src, fo), "liquify warp")


def test_brush_random_strokes(eng, oracle):
    rng = np.random.default_rng(8)
    w, h = 200, 150
    for trial in range(12):
        size, hard = float(rng.uniform(1, 50)), float(rng.uniform(-0.2, 1.2))
        aa, eraser = bool(trial & 1), bool(trial & 2)
        color = tuple(float(v) for v in rng.uniform(0, 1, 4))
        flow = float(rng.choice([1.0, 0.5, 0.02]))
        img_o = fx.random_rgba(rng, w, h) if trial % 3 else fx.solid(w, h, (0, 0, 0, 0))
        img_e = img_o.copy()
        sel = None if trial % 4 else (rng.random((h, w)) < 0.5).astype(np.uint8)
        bo = oracle.make_brush(size, hard, aa, color, flow=flow, is_eraser=eraser)
        be = eng.brush_desc(size, hard, aa, color, flow=flow, is_eraser=eraser)
        pts = rng.uniform(-20, 220, (5, 2))
        centres = []
        for a, b in zip(pts[:-1], pts[1:]):
            seg = (float(a[0]), float(a[1]), float(b[0]), float(b[1]))
            oracle.brush_line(img_o, bo, *seg, sel_mask=sel)
            centres.append(eng.brush_line_centres(w, h, *seg))
        centres = np.concatenate(centres) if centres else np.zeros((0, 2), np.float32)
        eng.brush_stamps(img_e, be, centres, selection_mask=sel)
        exact(img_e, img_o, f"stroke {trial}")


# ---------------------------------------------------------------------------------------------
# device tier == host tier; chained ops
# ---------------------------------------------------------------------------------------------
def test_device_tier_matches_host_tier(eng, oracle):
    import torch
    from paintfe_b200.engine import make_layer

    rng = np.random.default_rng(12)
    w, h = 320, 200
    ol, el = _stack(rng, oracle, w, h, 9, masks=True, adj=True)
    dev = [make_layer(None if L["rgba"] is None else torch.from_numpy(L["rgba"]).cuda(),
                      opacity=L["opacity"], blend=L["blend"], visible=L["visible"],
                      mask=None if L["mask"] is None else torch.from_numpy(L["mask"]).cuda(), kind=L["kind"], adj=L["adj"])
           for L in el]
    if all(L["rgba"] is None for L in el):
        pytest.skip("no raster layer drawn")
    flat_dev = eng.flatten(dev, w, h)
    exp = oracle.flatten(ol, w, h)
    exact(flat_dev.cpu().numpy(), exp, "dev flatten")
    blur_dev = eng.gaussian_blur(flat_dev, 3.0, exact=True)
    exact(blur_dev.cpu().numpy(), oracle.gaussian_blur(exp, 3.0), "dev gaussian")
    hsl_dev = eng.adjust(blur_dev, oracle.S_HSL, (10.0, 15.0, 0.0))
    exact(hsl_dev.cpu().numpy(), oracle.adjust(oracle.gaussian_blur(exp, 3.0), oracle.S_HSL, (10.0, 15.0, 0.0)), "dev hsl")
    exact(eng.flatten_gaussian(el, w, h, 3.0, exact=True), oracle.gaussian_blur(exp, 3.0), "fused host call")
    launches = eng.launches
    assert launches > 0


# ---------------------------------------------------------------------------------------------
# full-size (BASELINE config) properties
# ---------------------------------------------------------------------------------------------
def test_full_size_8k_flatten_and_gaussian_crops(eng, oracle):
    """7680x4320, 16 layers cycling modes (config 2) + Gaussian sigma=20 (headline): since flatten is
    per-pixel and the blur has finite support, any crop can be re-derived by the oracle from the
    same crop (+halo) of the inputs."""
    import torch
    from paintfe_b200.engine import make_layer

    w, h, n = 7680, 4320, 16
    g = torch.Generator(device="cuda").manual_seed(0x5EED)
    layers = [torch.randint(0, 256, (h, w, 4), dtype=torch.uint8, device="cuda", generator=g) for _ in range(n)]
    meta = [dict(blend=i % 25, opacity=0.25 + 0.05 * i) for i in range(n)]
    flat = eng.flatten([make_layer(t, **m) for t, m in zip(layers, meta)], w, h)
    blur = eng.gaussian_blur(flat, 20.0, exact=True)
    blur_fast = eng.gaussian_blur(flat, 20.0)
    r = 60
    rng = np.random.default_rng(1)
    crops = [(0, 0), (w - 256, h - 256), (w - 256, 0), (0, h - 256)] + [(int(rng.integers(0, w - 256)), int(rng.integers(0, h - 256))) for _ in range(3)]
    for (cx, cy) in crops:
        x0, y0, x1, y1 = max(cx - r, 0), max(cy - r, 0), min(cx + 256 + r, w), min(cy + 256 + r, h)
        sub = [oracle.make_layer(t[y0:y1, x0:x1].cpu().numpy(), **m) for t, m in zip(layers, meta)]
        exp_flat = oracle.flatten(sub, x1 - x0, y1 - y0)
        exact(flat[y0:y1, x0:x1].cpu().numpy(), exp_flat, f"flatten crop {cx},{cy}")
        exp_blur = oracle.gaussian_blur(exp_flat, 20.0)
        # pixels whose support lies inside the halo'd crop (or is clamped by a true image edge)
        ix0, iy0 = cx - x0, cy - y0
        got = blur[cy:cy + 256, cx:cx + 256].cpu().numpy()
        exact(got, exp_blur[iy0:iy0 + 256, ix0:ix0 + 256], f"gaussian crop {cx},{cy}")
        within1(blur_fast[cy:cy + 256, cx:cx + 256].cpu().numpy(), exp_blur[iy0:iy0 + 256, ix0:ix0 + 256])
    # idempotence-style property: flattening [flat] alone with Normal/1.0 returns flat where alpha==255
    again = eng.flatten([make_layer(flat)], w, h)
    opaque = flat[..., 3] == 255
    assert torch.equal(again[opaque], flat[opaque])


# ---------------------------------------------------------------------------------------------
# CLI batch mode end to end (BASELINE config 1 and a small config-5 style batch)
# ---------------------------------------------------------------------------------------------
def test_cli_flatten_pfe_and_script_batch(eng, oracle, tmp_path):
    from PIL import Image

    from paintfe_b200 import cli, pfe_io
    from test_pfe_io import _config1_project

    proj, _ = _config1_project()
    (tmp_path / "in").mkdir()
    (tmp_path / "in" / "project.pfe").write_bytes(pfe_io.save_pfe_v1(proj))
    rng = np.random.default_rng(3)
    shots = {}
    for i in range(3):
        img = fx.random_rgba(rng, 200 + 17 * i, 120 + 5 * i)
        shots[f"shot{i}"] = img
        Image.fromarray(img, "RGBA").save(tmp_path / "in" / f"shot{i}.png")
    script = tmp_path / "process.rhai"
    script.write_text("apply_blur(4.0); apply_hsl(10.0, 15.0, 0.0); apply_vignette(0.5, 0.3);\n")
    # config 1: --flatten on the 2-layer project, no script
    assert cli.main(["-i", str(tmp_path / "in" / "project.pfe"), "-o", str(tmp_path / "flat.png")]) == 0
    flats = [L.to_flat(1024, 1024) for L in proj.layers]
    exp = oracle.flatten([oracle.make_layer(f, opacity=L.opacity, blend=L.blend_mode) for f, L in zip(flats, proj.layers)], 1024, 1024)
    exact(np.array(Image.open(tmp_path / "flat.png").convert("RGBA")), exp, "cli --flatten project.pfe")
    # a v3 project: sparse layers, a hidden folder, adjustment layers - flattened from its chunk tables
    w3, h3 = 333, 200
    bg, top, hid = fx.random_rgba(rng, w3, h3, alpha="opaque"), fx.random_rgba(rng, w3, h3), fx.random_rgba(rng, w3, h3)
    top[:128, :192] = 0
    v3_layers = [pfe_io.layer_from_flat("bg", bg), pfe_io.layer_from_flat("hidden", hid),
                 pfe_io.PfeLayer("invert", True, 0.6, 0, adjustment=(3, ())),
                 pfe_io.layer_from_flat("top", top, opacity=0.8, blend_mode=21),
                 pfe_io.PfeLayer("exposure", True, 0.5, 0, adjustment=(1, (0.7,)))]
    v3_layers[1].folder_id = 7
    v3 = pfe_io.PfeProject(w3, h3, 0, v3_layers, folders=[dict(id=7, name="off", visible=False, collapsed=False,
                                                                insert_above_layer=None, color_index=None)])
    (tmp_path / "v3.pfe").write_bytes(pfe_io.save_pfe_v3(v3))
    assert cli.main(["-i", str(tmp_path / "v3.pfe"), "-o", str(tmp_path / "v3.png")]) == 0
    occ = v3_layers[0].occupancy(w3, h3) | v3_layers[3].occupancy(w3, h3)
    exp3 = oracle.flatten([oracle.make_layer(bg), oracle.make_layer(hid, visible=False), oracle.make_layer(None, opacity=0.6, kind=3),
                           oracle.make_layer(v3_layers[3].to_flat(w3, h3), opacity=0.8, blend=21),
                           oracle.make_layer(None, opacity=0.5, kind=1, adj=(float(np.float32(2.0) ** np.float32(0.7)),))],
                          w3, h3, active=occ)
    exact(np.array(Image.open(tmp_path / "v3.png").convert("RGBA")), exp3, "cli --flatten v3 project")
    # batch with a script over a glob; one unreadable file must not stop the batch (cli.rs:204-215)
    (tmp_path / "in" / "broken.png").write_bytes(b"not a png")
    rc = cli.main(["-i", str(tmp_path / "in" / "*.png"), "--script", str(script), "--output-dir", str(tmp_path / "out"), "--exact"])
    assert rc == 1
    for name, img in shots.items():
        e = oracle.gaussian_blur(img, 4.0)
        e = oracle.adjust(e, oracle.S_HSL, (10.0, 15.0, 0.0))
        e = oracle.vignette(e, 0.5, 0.3)
        exact(np.array(Image.open(tmp_path / "out" / f"{name}.png").convert("RGBA")), e, name)
    # the same batch one file at a time gives the same files; JPEG output drops alpha instead of failing (cli.rs:277-304)
    rc = cli.main(["-i", str(tmp_path / "in" / "shot*.png"), "--script", str(script), "--output-dir", str(tmp_path / "out1"),
                   "--exact", "--no-pipeline"])
    assert rc == 0
    for name in shots:
        exact(np.array(Image.open(tmp_path / "out1" / f"{name}.png")), np.array(Image.open(tmp_path / "out" / f"{name}.png")), name)
    assert cli.main(["-i", str(tmp_path / "in" / "shot*.png"), "--script", str(script), "--output-dir", str(tmp_path / "outj"), "-f", "jpg"]) == 0
    assert Image.open(tmp_path / "outj" / "shot0.jpg").mode == "RGB"
    # raster in, project out
    assert cli.main(["-i", str(tmp_path / "in" / "shot1.png"), "-o", str(tmp_path / "shot1.pfe")]) == 0
    back = pfe_io.load_pfe(str(tmp_path / "shot1.pfe"))
    exact(back.layers[0].to_flat(back.width, back.height), shots["shot1"], "png -> pfe")


def test_image_pipeline_keeps_order_and_contents(eng, oracle):
    """ImagePipeline (upload / script / download on three streams, 3 images in flight): eleven images of four different
    sizes, pageable and pinned inputs, results equal to running the script on each image alone."""
    import torch

    from paintfe_b200.pipeline import ImagePipeline
    from paintfe_b200.script import execute_script_sync

    script = "apply_blur(2.0); apply_brightness_contrast(10.0, 20.0); apply_invert();"
    rng = np.random.default_rng(8)
    sizes = [(64, 48), (130, 70), (33, 200), (256, 256)]
    imgs = [fx.random_rgba(rng, *sizes[i % 4]) for i in range(11)]
    pipe = ImagePipeline(eng, depth=3)
    got = {}

    def work(dev):
        return execute_script_sync(eng, script, dev, exact=True)

    for i, im in enumerate(imgs):
        if pipe.full():
            tag, out = pipe.collect()
            got[tag] = out.copy()
        src = im if i % 2 else torch.from_numpy(im).pin_memory()
        pipe.submit(src, work, tag=i)
    for tag, out in pipe.drain():
        got[tag] = out.copy()
    assert sorted(got) == list(range(11))
    for i, im in enumerate(imgs):
        e = oracle.adjust(oracle.adjust(oracle.gaussian_blur(im, 2.0), oracle.S_BRIGHTNESS_CONTRAST, (10.0, 20.0)), oracle.S_INVERT)
        exact(got[i], e, f"pipeline image {i}")
    assert pipe.h2d_bytes == pipe.d2h_bytes == sum(im.size for im in imgs)


# ---------------------------------------------------------------------------------------------
# widened scope: glow / pixelate / bulge / twist / noise / bilateral
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("w,h", [(64, 64), (67, 45), (1, 1), (300, 200)])
def test_widened_effects_random(eng, oracle, w, h):
    rng = np.random.default_rng(w * 13 + h)
    img = fx.random_rgba(rng, w, h)
    mask = (rng.random((h, w)) < 0.7).astype(np.uint8) * 255
    for r, i in ((3.0, 0.5), (0.0, 1.0), (6.5, 2.0)):
        exact(eng.glow(img, r, i, exact=True), oracle.glow(img, r, i), f"glow {r},{i}")
        d = np.abs(eng.glow(img, r, i).astype(int) - oracle.glow(img, r, i).astype(int)).max()
        assert d <= int(np.ceil(i)) + 1
    exact(eng.glow(img, 2.0, 0.7, mask=mask, exact=True), oracle.glow(img, 2.0, 0.7, mask=mask), "glow mask")
    for bs in (0, 1, 2, 8, 33, 1000):
        exact(eng.pixelate(img, bs), oracle.pixelate(img, bs), f"pixelate {bs}")
    exact(eng.pixelate(img, 5, mask=mask), oracle.pixelate(img, 5, mask=mask), "pixelate mask")
    for amt, org in ((0.5, (0.5, 0.5)), (-0.8, (0.2, 0.9)), (0.0, (0.5, 0.5)), (3.0, (1.5, -1.0))):
        exact(eng.bulge(img, amt, org), oracle.bulge(img, amt, org), f"bulge {amt},{org}")
    exact(eng.bulge(img, 0.7, mask=mask), oracle.bulge(img, 0.7, mask=mask), "bulge mask")
    for ang, org in ((45.0, (0.5, 0.5)), (-200.0, (0.1, 0.3)), (0.0, (0.5, 0.5))):
        assert within1(eng.twist(img, ang, org), oracle.twist(img, ang, org), f"twist {ang}") < 0.02
    for t, mono, seed, scale, octv, amt in ((0, False, 42, 1.0, 1, 30.0), (0, True, 7, 3.0, 1, 80.0), (2, False, 42, 5.0, 3, 50.0),
                                           (2, True, 1, 0.05, 9, 100.0), (1, False, 3, 2.0, 1, 40.0)):
        exact(eng.add_noise(img, amt, t, mono, seed, scale, octv), oracle.add_noise(img, amt, t, mono, seed, scale, octv),
              f"noise type {t} mono {mono}")
    assert within1(eng.add_noise(img, 30.0, 1, True, 42, 1.0, 1), oracle.add_noise(img, 30.0, 1, True, 42, 1.0, 1), "gaussian noise") < 0.02
    exact(eng.add_noise(img, 30.0, 0, False, 42, 1.0, 1, mask=mask), oracle.add_noise(img, 30.0, 0, False, 42, 1.0, 1, mask=mask), "noise mask")
    for st, r in ((0.5, 2), (10.0, 1), (50.0, 4), (0.0, 3)):
        assert within1(eng.reduce_noise(img, st, r), oracle.reduce_noise(img, st, r), f"bilateral {st},{r}") < 0.02
    assert within1(eng.reduce_noise(img, 20.0, 2, mask=mask), oracle.reduce_noise(img, 20.0, 2, mask=mask), "bilateral mask") < 0.02


def test_effects3_goldens_and_known_answers(eng):
    """tests/visual_filters.rs:193-236, :338-352 through the engine."""
    from test_oracle_golden import _square

    o = _EngAsOracle(eng, True)
    exact(o.drop_shadow(_square((255, 255, 255, 255)), 5, 5, 3.0, False, (0, 0, 0, 255), 0.8), fx.golden("filters", "drop_shadow"))
    within1(_EngAsOracle(eng, False).drop_shadow(_square((255, 255, 255, 255)), 5, 5, 3.0, False, (0, 0, 0, 255), 0.8),
            fx.golden("filters", "drop_shadow"))
    exact(o.outline(_square((255, 0, 0, 255)), 2, (0, 0, 255, 255), o.OUTLINE_OUTSIDE, True), fx.golden("filters", "outline_outside"))
    out = eng.canvas_border(fx.solid(8, 8, (10, 20, 30, 255)), 2, (200, 100, 50, 255))
    assert tuple(out[0, 0]) == (200, 100, 50, 255) and tuple(out[3, 3]) == (10, 20, 30, 255)
    g = fx.gradient(64, 64)
    exact(eng.color_filter(g, (255, 255, 255, 255), 0.0, 0), g, "colour filter identity")


@pytest.mark.parametrize("w,h", [(64, 64), (67, 45), (1, 1), (300, 200)])
def test_effects3_random(eng, oracle, w, h):
    """Every effect of effects3.cu against the oracle on random images (so alpha, clamping and the mask
    rule are exercised), ragged sizes, parameter extremes."""
    rng = np.random.default_rng(w * 17 + h)
    img = fx.random_rgba(rng, w, h)
    mask = (rng.random((h, w)) < 0.7).astype(np.uint8) * 255
    sparse = img.copy()
    sparse[rng.random((h, w)) < 0.6] = 0  # holes in alpha for shadow / outline
    col = (200, 40, 90, 180)

    def both(name, *a, src=img):
        exact(getattr(eng, name)(src, *a), getattr(oracle, name)(src, *a), f"{name}{a}")
        exact(getattr(eng, name)(src, *a, mask=mask), getattr(oracle, name)(src, *a, mask=mask), f"{name}{a} mask")

    for es, th in ((1.0, 0.5), (50.0, 3.0), (0.0, 0.0)):
        both("ink", es, th)
    for r, lv in ((3, 20), (1, 2), (10, 64), (0, 0), (25, 300)):
        both("oil_painting", r, lv)
    for mode in range(4):
        both("color_filter", col, 0.6, mode)
    both("color_filter", (255, 128, 0, 255), 1.5, 3)
    both("contours", 10.0, 5.0, 1.0, (0, 0, 0, 255), 42, 2, 0.5)
    both("contours", 0.1, 0.2, 6.0, col, 7, 9, 1.0)
    for cs, seed in ((16.0, 42), (2.0, 1), (0.5, 3), (7.3, 99), (500.0, 5)):
        both("crystallize", cs, seed)
    for args in ((20.0, 10.0, 42, 2, 0.5, False, False), (5.0, 30.0, 7, 4, 0.8, True, False), (8.0, 50.0, 3, 1, 0.3, True, True),
                 (0.1, 5.0, 1, 9, 0.5, False, True)):
        both("dents", *args)
    for shape in range(4):
        both("halftone", 4.0, 45.0, shape)
    both("halftone", 0.5, -30.0, 0)
    for r in (5.0, 0.4, 0.5, 1.0, 2.7, 12.0):
        both("bokeh_blur", r)
    both("zoom_blur", 0.5, 0.5, 0.3, 8)
    both("zoom_blur", 0.2, 0.9, 5.0, 1, (1.0, 0.5, 0.0, 1.0), 0.7)
    both("zoom_blur", 0.5, 0.5, 0.0005, 16)
    both("grid", 16, 16, 1, (0, 0, 0, 255), 0, 1.0)
    both("grid", 0, 5, 0, col, 1, 0.4)
    both("grid", 7, 3, 2, col, 0, 0.5)
    for bw in (0, 2, 1000):
        both("canvas_border", bw, col)
    for args in ((5, 5, 3.0, False, (0, 0, 0, 255), 0.8), (-3, 7, 2.0, True, col, 1.0), (0, 0, 0.4, False, col, 0.5),
                 (1000, 0, 1.0, True, col, 0.9), (2, -2, 0.2, True, col, 0.7)):
        exact(eng.drop_shadow(sparse, *args, exact=True), oracle.drop_shadow(sparse, *args), f"drop_shadow{args}")
        exact(eng.drop_shadow(sparse, *args, mask=mask, exact=True), oracle.drop_shadow(sparse, *args, mask=mask), f"drop_shadow{args} mask")
    for args in ((2, (0, 0, 255, 255), 0, True), (1, col, 1, False), (4, col, 2, True), (0, col, 2, False)):
        both("outline", *args, src=sparse)
    both("outline", 2, col, 0, True, src=np.zeros_like(img))  # nothing opaque: clone
    for args in ((42, 50.0, 20, 0.0), (7, 100.0, 1, 45.0), (3, 0.0, 9, 90.0), (9, 80.0, 500, 200.0)):
        both("pixel_drag", *args)
    both("rgb_displace", (5, 0), (0, 0), (-5, 0))
    both("rgb_displace", (-1000, 3), (2, -2), (0, 100000))


def test_effects3_device_tier_and_large(eng, oracle):
    """Device-pointer tier equals the host tier; an image larger than one block grid row / several cells."""
    import torch

    rng = np.random.default_rng(5)
    img = fx.random_rgba(rng, 1031, 517)
    dev = torch.from_numpy(img).cuda()
    for name, a in (("ink", (2.0, 0.5)), ("oil_painting", (4, 32)), ("crystallize", (23.0, 42)), ("bokeh_blur", (9.5,)),
                    ("halftone", (6.0, 15.0, 0)), ("zoom_blur", (0.4, 0.6, 0.5, 32)), ("dents", (20.0, 10.0, 42, 3, 0.5, True, True)),
                    ("pixel_drag", (42, 50.0, 40, 10.0)), ("contours", (30.0, 8.0, 2.0, (10, 20, 30, 200), 1, 4, 0.8))):
        want = getattr(oracle, name)(img, *a)
        exact(getattr(eng, name)(dev, *a).cpu().numpy(), want, f"{name} device tier")
        exact(getattr(eng, name)(img, *a), want, f"{name} host tier")
    exact(eng.drop_shadow(dev, 9, -4, 6.0, True, (5, 5, 5, 255), 0.9, exact=True).cpu().numpy(),
          oracle.drop_shadow(img, 9, -4, 6.0, True, (5, 5, 5, 255), 0.9), "drop_shadow device tier")
    exact(eng.outline(dev, 3, (1, 2, 3, 255), 2, True).cpu().numpy(), oracle.outline(img, 3, (1, 2, 3, 255), 2, True), "outline device")


# ---------------------------------------------------------------------------------------------
# geometry: flips / quarter turns / resize_canvas / affine / imageops::resize
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(GEOMETRY))
def test_geometry_golden(eng, oracle, name):
    out = GEOMETRY[name](_EngAsOracle(eng), fx.gradient(64, 48))
    if name != "flatten_single":
        out, _ = oracle.tiled_roundtrip(out)
    exact(out, fx.golden("transforms", name), name)


def test_geometry_more_goldens(eng):
    from paintfe_b200.engine import make_layer

    src = fx.gradient(32, 32)
    for name, rz, sc in (("affine_rotate_90", float(np.float32(np.pi / 2)), 1.0), ("affine_scale_half", 0.0, 0.5)):
        exact(eng.flatten([make_layer(eng.affine(src, 32, 32, rz, scale=sc))], 32, 32), fx.golden("transform", name), name)
    g = fx.gradient(64, 64)
    exact(eng.orient(g, 0), fx.golden("scripting", "flip_horizontal"))
    exact(eng.orient(g, 1), fx.golden("scripting", "flip_vertical"))


@pytest.mark.parametrize("w,h", [(64, 48), (67, 45), (1, 1), (1, 9), (300, 200), (33, 1031)])
def test_geometry_random(eng, oracle, w, h):
    import torch

    rng = np.random.default_rng(w * 19 + h)
    img = fx.random_rgba(rng, w, h)
    dev = torch.from_numpy(img).cuda()
    for op in range(5):
        want = oracle.orient(img, op)
        exact(eng.orient(img, op), want, f"orient {op}")
        exact(eng.orient(dev, op).cpu().numpy(), want, f"orient {op} device tier")
    for nw, nh, anchor in ((w + 32, h + 32, (1, 1)), (w + 16, h + 16, (0, 0)), (max(w // 2, 1), max(h // 2, 1), (2, 2)),
                           (w + 5, max(h - 3, 1), (1, 2)), (max(w - 7, 1), h + 9, (2, 0))):
        exact(eng.resize_canvas(img, nw, nh, anchor, (9, 8, 7, 200)), oracle.resize_canvas(img, nw, nh, anchor, (9, 8, 7, 200)),
              f"resize_canvas {nw}x{nh} {anchor}")
    for args in ((0.0, 0.0, 0.0, 1.0, (0.0, 0.0)), (33.0, 0.0, 0.0, 1.0, (0.0, 0.0)), (-120.0, 20.0, -35.0, 0.7, (5.5, -3.25)),
                 (90.0, 0.0, 0.0, 2.5, (0.0, 0.0)), (10.0, 89.0, 0.0, 1.0, (0.0, 0.0)), (0.0, 0.0, 0.0, 0.0, (1.0, 1.0))):
        for nearest in (False, True):
            exact(eng.affine(img, w, h, *args, nearest=nearest), oracle.affine(img, w, h, *args, nearest=nearest), f"affine {args} {nearest}")
    exact(eng.affine(dev, w + 11, h + 3, 17.0, 5.0, 5.0, 1.2, (2.0, 2.0)).cpu().numpy(),
          oracle.affine(img, w + 11, h + 3, 17.0, 5.0, 5.0, 1.2, (2.0, 2.0)), "affine other canvas, device tier")
    for nw, nh in ((w * 2, h * 2), (max(w // 2, 1), max(h // 2, 1)), (w + 3, max(h - 1, 1)), (max(w // 7, 1), h * 3), (w, h), (1, 1)):
        for f in range(4):
            exact(eng.resize(img, nw, nh, f), oracle.resize(img, nw, nh, f), f"resize {nw}x{nh} filter {f}")
    exact(eng.resize(dev, w + 9, h + 4, 3).cpu().numpy(), oracle.resize(img, w + 9, h + 4, 3), "resize device tier")


# ---------------------------------------------------------------------------------------------
# tile-native flatten and the device-resident TiledImage
# ---------------------------------------------------------------------------------------------
def _sparse_stack(rng, w, h, n):
    """Random layers with holes: whole chunks cleared, plus some chunks that are populated but fully
    transparent in one layer only."""
    cyn, cxn = (h + 63) // 64, (w + 63) // 64
    imgs, specs = [], []
    for i in range(n):
        im = fx.random_rgba(rng, w, h)
        keep = rng.random((cyn, cxn)) < (0.5 if i else 0.8)
        big = np.kron(keep, np.ones((64, 64), bool))[:h, :w]
        im[~big] = 0
        imgs.append(im)
    return imgs


@pytest.mark.parametrize("w,h,n", [(64, 64, 2), (200, 130, 5), (67, 45, 3), (1, 1, 1), (260, 257, 40), (512, 384, 16)])
def test_flatten_tiles_matches_oracle(eng, oracle, w, h, n):
    import torch
    from paintfe_b200.engine import make_layer

    rng = np.random.default_rng(w * 23 + h + n)
    imgs = _sparse_stack(rng, w, h, n)
    cyn, cxn = (h + 63) // 64, (w + 63) // 64
    o_layers, t_layers, d_layers, active = [], [], [], np.zeros((cyn, cxn), np.uint8)
    handles = []
    for i, im in enumerate(imgs):
        meta = dict(blend=int(rng.integers(0, 25)), opacity=float(rng.choice([1.0, rng.uniform(0, 1)])), visible=bool(rng.random() < 0.9))
        if i % 7 == 3:  # adjustment layer: only chunks some raster layer populates are touched
            meta.update(kind=int(rng.integers(1, 5)), adj=tuple(float(v) for v in rng.uniform(-0.5, 1.2, 16)))
            o_layers.append(oracle.make_layer(None, **meta))
            t_layers.append(dict(meta))
            d_layers.append(dict(meta))
            continue
        occ, tiles = eng.flat_to_tiles(im)
        table = [tiles[k] if occ.reshape(-1)[k] else None for k in range(occ.size)]
        mask_img = mask_plane = mask_table = None
        if i % 3 == 1:
            mask_img = fx.random_rgba(rng, w, h)
            mask_img[rng.random((h, w)) < 0.5] = 0
            mocc, mtiles = eng.flat_to_tiles(mask_img)
            mask_table = [mtiles[k] if mocc.reshape(-1)[k] else None for k in range(mocc.size)]
            mask_plane = np.ascontiguousarray(mask_img[..., 3])
        if meta["visible"]:
            active |= occ
        o_layers.append(oracle.make_layer(im, mask=mask_plane, **meta))
        t_layers.append(dict(meta, tiles=table, mask_tiles=mask_table))
        dt = eng.tiled(w, h).upload(table)
        dm = eng.tiled(w, h).from_flat(torch.from_numpy(mask_img).cuda()) if mask_img is not None else None
        handles += [dt, dm]
        d_layers.append(dict(meta, tiles=dt, mask_tiles=dm))
    want = oracle.flatten(o_layers, w, h, active=active)
    exact(eng.flatten_tiles(t_layers, w, h), want, "host tier chunk tables")
    exact(eng.flatten_tiles(d_layers, w, h).cpu().numpy(), want, "device-resident tiles")
    for hnd in handles:
        if hnd is not None:
            hnd.close()


def test_device_tiled_roundtrip(eng, oracle):
    """from_rgba_image / to_rgba_image on the device: occupancy and pixels equal the host marshalling."""
    import torch

    rng = np.random.default_rng(12)
    for (w, h) in [(64, 64), (65, 63), (200, 130), (1, 1), (129, 64)]:
        img = fx.random_rgba(rng, w, h)
        img[: h // 2, : w // 2, 3] = 0  # a fully transparent region with non-zero RGB
        t = eng.tiled(w, h).from_flat(torch.from_numpy(img).cuda())
        exp, exp_occ = oracle.tiled_roundtrip(img)
        occ, tiles = t.download()
        assert np.array_equal(occ.reshape(exp_occ.shape), exp_occ)
        exact(t.to_flat().cpu().numpy(), exp, "to_flat")
        hocc, htiles = eng.flat_to_tiles(img)
        for k in range(occ.size):
            if occ[k]:
                assert np.array_equal(tiles[k], htiles[k])
        t2 = eng.tiled(w, h).upload([htiles[k] if hocc.reshape(-1)[k] else None for k in range(hocc.size)])
        exact(t2.to_flat().cpu().numpy(), exp, "upload + to_flat")
        t.close(); t2.close()


def test_device_tiled_copy_on_write(eng, oracle):
    """Snapshots of a device TiledImage share chunks until one side writes (Arc::make_mut, tiled_image.rs:330, :868)."""
    import torch

    rng = np.random.default_rng(21)
    w, h = 300, 200  # 5 x 4 chunks
    img = fx.random_rgba(rng, w, h)
    img[:64, :128] = 0  # chunks 0 and 1 unpopulated
    a = eng.tiled(w, h).from_flat(torch.from_numpy(img).cuda())
    occ_a, tiles_a = a.download()
    assert list(np.flatnonzero(occ_a == 0)) == [0, 1]
    snap = a.clone()
    assert np.array_equal(a.chunk_ids(), snap.chunk_ids())  # nothing copied: every chunk shared
    exact(snap.to_flat().cpu().numpy(), oracle.tiled_roundtrip(img)[0], "clone reads like the original")
    # write chunk 7 (populated, shared) and chunk 0 (absent) of the original through make_mut
    a.make_mut([7, 0])
    ids_a, ids_s = a.chunk_ids(), snap.chunk_ids()
    assert ids_a[7] != ids_s[7] and ids_a[0] >= 0
    same = [k for k in range(a.n_chunks) if k not in (0, 7) and occ_a[k]]
    assert all(ids_a[k] == ids_s[k] for k in same)
    occ2, tiles2 = a.download()
    assert occ2[0] == 1 and not tiles2[0].any()                 # a fresh chunk is transparent
    assert np.array_equal(tiles2[7], tiles_a[7])                 # the private copy carries the pixels over
    exact(snap.to_flat().cpu().numpy(), oracle.tiled_roundtrip(img)[0], "the snapshot is untouched")
    # replacing the original's content leaves the snapshot alone, and the flatten reads each through its own table
    img2 = fx.random_rgba(rng, w, h)
    a.from_flat(torch.from_numpy(img2).cuda())
    assert not set(a.chunk_ids()[a.chunk_ids() >= 0]) & set(snap.chunk_ids()[snap.chunk_ids() >= 0])
    exact(snap.to_flat().cpu().numpy(), oracle.tiled_roundtrip(img)[0], "snapshot after the original was overwritten")
    exact(a.to_flat().cpu().numpy(), oracle.tiled_roundtrip(img2)[0], "original after from_flat")
    got = eng.flatten_tiles([dict(tiles=snap, opacity=1.0, blend=0), dict(tiles=a, opacity=0.6, blend=8)], w, h)
    exp = oracle.flatten([oracle.make_layer(oracle.tiled_roundtrip(img)[0]), oracle.make_layer(oracle.tiled_roundtrip(img2)[0], opacity=0.6, blend=8)], w, h)
    exact(got.cpu().numpy(), exp, "flatten of a snapshot under its successor")
    # dropping the snapshot returns its chunks to the pool: a new image reuses them
    freed = set(snap.chunk_ids()[snap.chunk_ids() >= 0])
    snap.close()
    b = eng.tiled(w, h).from_flat(torch.from_numpy(img).cuda())
    assert set(b.chunk_ids()[b.chunk_ids() >= 0]) & freed
    a.close(); b.close()


def test_gaussian_ring_pipeline_full_8k(eng, monkeypatch):
    """The V pass's producer/consumer ring (one CTA walks ~56 tiles back to back at 8K) against the
    direct-from-global V kernel on EVERY pixel of an 8K image, repeatedly: a chunk refilled while a slow
    warp still reads it would show up here and nowhere at test sizes where a CTA sees one or two tiles."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(1)
    img = torch.randint(0, 256, (4320, 7680, 4), dtype=torch.uint8, device="cuda", generator=g)
    for sigma in (20.0, 5.0, 35.0, 50.0):
        monkeypatch.setenv("PFE_GAUSS_V_DIRECT", "1")
        ref = eng.gaussian_blur(img, sigma).clone()
        monkeypatch.delenv("PFE_GAUSS_V_DIRECT")
        for rep in range(6):
            out = eng.gaussian_blur(img, sigma)
            bad = int((out != ref).any(dim=-1).sum())
            assert bad == 0, f"sigma {sigma} rep {rep}: {bad} pixels differ between the ring and the direct V pass"


def test_disp_reach_matches_host_formula(eng):
    """pfe_dev_disp_reach (halo sizing of the banded displacement warp) == the torch formula it replaces."""
    import torch

    rng = np.random.default_rng(9)
    for (w, rows, y0, h_total) in [(300, 128, 256, 1000), (67, 45, 0, 45), (1, 1, 5, 9), (1030, 64, 960, 1024)]:
        d = rng.normal(0, 40, (rows, w, 2)).astype(np.float32)
        d[0, 0] = (0.0, np.nan); d[-1, -1] = (1.0, np.inf); d[rows // 2, w // 2] = (0.0, -np.inf)
        if rows > 2:
            d[1, 0, 1] = 1e9; d[2, 0, 1] = -1e9
        t = torch.from_numpy(d).cuda()
        ys = torch.arange(y0, y0 + rows, dtype=torch.float32, device="cuda")[:, None]
        sy = (ys - torch.nan_to_num(t[..., 1], nan=0.0, posinf=0.0, neginf=0.0)).clamp(-1.0, float(h_total))
        want = [int(torch.floor(sy.min())), int(torch.floor(sy.max()))]
        assert eng.disp_reach(t, y0, h_total).tolist() == want, (w, rows, y0, h_total)


def test_error_statuses_of_widened_entry_points(eng):
    """Bad enums / unsupported sizes come back as statuses (the ABI never unwinds, never falls back) and the
    context stays usable afterwards."""
    from paintfe_b200._lib import PfeError

    img = fx.gradient(64, 48)

    def code(fn, *a, **k):
        with pytest.raises(PfeError) as e:
            fn(*a, **k)
        return e.value.code

    assert code(eng.halftone, img, 4.0, 45.0, 9) == -1            # unknown HalftoneShape
    assert code(eng.color_filter, img, (1, 2, 3, 4), 0.5, 7) == -1  # unknown ColorFilterMode
    assert code(eng.grid, img, 8, 8, 1, (0, 0, 0, 255), 5, 1.0) == -1
    assert code(eng.outline, img, 2, (0, 0, 0, 255), 3, True) == -1
    assert code(eng.outline, img, 100000, (0, 0, 0, 255), 0, True) == -2  # O(width^2) search refused
    assert code(eng.bokeh_blur, img, 1.0e6) == -2
    assert code(eng.orient, img, 9) == -1
    assert code(eng.resize, img, 10, 10, 9) == -1
    assert code(eng.resize, img, 0, 10, 1) == -1
    assert code(eng.adjust, img, 17) == -1                         # between the two op families
    assert code(eng.adjust, img, 14) == -1                         # gradient map without its LUT
    assert code(eng.flatten_tiles, [dict(kind=9)], 64, 48) == -1
    exact(eng.orient(eng.orient(img, 0), 0), img, "context still usable")


def test_script_host_on_device(eng, oracle):
    """paintfe_b200/rhai_host.py against the CUDA path: control flow around effect calls, pixel access between them and
    closures, on host arrays (host tier) and on a CUDA tensor (device tier, pending set_pixel writes carried over)."""
    import torch

    from paintfe_b200.rhai_host import run_script
    from test_rhai_host import OracleEngine

    class Ref(OracleEngine):
        def box_blur(self, img, radius, mask=None):
            return self.o.box_blur(img, radius, mask=mask)

        def median(self, img, radius, mask=None):
            return self.o.median(img, radius, mask=mask)

    source = """
        fn passes(n) { if n > 2 { 2 } else { n } }
        set_pixel(3, 2, 250, 10, 20, 255);
        for i in 0..passes(5) { apply_box_blur(i + 1); }
        let p = get_pixel(3, 2); print(`${p}`);
        select_ellipse(40.0, 30.0, 25.0, 18.0);
        if has_selection() && p[3] == 255 { apply_median(1); apply_invert(); }
        for_each_pixel(|x, y, r, g, b, a| { if is_selected(x, y) { [r, b, g, a] } else { [r / 2, g, b, a] } });
        set_pixel(0, 0, 1, 2, 3, 4);
        invert_selection(); fill_selected(9, 8, 7, 255);
        rotate_canvas_90cw();
        print(`${width()}x${height()} ${get_pixel(width() - 1, 0)}`);
        apply_pixelate(3);
    """
    img = fx.random_rgba(np.random.default_rng(12), 96, 64, alpha="opaque")
    want, console = run_script(Ref(oracle), source, img)
    got, con_host = run_script(eng, source, img)
    exact(got, want, "script host, host tier")
    dev, con_dev = run_script(eng, source, torch.from_numpy(img).cuda())
    assert dev.is_cuda
    exact(dev.cpu().numpy(), want, "script host, device tier")
    assert console == con_host == con_dev and console[1].startswith("64x96 ")
    exact(run_script(eng, "for_each_pixel(|x, y, r, g, b, a| { [255 - r, 255 - g, 255 - b, a] });", fx.gradient(64, 64))[0],
          fx.golden("scripting", "for_each_pixel_invert"))


def test_cli_script_canvas_ops_replayed_on_other_layers(eng, oracle, tmp_path, capsys):
    """cli.rs:247-266: console lines under --verbose, canvas-wide calls replayed on the non-active layers."""
    from PIL import Image

    from paintfe_b200 import cli, pfe_io

    rng = np.random.default_rng(21)
    w, h = 150, 90
    bg, top = fx.random_rgba(rng, w, h, alpha="opaque"), fx.random_rgba(rng, w, h)
    proj = pfe_io.PfeProject(w, h, 1, [pfe_io.layer_from_flat("bg", bg), pfe_io.layer_from_flat("top", top, opacity=0.7, blend_mode=3)])
    (tmp_path / "p.pfe").write_bytes(pfe_io.save_pfe_v3(proj))
    script = tmp_path / "s.rhai"
    script.write_text('let w = width();\nif w > 100 { rotate_canvas_90ccw(); resize_image(60, 100, "bicubic"); }\n'
                      'apply_invert(); print_line(`now ${width()}x${height()}`);\n')
    assert cli.main(["-i", str(tmp_path / "p.pfe"), "-s", str(script), "-o", str(tmp_path / "o.png"), "-v"]) == 0
    assert "  [script] now 60x100" in capsys.readouterr().out
    tf = lambda im: oracle.resize(oracle.orient(im, oracle.ROT90CCW), 60, 100, 2)
    top2 = oracle.adjust(tf(top), 32)
    exp = oracle.flatten([oracle.make_layer(tf(bg)), oracle.make_layer(top2, opacity=0.7, blend=3)], 60, 100)
    exact(np.array(Image.open(tmp_path / "o.png").convert("RGBA")), exp, "cli script with canvas ops")
    script.write_text("let x = 1 / 0;\n")
    assert cli.main(["-i", str(tmp_path / "p.pfe"), "-s", str(script), "-o", str(tmp_path / "o2.png")]) == 1
    assert "script error" in capsys.readouterr().err


def test_script_runner_covers_effect_api(eng, oracle):
    """The Rhai bindings' fixed arguments (scripting.rs:822-1165) through the script runner."""
    from paintfe_b200.script import execute_script_sync

    img = fx.gradient(64, 64)
    exact(execute_script_sync(eng, "apply_pixelate(4);", img), fx.golden("scripting", "apply_pixelate"))
    exact(execute_script_sync(eng, "apply_glow(3.0, 0.5);", img, exact=True), fx.golden("filters", "glow_r3_i05"))
    exact(execute_script_sync(eng, "apply_bulge(0.5);", img), fx.golden("filters", "bulge_05"))
    within1(execute_script_sync(eng, "apply_twist(45.0);", img), fx.golden("filters", "twist_45"))
    within1(execute_script_sync(eng, "apply_noise(30.0, true);", img), fx.golden("filters", "add_noise_gaussian_mono"))
    within1(execute_script_sync(eng, "apply_reduce_noise(0.5);", img), fx.golden("filters", "reduce_noise"))
    exact(execute_script_sync(eng, "apply_median(2); apply_box_blur(3); apply_motion_blur(45.0, 10.0);", img),
          oracle.motion_blur(oracle.box_blur(oracle.median(img, 2), 3.0), 45.0, 10.0))
    exact(execute_script_sync(eng, "apply_oil_painting(3);", img), fx.golden("filters", "oil_painting"))
    exact(execute_script_sync(eng, "apply_ink(1.0, 0.5);", img), fx.golden("filters", "ink"))
    exact(execute_script_sync(eng, "apply_crystallize(16);", img), fx.golden("filters", "crystallize_s16"))
    exact(execute_script_sync(eng, "apply_halftone(4.0);", img), fx.golden("filters", "halftone_circle"))
    exact(execute_script_sync(eng, "flip_horizontal();", img), fx.golden("scripting", "flip_horizontal"))
    exact(execute_script_sync(eng, "flip_vertical();", img), fx.golden("scripting", "flip_vertical"))
    exact(execute_script_sync(eng, "rotate_180(); flip_canvas_horizontal(); flip_canvas_vertical();", img), img)
    g = fx.gradient(64, 48)
    exact(execute_script_sync(eng, "rotate_canvas_90cw();", g), oracle.orient(g, oracle.ROT90CW))
    exact(execute_script_sync(eng, "rotate_canvas_90ccw(); rotate_canvas_180();", g), oracle.orient(g, oracle.ROT90CW))
    exact(execute_script_sync(eng, 'resize_image(32, 24, "lanczos");', g), fx.golden("transforms", "resize_half_lanczos"))
    exact(execute_script_sync(eng, 'resize_image(128, 96, "nn"); resize_image(128, 96, "bicubic");', g), fx.golden("transforms", "resize_2x_nearest"))
    exact(execute_script_sync(eng, 'resize_canvas(96, 80, "center");', g), fx.golden("transforms", "resize_canvas_center"))
    exact(execute_script_sync(eng, 'resize_canvas(70, 50, "se");', g), oracle.resize_canvas(g, 70, 50, (2, 2), (0, 0, 0, 0)))
    with pytest.raises(ValueError):
        execute_script_sync(eng, "for_each_pixel(3);", img)


def test_script_selection_api(eng, oracle):
    """tests/scripting.rs:266-400: select_rect / select_ellipse / invert / clear / fill / delete drive the
    mask the effects honour."""
    import torch
    from paintfe_b200.script import execute_script_sync

    img = fx.gradient(64, 64)
    rect = np.zeros((64, 64), np.uint8); rect[10:40, 5:30] = 255
    exact(execute_script_sync(eng, "select_rect(5, 10, 30, 40); apply_blur(3.0);", img, exact=True), oracle.gaussian_blur(img, 3.0, mask=rect))
    exact(execute_script_sync(eng, "select_rect(5, 10, 30, 40); invert_selection(); apply_box_blur(2);", img),
          oracle.box_blur(img, 2.0, mask=255 - rect))
    exact(execute_script_sync(eng, "select_rect(-5, -5, 500, 500); clear_selection(); apply_median(1);", img), oracle.median(img, 1))
    yy, xx = np.mgrid[0:64, 0:64].astype(np.float64)
    ell = np.where(((xx - 32.0) ** 2) / 100.0 + ((yy - 20.0) ** 2) / 400.0 <= 1.0, 255, 0).astype(np.uint8)
    exact(execute_script_sync(eng, "select_ellipse(32.0, 20.0, 10.0, 20.0); apply_vignette(0.8, 0.5);", img), oracle.vignette(img, 0.8, 0.5, mask=ell))
    want = img.copy(); want[rect > 0] = (9, 8, 7, 200)
    exact(execute_script_sync(eng, "select_rect(5, 10, 30, 40); fill_selected(9, 8, 7, 200);", img), want)
    want = img.copy(); want[rect > 0] = 0
    dev = execute_script_sync(eng, "select_rect(5, 10, 30, 40); delete_selected();", torch.from_numpy(img).cuda())
    exact(dev.cpu().numpy(), want)
    exact(execute_script_sync(eng, "invert_selection(); fill_selected(1, 2, 3, 4);", img), img)  # nothing selected
    exact(execute_script_sync(eng, "delete_selected();", img), np.zeros_like(img))                 # no mask = everything


def test_host_tier_band_pipeline(eng, oracle):
    """pfe_flatten / pfe_flatten_gaussian pipeline uploads, compute and downloads in row bands once the
    image exceeds 8 MB; the result must equal the unpipelined device tier, including a blur radius larger
    than a band and masks / adjustment layers / the active-chunk bitmap crossing band boundaries."""
    import torch
    from paintfe_b200.engine import make_layer

    rng = np.random.default_rng(31)
    w, h = 2048, 1100  # 9 MB per layer -> 6 bands of 192 rows, ragged last band
    imgs = [fx.random_rgba(rng, w, h) for _ in range(4)]
    imgs[1][:640, :1024] = 0
    mask = rng.integers(0, 256, (h, w), dtype=np.uint8)
    specs = [dict(rgba=imgs[0], blend=0, opacity=1.0), dict(rgba=imgs[1], blend=8, opacity=0.6, mask=mask),
             dict(kind=3, opacity=0.5), dict(rgba=imgs[2], blend=21, opacity=0.9), dict(rgba=imgs[3], blend=13, opacity=0.4)]
    host_layers = [make_layer(**s) for s in specs]
    dev_layers = [make_layer(**{k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in s.items()}) for s in specs]
    active = np.ones(((h + 63) // 64, (w + 63) // 64), np.uint8)
    active[3:9, 5:20] = 0
    flat_dev = eng.flatten(dev_layers, w, h, active=torch.from_numpy(active).cuda())
    exact(eng.flatten(host_layers, w, h, active=active), flat_dev.cpu().numpy(), "pipelined flatten")
    exact(flat_dev[:320, :256].cpu().numpy(),
          oracle.flatten([oracle.make_layer(**{k: (v[:320, :256] if isinstance(v, np.ndarray) else v) for k, v in s.items()}) for s in specs],
                         256, 320, active=active[:5, :4]), "flatten crop vs oracle")
    for sigma in (20.0, 80.0):  # radius 60 < band height 192 < radius 240
        for ex in (True, False):
            want = eng.gaussian_blur(flat_dev, sigma, exact=ex).cpu().numpy()
            exact(eng.flatten_gaussian(host_layers, w, h, sigma, active=active, exact=ex), want, f"pipelined flatten+gaussian s={sigma} exact={ex}")


def test_warp_band_window_check_is_asynchronous(eng):
    """pfe_dev_warp_band no longer synchronises: a source window that does not cover the warp's reach is recorded
    in the context's sticky flag and reported (once) by pfe_ctx_check_async."""
    import torch
    from paintfe_b200._lib import PfeError

    rng = np.random.default_rng(5)
    w, h = 64, 96
    src = torch.from_numpy(rng.integers(0, 256, (h, w, 4), dtype=np.uint8)).cuda()
    disp = torch.zeros((32, w, 2), dtype=torch.float32, device="cuda")
    disp[..., 1] = 20.0  # rows 32..63 sample rows 12..43
    ok = eng.warp_band(src[8:70].contiguous(), h, 8, w, h, 32, 32, disp_band=disp)
    eng.check_async()
    whole = torch.zeros((h, w, 2), dtype=torch.float32, device="cuda")
    whole[32:64] = disp
    assert torch.equal(ok, eng.warp_displacement(src, whole)[32:64])
    eng.warp_band(src[30:70].contiguous(), h, 30, w, h, 32, 32, disp_band=disp)  # rows 12..29 are missing
    with pytest.raises(PfeError):
        eng.check_async()
    eng.check_async()


def test_warp_displacement_region(eng, oracle):
    """warp_displacement_region (transform.rs:1206-1285): prev outside the dirty rect, the warp inside; the reference's
    rect clamping (negative x1 = to the edge), an inverted rect, in place on prev; host tier == device tier."""
    import torch

    rng = np.random.default_rng(12)
    w, h = 150, 97
    src, prev = fx.random_rgba(rng, w, h), fx.random_rgba(rng, w, h)
    disp = rng.normal(0, 6, (h, w, 2)).astype(np.float32)
    for rect in ((10, 20, 90, 70), (-5, -5, 400, 400), (30, 10, -1, 50), (80, 40, 20, 60), (0, 0, 0, 0), (149, 96, 150, 97)):
        exp = oracle.warp_displacement_region(src, disp, prev, rect)
        exact(eng.warp_displacement_region(src, disp, prev, rect), exp, f"region {rect} host tier")
        d = eng.warp_displacement_region(torch.from_numpy(src).cuda(), torch.from_numpy(disp).cuda(), torch.from_numpy(prev).cuda(), rect)
        exact(d.cpu().numpy(), exp, f"region {rect} device tier")
    p = torch.from_numpy(prev).cuda()
    eng.warp_displacement_region(torch.from_numpy(src).cuda(), torch.from_numpy(disp).cuda(), p, (10, 20, 90, 70), out=p)
    exact(p.cpu().numpy(), oracle.warp_displacement_region(src, disp, prev, (10, 20, 90, 70)), "region in place")
    # inside a full-canvas rect it is the full warp
    exact(eng.warp_displacement_region(src, disp, prev, (0, 0, w, h)), oracle.warp_displacement(src, disp), "full rect == full warp")


def test_flatten_peer_stores_twice_and_flags(eng):
    """pfe_dev_flatten_peer with a local "peer": both destinations hold the flatten, the flag carries the value, a
    wait on it passes and a wait on a value nobody will write times out into the sticky error."""
    import torch
    from paintfe_b200.engine import make_layer

    rng = np.random.default_rng(31)
    for (w, h, n) in ((256, 40, 6), (13, 9, 3), (64, 16, 40)):  # one launch; scalar tail (odd width); chained launches
        imgs = [torch.from_numpy(rng.integers(0, 256, (h, w, 4), dtype=np.uint8)).cuda() for _ in range(n)]
        layers = [make_layer(t, blend=(5 * i) % 25, opacity=0.2 + 0.02 * i) for i, t in enumerate(imgs)]
        want = eng.flatten(layers, w, h)
        prep = eng.prepare_layers(layers, w, h)
        out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        far = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        flags = torch.zeros(4, dtype=torch.int32, device="cuda")
        eng.flatten_prepared_peer(prep, out, far.data_ptr(), flags.data_ptr() + 4, 7)
        eng.peer_wait(flags.data_ptr() + 4, 1, 7, timeout_ms=1000)
        assert torch.equal(out, want) and torch.equal(far, want)
        assert flags.tolist() == [0, 7, 0, 0]
        eng.check_async()
    # without a flag: stores only
    far.zero_()
    eng.flatten_prepared_peer(prep, out, far.data_ptr())
    assert torch.equal(far, want)
    # counters wrap: 0xFFFFFFFE has "reached" 0xFFFFFFFD but not 2
    flags[0] = -2
    eng.peer_wait(flags.data_ptr(), 1, 0xFFFFFFFD, timeout_ms=1000)
    eng.check_async()
    eng.peer_wait(flags.data_ptr(), 1, 2, timeout_ms=20)
    with pytest.raises(Exception, match="did not arrive"):
        eng.check_async()
    eng.check_async()


def test_flatten_plain_instantiation_matches_general(eng, oracle):
    """A stack of raster layers without masks takes the PLAIN instantiation of the flatten kernel (kind and mask tests
    compiled out); PFE_FLATTEN_NO_PLAIN forces the general one. Both must equal the oracle."""
    import os

    rng = np.random.default_rng(77)
    w, h = 332, 75
    imgs = [fx.random_rgba(rng, w, h) for _ in range(25)]
    imgs[3][:, : w // 2] = 0  # a half-transparent layer: the warp-level skip fires
    layers = [dict(rgba=im, blend=i, opacity=[1.0, 0.6, 0.25][i % 3]) for i, im in enumerate(imgs)]
    exp = oracle.flatten([oracle.make_layer(**L) for L in layers], w, h)
    from paintfe_b200.engine import make_layer
    got = eng.flatten([make_layer(**L) for L in layers], w, h)
    exact(got, exp, "flatten PLAIN")
    os.environ["PFE_FLATTEN_NO_PLAIN"] = "1"
    try:
        exact(eng.flatten([make_layer(**L) for L in layers], w, h), exp, "flatten general")
    finally:
        del os.environ["PFE_FLATTEN_NO_PLAIN"]
