"""Host-side multi-GPU logic on CPU: world_size-2/3 gloo process groups, the oracle standing in for
the engine (tests may use the oracle; the product never does).  Property checked everywhere:
band-split result, concatenated over ranks, == the whole-image result, bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fixtures as fx


class OracleEngine:
    """Engine-shaped stand-in (same method names/arguments as paintfe_b200.engine.Engine)."""

    def __init__(self):
        from oracle import pfo
        self.o = pfo

    def gaussian_blur(self, img, sigma, mask=None, exact=False, out=None): return self.o.gaussian_blur(img, sigma, mask=mask)
    def box_blur(self, img, radius, mask=None, out=None): return self.o.box_blur(img, radius, mask=mask)
    def median(self, img, radius, mask=None, out=None): return self.o.median(img, radius, mask=mask)
    def sharpen(self, img, amount, radius, mask=None, exact=False, out=None): return self.o.sharpen(img, amount, radius, mask=mask)
    def ink(self, img, es, th, mask=None, out=None): return self.o.ink(img, es, th, mask=mask)
    def oil_painting(self, img, r, lv, mask=None, out=None): return self.o.oil_painting(img, r, lv, mask=mask)
    def bokeh_blur(self, img, r, mask=None, out=None): return self.o.bokeh_blur(img, r, mask=mask)
    def reduce_noise(self, img, s, r, mask=None, out=None): return self.o.reduce_noise(img, s, r, mask=mask)
    def motion_blur(self, img, a, d, mask=None, out=None): return self.o.motion_blur(img, a, d, mask=mask)
    def adjust(self, img, op, params=(), luts=None, mask=None, occupancy=None, out=None):
        return self.o.adjust(np.asarray(img), op, params, luts=luts, mask=None if mask is None else np.asarray(mask),
                             occupancy=None if occupancy is None else np.asarray(occupancy))

    def flatten(self, layers, w, h, active=None, out=None):
        return self.o.flatten([self.o.make_layer(**{k: (np.asarray(v) if k in ("rgba", "mask") and v is not None else v)
                                                    for k, v in L.items()}) for L in layers], w, h, active=active)

    def prepare_layers(self, layers, w, h):
        return (list(layers), int(w), int(h))

    def flatten_prepared(self, prepared, out, active=None):
        layers, w, h = prepared
        out.copy_(torch.from_numpy(self.flatten(layers, w, h, active=active)))
        return out

    def warp_band(self, src_rows, src_h, src_y0, w, h, y0, rows_out, disp_band=None, original=None, deformed=None,
                  cols=0, rows=0, out=None):
        src_rows = np.asarray(src_rows)
        sw = src_rows.shape[1]
        full = np.zeros((src_h, sw, 4), np.uint8)
        full[src_y0:src_y0 + src_rows.shape[0]] = src_rows
        if disp_band is not None:
            d = np.zeros((h, w, 2), np.float32)
            d[y0:y0 + rows_out] = np.asarray(disp_band)
            res = self.o.warp_displacement(full, d)
        else:
            res = self.o.mesh_warp(full, original, deformed, cols, rows, w, h)
        # emulate the library's window check: a tap outside the provided rows must not be needed
        probe = np.full((src_h, sw, 4), 255, np.uint8)
        probe[src_y0:src_y0 + src_rows.shape[0]] = src_rows
        if disp_band is not None:
            res2 = self.o.warp_displacement(probe, d)
        else:
            res2 = self.o.mesh_warp(probe, original, deformed, cols, rows, w, h)
        assert np.array_equal(res[y0:y0 + rows_out], res2[y0:y0 + rows_out]), "halo too small for the warp's reach"
        return res[y0:y0 + rows_out]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from paintfe_b200 import dist as pd

        eng = OracleEngine()
        rng = np.random.default_rng(1234)
        w, h = 96, 200 if case != "thin" else 130
        img = fx.random_rgba(rng, w, h)
        bounds = pd.band_bounds(h, world) if case != "thin" else pd.band_bounds(h, world, align=16)
        y0, y1 = bounds[rank]
        band = torch.from_numpy(img[y0:y1].copy())
        if case == "gaussian":
            out = pd.gaussian_blur_banded(eng, band, h, 7.0, exact=True, bounds=bounds)
            exp = eng.gaussian_blur(img, 7.0)
        elif case == "thin":  # halo (30 rows) wider than a band (16-48 rows): multi-neighbour exchange
            out = pd.gaussian_blur_banded(eng, band, h, 10.0, exact=True, bounds=bounds)
            exp = eng.gaussian_blur(img, 10.0)
        elif case == "box":
            out = pd.box_blur_banded(eng, band, h, 5.0, bounds=bounds)
            exp = eng.box_blur(img, 5.0)
        elif case == "median":
            out = pd.median_banded(eng, band, h, 3, bounds=bounds)
            exp = eng.median(img, 3)
        elif case == "sharpen":
            out = pd.sharpen_banded(eng, band, h, 1.5, 2.0, exact=True, bounds=bounds)
            exp = eng.sharpen(img, 1.5, 2.0)
        elif case == "effects":  # translation-invariant neighbourhood effects through the generic halo path
            out = pd.ink_banded(eng, band, h, 2.0, 0.5, bounds=bounds)
            exp = eng.ink(img, 2.0, 0.5)
            for got, want in ((pd.oil_painting_banded(eng, band, h, 4, 24, bounds=bounds), eng.oil_painting(img, 4, 24)),
                              (pd.bokeh_blur_banded(eng, band, h, 6.5, bounds=bounds), eng.bokeh_blur(img, 6.5)),
                              (pd.reduce_noise_banded(eng, band, h, 12.0, 3, bounds=bounds), eng.reduce_noise(img, 12.0, 3)),
                              (pd.motion_blur_banded(eng, band, h, 70.0, 9.0, bounds=bounds), eng.motion_blur(img, 70.0, 9.0))):
                assert np.array_equal(np.asarray(got), want[y0:y1])
        elif case == "flatten":
            imgs = [fx.random_rgba(rng, w, h) for _ in range(5)]
            meta = [dict(blend=(3 * i) % 25, opacity=0.3 + 0.15 * i) for i in range(5)]
            layers = [dict(rgba=torch.from_numpy(im[y0:y1].copy()), **m) for im, m in zip(imgs, meta)]
            out = torch.from_numpy(pd.flatten_banded(eng, layers, w, y1 - y0))
            exp = eng.o.flatten([eng.o.make_layer(im, **m) for im, m in zip(imgs, meta)], w, h)
        elif case in ("flatten_blur", "flatten_blur_small"):  # BandedFlattenBlur on the CPU stand-in: NCCL-shaped plan over gloo, sequential step
            sigma = 7.0 if case == "flatten_blur" else 2.0
            imgs = [fx.random_rgba(rng, w, h) for _ in range(4)]
            meta = [dict(blend=(5 * i) % 25, opacity=0.4 + 0.15 * i) for i in range(4)]
            layers = [dict(rgba=torch.from_numpy(im[y0:y1].copy()), **m) for im, m in zip(imgs, meta)]
            fb = pd.BandedFlattenBlur(eng, layers, w, h, sigma, exact=True, bounds=bounds)
            assert fb.transport == "nccl" and fb.peer is None  # peer memory needs CUDA: "auto" settles on the collective plan
            assert sum(b - a for a, b, _ in fb.parts) == y1 - y0
            for _ in range(2):  # buffers are reused step after step
                out = fb.step().clone()
            exp = eng.gaussian_blur(eng.o.flatten([eng.o.make_layer(im, **m) for im, m in zip(imgs, meta)], w, h), sigma)
            with pytest.raises(pd.PeerUnavailable):
                pd.BandedFlattenBlur(eng, layers, w, h, sigma, exact=True, bounds=bounds, transport="peer")
        elif case == "adjust":  # per-pixel adjustment with an occupancy bitmap whose populated chunks cross the band edges
            cyn, cxn = (h + 63) // 64, (w + 63) // 64
            occ = (rng.random((cyn, cxn)) < 0.6).astype(np.uint8)
            for r in range(1, world):  # chunk rows either side of every band edge: one populated, one not, per column
                e = bounds[r][0] // 64
                if 0 < e < cyn:
                    occ[e - 1], occ[e] = np.arange(cxn) % 2, (np.arange(cxn) + 1) % 2
            mask = (rng.random((h, w)) < 0.8).astype(np.uint8) * 255
            out = pd.adjust_banded(eng, band, h, 5, (30.0, -20.0, 10.0), mask_band=torch.from_numpy(mask[y0:y1].copy()),
                                   occupancy=occ, bounds=bounds)
            exp = eng.o.adjust(img, 5, (30.0, -20.0, 10.0), mask=mask, occupancy=occ)
        elif case == "warp":
            disp = rng.normal(0, 9, (h, w, 2)).astype(np.float32)
            out = pd.warp_displacement_banded(eng, band, torch.from_numpy(disp[y0:y1].copy()), h, bounds=bounds)
            exp = eng.o.warp_displacement(img, disp)
        elif case == "mesh":
            orig = fx.uniform_grid(6, 6, float(w), float(h))
            deformed = orig.copy()
            for i in range(7):
                for j in range(7):
                    deformed[i * 7 + j] += 8.0 * np.sin(i) * np.cos(j)  # SURVEY §8d config 4 mesh
            out = pd.mesh_warp_banded(eng, band, orig, deformed, 6, 6, w, h, bounds=bounds)
            exp = eng.o.mesh_warp(img, orig, deformed, 6, 6, w, h)
        else:
            raise ValueError(case)
        ok = np.array_equal(np.asarray(out), exp[y0:y1])
        q.put((rank, bool(ok), int(np.asarray(out).shape[0]), y1 - y0))
    finally:
        dist.destroy_process_group()


def _run(case, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, rows, want in res:
        assert rows == want, (case, rank, rows, want)
        assert ok, f"{case}: rank {rank} band differs from the whole-image result"


@pytest.mark.parametrize("case", ["gaussian", "box", "median", "sharpen", "effects", "flatten", "flatten_blur", "flatten_blur_small", "adjust", "warp", "mesh"])
def test_band_split_equals_whole_world2(case):
    _run(case, 2)


def test_band_split_halo_spans_several_ranks_world3():
    _run("thin", 3)


def test_banded_adjust_world3():
    _run("adjust", 3)


def test_band_bounds_and_sharding():
    from paintfe_b200 import dist as pd

    for h, world in [(4320, 1), (4320, 2), (4320, 4), (4320, 8), (16384, 4), (100, 8), (64, 3), (1, 2)]:
        b = pd.band_bounds(h, world)
        assert b[0][0] == 0 and b[-1][1] == h and len(b) == world
        for (a0, a1), (b0, b1) in zip(b, b[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(y0 % 64 == 0 for y0, y1 in b if y1 > y0)  # non-empty bands start on a chunk row
        units = [-(-(y1 - y0) // 64) for y0, y1 in b]
        assert max(units) - min(units) <= 1
    idx = [pd.shard_indices(1024, r, 8) for r in range(8)]
    assert sorted(sum(idx, [])) == list(range(1024)) and all(len(i) == 128 for i in idx)
    assert pd.shard_indices(3, 5, 8) == []
    assert pd.gaussian_radius(20.0) == 60 and pd.gaussian_radius(0.0) == 0 and pd.gaussian_radius(0.1) == 1


def test_script_parser_and_cli_helpers(tmp_path):
    from paintfe_b200 import cli, script

    calls = script.parse("apply_blur(4.0); // comment\n apply_hsl(10.0, 15.0, 0.0);\napply_vignette(0.5,0.3);apply_invert();")
    assert calls == [("apply_blur", (4.0,)), ("apply_hsl", (10.0, 15.0, 0.0)), ("apply_vignette", (0.5, 0.3)), ("apply_invert", ())]
    with pytest.raises(ValueError):
        script.parse("for x in 0..10 { apply_blur(1.0); }")
    for n in ("b.png", "a.png", "c.jpg"):
        (tmp_path / n).write_bytes(b"x")
    got = cli.resolve_inputs([str(tmp_path / "*.png"), str(tmp_path / "a.png"), str(tmp_path / "nope*.png")])
    assert [os.path.basename(p) for p in got] == ["a.png", "b.png"]
    assert cli.build_output_path("x/y/shot.jpg", None, "out", "png") == os.path.join("out", "shot.png")
    assert cli.build_output_path("shot.jpg", "r.png", None, "png") == "r.png"
    assert cli.build_output_path("shot.jpg", None, None, "png") is None


def test_peer_targets_address_arithmetic():
    """Where a rank's edge rows land in its neighbours' blocks (paintfe_b200.dist.peer_targets): three ranks, 60-row halos."""
    from paintfe_b200 import dist as pd

    row_bytes, header = 520 * 4, 4096
    infos = [(b"a", 0, 236, 60, 4096 * 155), (b"b", 60, 232, 60, 4096 * 180), (b"c", 60, 232, 0, 4096 * 150)]
    up, down = 1 << 30, 2 << 30
    # rank 1, in the middle: both neighbours
    rows_up, put_up, rows_down, put_down = pd.peer_targets(infos, 1, up, down, row_bytes, header)
    assert (rows_up, rows_down) == (60, 60)
    for p in (0, 1):
        # my first 60 rows -> rank 0's BOTTOM halo (after its 0 top rows and 236 band rows), flagged as "from below" (side 1)
        assert put_up[p] == (up + header + p * infos[0][4] + (0 + 236) * row_bytes, up + 4 * (2 * p + 1))
        # my last 60 rows -> rank 2's TOP halo (row 0 of its extended band), flagged as "from above" (side 0)
        assert put_down[p] == (down + header + p * infos[2][4], down + 4 * (2 * p))
        assert put_up[p][0] % 16 == 0 and put_down[p][0] % 16 == 0  # the flatten kernel's 16-byte stores
    # the image edges have one neighbour
    assert pd.peer_targets(infos, 0, None, down, row_bytes, header)[:2] == (0, None)
    r = pd.peer_targets(infos, 2, up, None, row_bytes, header)
    assert r[2:] == (0, None) and r[0] == 60 and r[1][0][0] == up + header + (60 + 232) * row_bytes
