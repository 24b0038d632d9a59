"""Row-band split across real GPUs (NCCL halo exchange) == single-GPU whole-image result.
Skipped on boxes with fewer than 2 GPUs; the same logic is covered on CPU/gloo in test_dist_cpu.py."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from paintfe_b200 import dist as pd
        from paintfe_b200.engine import Engine

        eng = Engine(rank)
        rng = np.random.default_rng(77)
        w, h = 1000, 700
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        disp = rng.normal(0, 12, (h, w, 2)).astype(np.float32)
        full = torch.from_numpy(img).cuda()
        bounds = pd.band_bounds(h, world)
        y0, y1 = bounds[rank]
        band = full[y0:y1].contiguous()
        res = {}
        for exact in (True, False):
            got = pd.gaussian_blur_banded(eng, band, h, 20.0, exact=exact, bounds=bounds)
            res[f"gauss_exact={exact}"] = torch.equal(got, eng.gaussian_blur(full, 20.0, exact=exact)[y0:y1])
        # the producer writes straight into the plan's core rows (what bench.py's strong leg does): nothing is copied
        plan = pd.halo_plan(band, pd.gaussian_radius(20.0), pd.gaussian_radius(20.0), bounds)
        plan.core.copy_(band)
        out = torch.empty_like(band)
        for _ in range(3):  # the plan's buffers are reused call after call
            pd.gaussian_blur_banded(eng, plan.core, h, 20.0, bounds=bounds, out=out)
        res["gauss_plan_core"] = torch.equal(out, eng.gaussian_blur(full, 20.0)[y0:y1])
        # flatten + blur of one canvas with the edge-rows-first schedule (bench.py's strong leg)
        from paintfe_b200.engine import make_layer
        lrng = np.random.default_rng(5)
        limgs = [torch.from_numpy(lrng.integers(0, 256, (h, w, 4), dtype=np.uint8)).cuda() for _ in range(5)]
        lmeta = [dict(blend=(7 * i) % 25, opacity=0.3 + 0.15 * i) for i in range(5)]
        want = eng.gaussian_blur(eng.flatten([make_layer(t, **m) for t, m in zip(limgs, lmeta)], w, h), 20.0)[y0:y1]
        fb = pd.BandedFlattenBlur(eng, [make_layer(t[y0:y1], **m) for t, m in zip(limgs, lmeta)], w, h, 20.0, bounds=bounds)
        for _ in range(3):
            got = fb.step()
        res["flatten_blur_edge_first"] = torch.equal(got, want) and len(fb.parts) >= 2  # edge rows are their own launches
        res["flatten_blur_default_transport_is_peer"] = fb.transport == "peer"
        fbn = pd.BandedFlattenBlur(eng, [make_layer(t[y0:y1], **m) for t, m in zip(limgs, lmeta)], w, h, 20.0, bounds=bounds, transport="nccl")
        for _ in range(3):
            got = fbn.step()
        res["flatten_blur_nccl_transport"] = torch.equal(got, want) and fbn.transport == "nccl"
        for t in limgs:  # new pixels every step: a stale halo row would show
            t.add_(3)
        want = eng.gaussian_blur(eng.flatten([make_layer(t, **m) for t, m in zip(limgs, lmeta)], w, h), 20.0)[y0:y1]
        res["flatten_blur_peer_new_pixels"] = torch.equal(fb.step(), want)
        eng.check_async()  # no wait timed out
        fb.close()
        fb2 = pd.BandedFlattenBlur(eng, [make_layer(t[y0:y1], **m) for t, m in zip(limgs, lmeta)], w, h, 4.0, exact=True, bounds=bounds)
        want2 = eng.gaussian_blur(eng.flatten([make_layer(t, **m) for t, m in zip(limgs, lmeta)], w, h), 4.0, exact=True)[y0:y1]
        res["flatten_blur_small_radius"] = torch.equal(fb2.step(), want2)
        res["gauss_small_radius"] = torch.equal(pd.gaussian_blur_banded(eng, band, h, 3.0, exact=True, bounds=bounds),
                                                eng.gaussian_blur(full, 3.0, exact=True)[y0:y1])
        res["box"] = torch.equal(pd.box_blur_banded(eng, band, h, 9.0, bounds=bounds), eng.box_blur(full, 9.0)[y0:y1])
        res["median"] = torch.equal(pd.median_banded(eng, band, h, 2, bounds=bounds), eng.median(full, 2)[y0:y1])
        res["sharpen"] = torch.equal(pd.sharpen_banded(eng, band, h, 1.0, 2.0, bounds=bounds), eng.sharpen(full, 1.0, 2.0)[y0:y1])
        dfull = torch.from_numpy(disp).cuda()
        fband = dfull[y0:y1].contiguous()
        reach = pd.displacement_reach(eng, fband, h, bounds=bounds)  # once per field
        res["warp"] = torch.equal(pd.warp_displacement_banded(eng, band, fband, h, bounds=bounds, reach=reach),
                                  eng.warp_displacement(full, dfull)[y0:y1])
        res["warp_reach_computed_inside"] = torch.equal(pd.warp_displacement_banded(eng, band, fband, h, bounds=bounds),
                                                        eng.warp_displacement(full, dfull)[y0:y1])
        eng.check_async()  # no tap fell outside the exchanged rows
        # a halo that is too small is reported asynchronously
        pd.warp_displacement_banded(eng, band, fband, h, bounds=bounds, reach=(0, 0))
        try:
            eng.check_async()
            res["warp_short_halo_detected"] = False
        except Exception:
            res["warp_short_halo_detected"] = True
        eng.check_async()  # the flag was cleared
        # per-pixel adjustment with an occupancy bitmap whose populated chunks differ either side of the band edge (row e3)
        cyn, cxn = (h + 63) // 64, (w + 63) // 64
        occ = (rng.random((cyn, cxn)) < 0.6).astype(np.uint8)
        e = bounds[1][0] // 64
        occ[e - 1], occ[e] = np.arange(cxn) % 2, (np.arange(cxn) + 1) % 2
        got = pd.adjust_banded(eng, band, h, 5, (30.0, -20.0, 10.0), occupancy=torch.from_numpy(occ).cuda(), bounds=bounds)
        res["adjust_occupancy"] = torch.equal(got, eng.adjust(full, 5, (30.0, -20.0, 10.0), occupancy=occ)[y0:y1])
        orig = np.array([[c / 6 * w, r / 6 * h] for r in range(7) for c in range(7)], np.float32)
        deformed = orig + np.array([[8 * np.sin(i) * np.cos(j)] * 2 for i in range(7) for j in range(7)], np.float32)
        res["mesh"] = torch.equal(pd.mesh_warp_banded(eng, band, orig, deformed, 6, 6, w, h, bounds=bounds),
                                  eng.mesh_warp(full, orig, deformed, 6, 6, w, h)[y0:y1])
        q.put((rank, res))
        eng.close()
    finally:
        dist.destroy_process_group()


def _peer_worker(rank, world, port, q):
    """Ranks of ONE node that share GPU 0: the rendezvous is gloo, the halo rows travel through CUDA-IPC mapped memory
    exactly as they do between two GPUs (the mapping is then local instead of NVLink)."""
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0")
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from paintfe_b200 import dist as pd
        from paintfe_b200.engine import Engine, make_layer

        eng = Engine(0)
        w, h = 520, 700
        bounds = pd.band_bounds(h, world, align=4)  # dense layers: bands need not sit on chunk rows (236 / 232 / 232)
        y0, y1 = bounds[rank]
        lrng = np.random.default_rng(5)
        limgs = [torch.from_numpy(lrng.integers(0, 256, (h, w, 4), dtype=np.uint8)).cuda() for _ in range(5)]
        lmeta = [dict(blend=(7 * i) % 25, opacity=0.3 + 0.15 * i) for i in range(5)]
        res = {}
        fb = pd.BandedFlattenBlur(eng, [make_layer(t[y0:y1], **m) for t, m in zip(limgs, lmeta)], w, h, 20.0, bounds=bounds,
                                  transport="peer", timeout_ms=20000)
        res["transport"] = fb.transport == "peer"
        ok = True
        for step in range(5):  # new pixels every step, both buffers used more than once
            for t in limgs:
                t.add_(1 + step)
            want = eng.gaussian_blur(eng.flatten([make_layer(t, **m) for t, m in zip(limgs, lmeta)], w, h), 20.0)[y0:y1]
            ok = ok and torch.equal(fb.step(), want)
        res["five_steps_bit_equal"] = ok
        eng.check_async()
        res["flags"] = [int(v) for v in fb.peer.flags.cpu()]
        dist.barrier()
        fc = pd.BandedFlattenBlur(eng, [make_layer(t[y0:y1], **m) for t, m in zip(limgs, lmeta)], w, h, 20.0, bounds=bounds,
                                  transport="peer", timeout_ms=20000, peer_put="store")  # the flatten kernel's own stores
        okc = True
        for step in range(3):
            for t in limgs:
                t.add_(7)
            want = eng.gaussian_blur(eng.flatten([make_layer(t, **m) for t, m in zip(limgs, lmeta)], w, h), 20.0)[y0:y1]
            okc = okc and torch.equal(fc.step(), want)
        res["store_put_bit_equal"] = okc and fb.peer_put == "copy"
        eng.check_async()
        fc.close()
        # one rank cannot map its neighbour: every rank learns of it, releases what it had and reports PeerUnavailable
        real_open = eng.peer_open
        if rank == 1:
            def broken(handle):
                raise RuntimeError("boom: no mapping on this rank")
            eng.peer_open = broken
        try:
            pd.BandedFlattenBlur(eng, [make_layer(t[y0:y1], **m) for t, m in zip(limgs, lmeta)], w, h, 20.0, bounds=bounds, transport="peer")
            res["setup_failure_is_collective"] = False
        except pd.PeerUnavailable as e:
            res["setup_failure_is_collective"] = "boom" in str(e)
        eng.peer_open = real_open
        # a neighbour that never produces its rows: the wait gives up and says so instead of hanging the GPU
        if rank == 0:
            eng.peer_wait(fb.peer.wait_args(0)[0], 1, 1000, timeout_ms=50)
            try:
                eng.check_async()
                res["timeout_reported"] = False
            except Exception as e:
                res["timeout_reported"] = "did not arrive" in str(e)
            eng.check_async()
        fb.close()
        q.put((rank, res))
        eng.close()
    finally:
        dist.destroy_process_group()


def test_peer_memory_halo_three_ranks_on_one_gpu():
    import torch.multiprocessing as mp

    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        flags = res.pop("flags")
        assert all(res.values()), (rank, res)
        # side 0 = from the upper neighbour, side 1 = from the lower; steps 4 and 5 were the last into buffers 0 and 1
        want = [4 if rank > 0 else 0, 4 if rank < world - 1 else 0, 5 if rank > 0 else 0, 5 if rank < world - 1 else 0]
        assert flags == want, (rank, flags)


def test_row_bands_over_nccl_match_single_gpu():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        assert all(res.values()), (rank, res)
