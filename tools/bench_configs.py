#!/usr/bin/env python
"""Secondary benchmarks for BASELINE.json configs 2-5 (bench.py carries the headline metric).

  python tools/bench_configs.py [--config 2,3,4,5] [--steps K]
  torchrun --nproc-per-node N tools/bench_configs.py --config 4,5      # row bands / batch shards

Prints one JSON line per measurement (rank 0).  Device-resident inputs, CUDA-event timing on the
engine's stream, max over ranks.  Numbers taken here are reported in BASELINE.md §5.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from paintfe_b200 import dist as pd
from paintfe_b200.engine import Engine, make_layer
from paintfe_b200.script import execute_script_sync

W8K, H8K = 7680, 4320


def timed(fn, steps, warmup=3, world=1):
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="2,3,5")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch-per-gpu", type=int, default=128)
    ap.add_argument("--canvas", type=int, default=16384)
    args = ap.parse_args()
    configs = {int(c) for c in args.config.split(",")}
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout to the JSON lines
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine(local)
    eng.use_torch_stream()
    peak = 6557.8
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    gen = torch.Generator(device=dev).manual_seed(0x5EED + rank)
    px8k = W8K * H8K

    if 2 in configs:  # 16-layer 8K flatten, three stacks
        layers = [torch.randint(0, 256, (H8K, W8K, 4), dtype=torch.uint8, device=dev, generator=gen) for _ in range(16)]
        out = torch.empty((H8K, W8K, 4), dtype=torch.uint8, device=dev)
        for name, off, binary in (("modes 0-15, uniform alpha", 0, False), ("modes 16-24,0-6, uniform alpha", 16, False),
                                  ("modes 0-15, alpha in {0,255} (fast-path regime)", 0, True)):
            ls = layers
            if binary:
                ls = [t.clone() for t in layers]
                for t in ls:
                    t[..., 3] = torch.where(t[..., 3] > 127, 255, 0).to(torch.uint8)
            dl = [make_layer(t, blend=(i + off) % 25, opacity=0.25 + 0.05 * i) for i, t in enumerate(ls)]
            ms = timed(lambda: eng.flatten(dl, W8K, H8K, out=out), args.steps)
            gbs = 68 * px8k / (ms * 1e-3) / 1e9
            emit({"config": 2, "workload": f"8K 16-layer flatten, {name}", "ms": ms, "mpx_s": px8k / ms / 1e3,
                  "achieved_gbs": gbs, "frac_of_hbm": gbs / peak, "n_gpus": 1})
            del ls
        del layers, dl

    if 3 in configs:  # Gaussian sigma=50 + HSL + unsharp on 8K
        img = torch.randint(0, 256, (H8K, W8K, 4), dtype=torch.uint8, device=dev, generator=gen)
        a, b = torch.empty_like(img), torch.empty_like(img)

        def pipe():
            eng.gaussian_blur(img, 50.0, out=a)
            eng.adjust(a, 5, (30.0, -20.0, 10.0), out=b)  # PFE_ADJ_HSL
            eng.sharpen(b, 1.0, 2.0, out=a)

        ms = timed(pipe, args.steps)
        # each stage on its own, same buffers (CUDA events around a loop of that stage only)
        stages = {"gaussian sigma=50": timed(lambda: eng.gaussian_blur(img, 50.0, out=a), args.steps),
                  "adjust HSL": timed(lambda: eng.adjust(a, 5, (30.0, -20.0, 10.0), out=b), args.steps),
                  "sharpen(1.0, 2.0)": timed(lambda: eng.sharpen(b, 1.0, 2.0, out=a), args.steps)}
        emit({"config": 3, "workload": "8K Gaussian sigma=50 -> HSL(30,-20,10) -> sharpen(1.0, 2.0)", "ms": ms,
              "mpx_s": px8k / ms / 1e3, "stages_ms": stages, "n_gpus": 1})
        avg = stages["adjust HSL"]
        emit({"config": 3, "kernel": "adjust (HSL)", "avg_ms": avg, "achieved_gbs": 8 * px8k / (avg * 1e-3) / 1e9,
              "frac_of_hbm": 8 * px8k / (avg * 1e-3) / 1e9 / peak})
        del img, a, b

    if 4 in configs:  # mesh warp 6x6 + liquify on a big canvas, row bands across ranks
        S = args.canvas
        bounds = pd.band_bounds(S, world)
        y0, y1 = bounds[rank]
        band = torch.randint(0, 256, (y1 - y0, S, 4), dtype=torch.uint8, device=dev, generator=gen)
        orig = np.zeros((49, 2), np.float32)
        for r in range(7):
            for c in range(7):
                orig[r * 7 + c] = (np.float32(c) / np.float32(6) * S, np.float32(r) / np.float32(6) * S)
        deformed = orig.copy()
        for i in range(7):
            for j in range(7):
                deformed[i * 7 + j] += np.float32(8.0 * np.sin(i) * np.cos(j))
        # liquify field: 64 pushes, r=200, strength 0.8; built per band on device
        field = torch.zeros((S if world == 1 else y1 - y0, S, 2), dtype=torch.float32, device=dev)
        prng = np.random.default_rng(0x5EED)
        pushes = [(float(prng.uniform(0, S)), float(prng.uniform(0, S)), float(prng.uniform(-20, 20)), float(prng.uniform(-20, 20))) for _ in range(64)]
        if world == 1:
            for cx, cy, dx, dy in pushes:
                eng.liquify(field, 0, cx, cy, 200.0, 0.8, dx, dy)
            fband = field
        else:
            # the field of a band is the band of the field: shift brush centres into band coordinates
            for cx, cy, dx, dy in pushes:
                eng.liquify(field, 0, cx, cy - y0, 200.0, 0.8, dx, dy)
            fband = field

        def step():
            m = pd.mesh_warp_banded(eng, band, orig, deformed, 6, 6, S, S, bounds=bounds)
            return pd.warp_displacement_banded(eng, m, fband, S, bounds=bounds)

        ms = timed(step, max(3, args.steps // 4), warmup=2, world=world)
        emit({"config": 4, "workload": f"{S}x{S} mesh warp 6x6 (fused Catmull-Rom) + liquify warp, row bands + halo exchange",
              "ms": ms, "mpx_s": S * S / ms / 1e3, "n_gpus": world, "band_rows": y1 - y0})
        del band, field

    if 5 in configs:  # batch of 4K images through process.rhai, sharded by index
        script = "apply_blur(4.0); apply_hsl(10.0, 15.0, 0.0); apply_vignette(0.5, 0.3);"
        w, h = 3840, 2160
        n_total = args.batch_per_gpu * world
        mine = pd.shard_indices(n_total, rank, world)
        img = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)

        def run_batch():
            for k in mine:  # synthetic image generated on the fly from seed+index, no disk I/O
                g = torch.Generator(device=dev).manual_seed(0x5EED + k)
                img.random_(0, 256, generator=g)
                execute_script_sync(eng, script, img)

        ms = timed(run_batch, 1, warmup=1, world=world)
        emit({"config": 5, "workload": f"{n_total} x 4K images, script: {script}", "ms_total": ms, "images_s": n_total / ms * 1e3,
              "mpx_s": n_total * w * h / ms / 1e3, "n_gpus": world, "tier": "device-resident (images generated on the device)",
              "note": "includes on-device synthetic image generation"})

        # ---- end to end from HOST buffers through the batch pipeline the CLI uses (paintfe_b200/pipeline.py):
        # pinned RGBA in, RGBA out; decoding / encoding excluded (harness plumbing, PIL)
        from paintfe_b200.pipeline import ImagePipeline

        pool = [torch.randint(0, 256, (h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(8)]  # distinct inputs, reused in turn
        pipe = ImagePipeline(eng, depth=3)
        work = lambda dev: execute_script_sync(eng, script, dev)
        checksum = [0]

        def run_e2e():
            for k in range(len(mine)):
                if pipe.full():
                    _, out = pipe.collect()
                    checksum[0] += int(out[0, 0, 0])  # the result is on the host (touch it)
                pipe.submit(pool[k % len(pool)], work, tag=k)
            for _, out in pipe.drain():
                checksum[0] += int(out[0, 0, 0])

        def wall(fn):
            import time
            fn()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t[0])
            return dt

        ms_e2e = wall(run_e2e)
        # the PCIe bound of the same traffic: one 4K image up and one down per image, both directions at once
        dev_a, dev_b = torch.empty((h, w, 4), dtype=torch.uint8, device=dev), torch.empty((h, w, 4), dtype=torch.uint8, device=dev)
        host_out = torch.empty((h, w, 4), dtype=torch.uint8).pin_memory()
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def bare():
            for k in range(len(mine)):
                with torch.cuda.stream(s_up):
                    dev_a.copy_(pool[k % len(pool)], non_blocking=True)
                with torch.cuda.stream(s_dn):
                    host_out.copy_(dev_b, non_blocking=True)

        ms_bare = wall(bare)
        emit({"config": 5, "workload": f"{n_total} x 4K images, script: {script}", "tier": "e2e from pinned host buffers through ImagePipeline (3 streams, depth 3)",
              "ms_total": ms_e2e, "images_s": n_total / ms_e2e * 1e3, "mpx_s": n_total * w * h / ms_e2e / 1e3, "n_gpus": world,
              "h2d_bytes_per_image": w * h * 4, "d2h_bytes_per_image": w * h * 4,
              "bare_copy_ms_total": ms_bare, "bare_copy_images_s": n_total / ms_bare * 1e3,
              "frac_of_pcie_bound": ms_bare / ms_e2e,
              "note": "bare copy = the same uploads and downloads (both directions concurrently) with no compute; codec excluded"})
        del pool

    if 6 in configs:  # tile-native flatten: dense and sparse 8K stacks, device-resident and host tier
        import time

        cyn, cxn = (H8K + 63) // 64, (W8K + 63) // 64
        out = torch.empty((H8K, W8K, 4), dtype=torch.uint8, device=dev)
        for name, fill in (("every chunk populated", 1.0), ("25% of each layer's chunks populated", 0.25)):
            prng = np.random.default_rng(7)
            flat_layers, tiled, metas = [], [], []
            for i in range(16):
                t = torch.randint(0, 256, (H8K, W8K, 4), dtype=torch.uint8, device=dev, generator=gen)
                if fill < 1.0:
                    keep = torch.from_numpy(prng.random((cyn, cxn)) < fill).to(dev)
                    big = keep.repeat_interleave(64, 0).repeat_interleave(64, 1)[:H8K, :W8K]
                    t[~big] = 0
                flat_layers.append(t)
                tiled.append(eng.tiled(W8K, H8K).from_flat(t))
                metas.append(dict(blend=i % 25, opacity=0.25 + 0.05 * i))
            dl = [make_layer(t, **m) for t, m in zip(flat_layers, metas)]
            tl = [dict(m, tiles=t) for t, m in zip(tiled, metas)]
            ms_dense = timed(lambda: eng.flatten(dl, W8K, H8K, out=out), args.steps)
            ref = out.clone()
            ms_tiles = timed(lambda: eng.flatten_tiles(tl, W8K, H8K, out=out), args.steps)
            same = bool(torch.equal(ref, out))
            populated = sum(int(t.download(False)[0].sum()) for t in tiled)
            emit({"config": 6, "workload": f"8K 16-layer flatten, {name}", "dense_kernel_ms": ms_dense, "tile_kernel_ms": ms_tiles,
                  "mpx_s_tiles": px8k / ms_tiles / 1e3, "populated_chunks": populated, "identical": same, "n_gpus": 1})
            # host tier: chunk tables of host pointers vs. dense host layers (pageable numpy memory on both sides)
            host_tabs = []
            for t in tiled:
                occ, tiles = t.download(True)
                host_tabs.append([tiles[k] if occ[k] else None for k in range(occ.size)])
            host_flat = [t.cpu().numpy() for t in flat_layers]
            hl_t = [dict(m, tiles=tab) for tab, m in zip(host_tabs, metas)]
            hl_d = [make_layer(a, **m) for a, m in zip(host_flat, metas)]
            res = {}
            for key, fn in (("host_tier_tiles_ms", lambda: eng.flatten_tiles(hl_t, W8K, H8K)), ("host_tier_dense_ms", lambda: eng.flatten(hl_d, W8K, H8K))):
                fn()
                t0 = time.perf_counter()
                for _ in range(3):
                    r = fn()
                res[key] = (time.perf_counter() - t0) / 3 * 1e3
            emit({"config": 6, "workload": f"host tier, {name}", **res, "h2d_bytes_tiles": populated * 16384, "h2d_bytes_dense": 16 * px8k * 4,
                  "note": "pageable host memory; wall clock incl. the Python marshalling of the chunk tables"})
            for t in tiled:
                t.close()
            del flat_layers, tiled, dl, tl, host_tabs, host_flat, hl_t, hl_d

    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
