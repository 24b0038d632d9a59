"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing (NCCL on GPUs,
gloo in the CPU tests).

The reference has no multi-GPU or multi-process compute at all; this module is the
B200-native answer to its two scaling axes:

  * a batch of independent images (`paintfe -i glob ...`, the serial loop at src/cli.rs:159-209)
    shards by image index with NO communication - `shard_indices`;
  * one canvas too large for a single GPU splits into contiguous row bands (64-row aligned, so a
    band is a whole number of TiledImage chunk rows).  Flatten and per-pixel adjustments are band-
    local.  Blurs need `ceil(3*sigma)` rows from each neighbour, warps need as many rows as the
    displacement reaches: ONE point-to-point halo exchange of u8 source rows per op, then the
    ordinary single-GPU kernel runs on the extended band.  That is the only collective on the path.

Every function takes the compute engine as its first argument (paintfe_b200.engine.Engine on a
GPU; the CPU tests pass a stand-in built on the oracle) and never computes pixels itself.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

CHUNK = 64


def band_bounds(h: int, world: int, align: int = CHUNK) -> List[Tuple[int, int]]:
    """Contiguous row bands [y0, y1) covering h rows, sizes as equal as `align` allows.
    Ranks beyond the number of aligned units get an empty band."""
    units = (h + align - 1) // align
    out, u0 = [], 0
    for r in range(world):
        n = units // world + (1 if r < units % world else 0)
        y0, y1 = min(u0 * align, h), min((u0 + n) * align, h)
        out.append((y0, y1))
        u0 += n
    return out


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Image k -> rank k mod world (CLI batch): independent units, no collective."""
    return list(range(rank, n_items, world))


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin this process (CPU affinity + preferred memory node) to the NUMA node its GPU hangs off, so the
    pinned staging buffers of the host tier are allocated next to the GPU's PCIe root port.  With one
    process per GPU on a two-socket box this keeps every rank's H2D traffic off the inter-socket link.
    Best effort: returns what was done, never raises."""
    import ctypes
    import os

    info = {"node": None, "cpus": None}
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["cpus"] = len(cpus)
        mask = ctypes.c_ulong(1 << node)
        libc = ctypes.CDLL(None, use_errno=True)
        if libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64)) == 0:  # set_mempolicy(MPOL_PREFERRED)
            info["node"] = node
    except Exception:
        pass
    return info


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def exchange_halo(band: torch.Tensor, halo_up: int, halo_down: int, bounds: Sequence[Tuple[int, int]], group=None):
    """Point-to-point halo exchange of whole rows between row-band neighbours.

    `band` is this rank's rows (rows, ...) of an image split by `bounds`; returns (top, bottom):
    up to `halo_up` rows that sit directly above the band and up to `halo_down` rows directly
    below it (fewer at the image edges).  A halo may span several neighbours when bands are
    thinner than the halo; every transfer is one isend/irecv pair, batched in one group so NCCL
    issues them as a single ncclGroupStart/End.
    """
    rank, world = _world(group)
    y0, y1 = bounds[rank]
    h = bounds[-1][1]
    want_top = (max(y0 - halo_up, 0), y0)
    want_bot = (y1, min(y1 + halo_down, h))
    row_shape = tuple(band.shape[1:])

    def overlap(a, b):
        lo, hi = max(a[0], b[0]), min(a[1], b[1])
        return (lo, hi) if hi > lo else None

    ops, recvs = [], []
    for peer in range(world):
        if peer == rank:
            continue
        py0, py1 = bounds[peer]
        # rows the peer wants from me (same halo sizes everywhere)
        for want in ((max(py0 - halo_up, 0), py0), (py1, min(py1 + halo_down, h))):
            ov = overlap(want, (y0, y1))
            if ov:
                ops.append(dist.P2POp(dist.isend, band[ov[0] - y0:ov[1] - y0].contiguous(), peer, group))
        # rows I want from the peer
        for which, want in (("top", want_top), ("bot", want_bot)):
            ov = overlap(want, (py0, py1))
            if ov:
                buf = torch.empty((ov[1] - ov[0],) + row_shape, dtype=band.dtype, device=band.device)
                ops.append(dist.P2POp(dist.irecv, buf, peer, group))
                recvs.append((which, ov[0], buf))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    top = [b for w, _, b in sorted((r for r in recvs if r[0] == "top"), key=lambda r: r[1])]
    bot = [b for w, _, b in sorted((r for r in recvs if r[0] == "bot"), key=lambda r: r[1])]
    empty = band[:0]
    return (torch.cat(top) if top else empty), (torch.cat(bot) if bot else empty)


def _as_tensor(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def gaussian_radius(sigma: float) -> int:
    """`(sigma * 3.0).ceil() as usize` in f32, src/ops/filters.rs:215."""
    v = float(np.ceil(np.float32(sigma) * np.float32(3.0)))
    return int(v) if v > 0 else 0


def _windowed(eng, op, band, halo, bounds, group):
    """Run a windowed filter on a band: exchange `halo` rows, filter the extended band, keep the core."""
    band = _as_tensor(band)
    top, bot = exchange_halo(band, halo, halo, bounds, group)
    ext = torch.cat([top, band, bot]) if (len(top) or len(bot)) else band
    if ext.shape[0] == 0:
        return band
    out = _as_tensor(op(ext if ext.is_cuda else ext.numpy()))
    return out[len(top):len(top) + band.shape[0]]


def gaussian_blur_banded(eng, band, h_total: int, sigma: float, exact: bool = False, group=None, bounds=None):
    """parallel_gaussian_blur (filters.rs:242-316) of a row-split image. The H pass is band-local; the
    V pass reads `r = ceil(3 sigma)` rows either side, so each rank receives r rows of *u8 input*
    from its neighbours (4 B/px instead of exchanging the 16 B/px f32 intermediate) and recomputes
    the H pass on them.  Clamp-to-edge at the true image border falls out of the extended band's own
    edges because only ranks at the border lack a halo there.  Bit-identical to the unsplit blur."""
    _, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    return _windowed(eng, lambda ext: eng.gaussian_blur(ext, sigma, exact=exact), band, gaussian_radius(sigma), bounds, group)


def box_blur_banded(eng, band, h_total: int, radius: float, group=None, bounds=None):
    """box_blur_core (effects/blur.rs:233-318); halo = ceil(radius)."""
    _, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    r = int(math.ceil(radius)) if radius >= 0.5 else 0
    return _windowed(eng, lambda ext: eng.box_blur(ext, radius), band, r, bounds, group)


def median_banded(eng, band, h_total: int, radius: int, group=None, bounds=None):
    """median_core (effects/noise.rs:357-410); halo = max(radius, 1)."""
    _, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    return _windowed(eng, lambda ext: eng.median(ext, radius), band, max(int(radius), 1), bounds, group)


def sharpen_banded(eng, band, h_total: int, amount: float, radius: float, exact: bool = False, group=None, bounds=None):
    """sharpen_core (effects/stylize.rs:96-141): unsharp mask = blur(sigma=radius) + per-pixel epilogue."""
    _, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    return _windowed(eng, lambda ext: eng.sharpen(ext, amount, radius, exact=exact), band, gaussian_radius(radius), bounds, group)


def neighbourhood_banded(eng, band, h_total: int, op: str, args: tuple = (), halo: int = 0, group=None, bounds=None, **kw):
    """Any translation-invariant windowed effect of a row-split image: exchange `halo` input rows, run
    `eng.<op>(extended band, *args)`, keep the core rows.  Valid for effects whose result at a pixel
    depends only on the pixels within `halo` rows of it and not on absolute coordinates - ink (halo 1),
    oil painting (halo = radius), bokeh (ceil(radius)), bilateral reduce_noise (radius), motion blur
    (ceil(distance) + 1), glow (ceil(3 sigma)), rgb_displace (max |dy|).  Effects that use absolute
    coordinates (vignette, halftone, zoom, noise, contours, crystallize, pixel_drag, dents) do not split
    this way; whole images shard across GPUs instead."""
    _, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    return _windowed(eng, lambda ext: getattr(eng, op)(ext, *args, **kw), band, int(halo), bounds, group)


def ink_banded(eng, band, h_total, edge_strength, threshold, group=None, bounds=None):
    """ink_core (effects/artistic.rs:31-99): 3x3 Sobel."""
    return neighbourhood_banded(eng, band, h_total, "ink", (edge_strength, threshold), 1, group, bounds)


def oil_painting_banded(eng, band, h_total, radius, levels, group=None, bounds=None):
    """oil_painting_core (artistic.rs:123-217): window radius clamps to 1..10."""
    return neighbourhood_banded(eng, band, h_total, "oil_painting", (radius, levels), min(max(int(radius), 1), 10), group, bounds)


def bokeh_blur_banded(eng, band, h_total, radius, group=None, bounds=None):
    """bokeh_blur_core (effects/blur.rs:22-115): disc of ceil(radius) rows either side."""
    return neighbourhood_banded(eng, band, h_total, "bokeh_blur", (radius,), int(math.ceil(radius)) if radius >= 0.5 else 0, group, bounds)


def reduce_noise_banded(eng, band, h_total, strength, radius, group=None, bounds=None):
    """reduce_noise_core (effects/noise.rs:172-262)."""
    return neighbourhood_banded(eng, band, h_total, "reduce_noise", (strength, radius), max(int(radius), 1), group, bounds)


def motion_blur_banded(eng, band, h_total, angle_deg, distance, group=None, bounds=None):
    """motion_blur_core (effects/blur.rs:144-210): samples reach ceil(distance) pixels along the direction."""
    return neighbourhood_banded(eng, band, h_total, "motion_blur", (angle_deg, distance), int(math.ceil(abs(distance))) + 1, group, bounds)


def flatten_banded(eng, layer_bands, w: int, band_rows: int, active=None):
    """CanvasState::composite on this rank's band of every layer: pixel-independent, no collective."""
    return eng.flatten(layer_bands, w, band_rows, active=active)


def _warp_band(eng, band, h_total, w_out, y0, rows_out, reach, bounds, group, uniform=False, **warp_kw):
    """`reach` = (up, down) halo rows this rank needs: python ints, or a 2-element device tensor. One
    max-reduction (a single collective and a single host read) sizes the halo identically on every rank
    (SURVEY 8e); `uniform` = every rank already holds the same numbers (mesh warps), no collective at all."""
    band = _as_tensor(band)
    rank, world = _world(group)
    if world > 1 and not uniform:
        t = reach if isinstance(reach, torch.Tensor) else torch.tensor(list(reach), dtype=torch.int32, device=band.device)
        if t.device.type == "cpu" and dist.get_backend(group) == "nccl":
            t = t.to(band.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        up, down = (int(v) for v in t.tolist())
    else:
        up, down = (int(v) for v in (reach.tolist() if isinstance(reach, torch.Tensor) else reach))
    top, bot = exchange_halo(band, up, down, bounds, group)
    window = torch.cat([top, band, bot]) if (len(top) or len(bot)) else band
    src_y0 = bounds[rank][0] - len(top)
    out = eng.warp_band(window if window.is_cuda else window.numpy(), h_total, src_y0, w_out, h_total, y0, rows_out, **warp_kw)
    return _as_tensor(out)


def warp_displacement_banded(eng, band, disp_band, h_total: int, group=None, bounds=None):
    """warp_displacement_full (transform.rs:1288-1345) on a row-split canvas (source and output share
    the split). Reach = how far (y - dy) leaves the band, taken from the band's own field: on the device in
    one pass (pfe_dev_disp_reach), folded into the ranks' max-reduction before the host reads it."""
    rank, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    y0, y1 = bounds[rank]
    d = _as_tensor(disp_band)
    if world == 1:
        reach = (0, 0)  # the band is the whole image: nothing to exchange, nothing to size
    elif y1 > y0 and d.is_cuda and hasattr(eng, "disp_reach"):
        mm = eng.disp_reach(d, y0, h_total)  # [min, max] of floor(clamp(y - dy, -1, h))
        reach = torch.stack([(y0 - mm[0]).clamp(min=0), (mm[1] + 2 - y1).clamp(min=0)]).to(torch.int32)
    elif y1 > y0:
        ys = torch.arange(y0, y1, dtype=torch.float32, device=d.device)[:, None]
        sy = ys - torch.nan_to_num(d[..., 1], nan=0.0, posinf=0.0, neginf=0.0)
        sy = sy.clamp(-1.0, float(h_total))
        reach = (math.ceil(max(0.0, y0 - float(torch.floor(sy.min())))), math.ceil(max(0.0, float(torch.floor(sy.max())) + 2 - y1)))
    else:
        reach = (0, 0)
    return _warp_band(eng, band, h_total, int(d.shape[1]), y0, y1 - y0, reach, bounds, group,
                      disp_band=d if d.is_cuda else d.numpy())


def mesh_reach(original, deformed) -> int:
    """Upper bound on |dy| of generate_displacement_from_mesh: the field is the Catmull-Rom surface of
    (deformed - original) and a 2-D cardinal-spline weight set has L1 norm <= 1.25^2."""
    o = np.asarray(original, np.float32).reshape(-1, 2)
    d = np.asarray(deformed, np.float32).reshape(-1, 2)
    return int(math.ceil(1.5625 * float(np.abs(d[:, 1] - o[:, 1]).max()))) + 2


def mesh_warp_banded(eng, band, original, deformed, cols: int, rows: int, w: int, h_total: int, group=None, bounds=None):
    """warp_mesh_catmull_rom (transform.rs:1743-1761) on a row-split canvas, displacement fused."""
    rank, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    y0, y1 = bounds[rank]
    reach = mesh_reach(original, deformed)  # from the control points alone: identical on every rank
    return _warp_band(eng, band, h_total, w, y0, y1 - y0, (reach, reach), bounds, group, uniform=True,
                      original=original, deformed=deformed, cols=cols, rows=rows)
