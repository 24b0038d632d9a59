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
import os
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


class HaloPlan:
    """A band extended by its halo rows, allocated once, plus the point-to-point transfers that fill it.

    `ext` holds [top halo | band | bottom halo] rows; `core` is the view of the band's own rows - producers write
    into it directly (e.g. `eng.flatten(..., out=plan.core)`), so no band is ever concatenated or copied to make
    room for its halo.  The halo may span several neighbours when bands are thinner than it; every transfer is one
    isend / irecv of a contiguous row range of `ext`, all batched in one group (one ncclGroupStart/End).

    On CUDA `exchange_async()` runs the group on a side stream: the caller keeps enqueueing band-local work (the
    interior H pass of a blur) on its own stream and calls `wait()` right before the first kernel that reads halo
    rows.  `wait()` also orders the sends' reads of `core` before whatever overwrites `core` next.
    """

    def __init__(self, row_shape, dtype, device, halo_up: int, halo_down: int, bounds: Sequence[Tuple[int, int]], group=None):
        self.group = group
        self.rank, self.world = _world(group)
        self.bounds = list(bounds)
        y0, y1 = self.bounds[self.rank]
        h = self.bounds[-1][1]
        self.y0, self.rows = y0, y1 - y0
        want_top = (max(y0 - halo_up, 0), y0)
        want_bot = (y1, min(y1 + halo_down, h))
        self.top, self.bot = want_top[1] - want_top[0], want_bot[1] - want_bot[0]
        self.ext = torch.empty((self.top + self.rows + self.bot,) + tuple(row_shape), dtype=dtype, device=device)
        self.core = self.ext[self.top:self.top + self.rows]
        ext_y0 = y0 - self.top  # image row of ext row 0

        def overlap(a, b):
            lo, hi = max(a[0], b[0]), min(a[1], b[1])
            return (lo, hi) if hi > lo else None

        self.sends, self.recvs = [], []  # (peer, ext row lo, ext row hi)
        for peer in range(self.world):
            if peer == self.rank:
                continue
            py0, py1 = self.bounds[peer]
            for want in ((max(py0 - halo_up, 0), py0), (py1, min(py1 + halo_down, h))):  # rows the peer wants from me
                ov = overlap(want, (y0, y1))
                if ov:
                    self.sends.append((peer, ov[0] - ext_y0, ov[1] - ext_y0))
            for want in (want_top, want_bot):  # rows I want from the peer
                ov = overlap(want, (py0, py1))
                if ov:
                    self.recvs.append((peer, ov[0] - ext_y0, ov[1] - ext_y0))
        self.halo_bytes = sum((b - a) for _, a, b in self.recvs) * int(np.prod(row_shape)) * self.ext.element_size()
        self._side = self._ready = self._done = None
        self._op_list = None
        if self.ext.is_cuda:
            self._side = torch.cuda.Stream(device=device)
            self._ready = torch.cuda.Event()
            self._done = torch.cuda.Event()

    def _ops(self):
        if self._op_list is None:  # the buffers never move: build the descriptors once
            ops = [dist.P2POp(dist.isend, self.ext[a:b], peer, self.group) for peer, a, b in self.sends]
            ops += [dist.P2POp(dist.irecv, self.ext[a:b], peer, self.group) for peer, a, b in self.recvs]
            self._op_list = ops
        return self._op_list

    def exchange(self):
        """Fill the halo rows; on return (CPU) / in stream order (CUDA) `ext` is complete."""
        self.exchange_async()
        self.wait()

    def exchange_async(self):
        ops = self._ops()
        if not ops:
            self._pending = False
            return
        self._pending = True
        if self._side is None:  # CPU (gloo): nothing to overlap with
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            self._pending = False
            return
        cur = torch.cuda.current_stream(self.ext.device)
        self._ready.record(cur)  # core is written, earlier readers of the halo rows are done
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ready)
            for req in dist.batch_isend_irecv(ops):
                req.wait()  # NCCL: orders the side stream after the transfers, the host does not block
            self._done.record(self._side)

    def wait(self):
        if getattr(self, "_pending", False) and self._side is not None:
            torch.cuda.current_stream(self.ext.device).wait_event(self._done)
        self._pending = False

    def load(self, band: torch.Tensor):
        """Put the band's rows into `core` (a no-op when the caller produced them there)."""
        if band.data_ptr() != self.core.data_ptr():
            self.core.copy_(band)
        return self


_plans = {}


def halo_plan(band: torch.Tensor, halo_up: int, halo_down: int, bounds, group=None, cache: bool = True) -> HaloPlan:
    """The (cached) HaloPlan for bands shaped like `band`. A cached plan's buffers are reused by the next call with
    the same geometry: results that alias `plan.ext` must be consumed (or copied) before then."""
    key = (tuple(band.shape[1:]), band.dtype, str(band.device), int(halo_up), int(halo_down), tuple(bounds), id(group))
    plan = _plans.get(key) if cache else None
    if plan is None:
        plan = HaloPlan(tuple(band.shape[1:]), band.dtype, band.device, halo_up, halo_down, bounds, group)
        if cache:
            if len(_plans) > 16:
                _plans.clear()
            _plans[key] = plan
    return plan


def exchange_halo(band: torch.Tensor, halo_up: int, halo_down: int, bounds: Sequence[Tuple[int, int]], group=None):
    """Point-to-point halo exchange of whole rows between row-band neighbours: returns (top, bottom), up to
    `halo_up` rows that sit directly above the band and up to `halo_down` rows directly below it (fewer at the
    image edges)."""
    plan = halo_plan(band, halo_up, halo_down, bounds, group, cache=False).load(band)
    plan.exchange()
    return plan.ext[:plan.top], plan.ext[plan.top + plan.rows:]


def _as_tensor(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def gaussian_radius(sigma: float) -> int:
    """`(sigma * 3.0).ceil() as usize` in f32, src/ops/filters.rs:215."""
    v = float(np.ceil(np.float32(sigma) * np.float32(3.0)))
    return int(v) if v > 0 else 0


def _windowed(eng, op, band, halo, bounds, group):
    """Run a windowed filter on a band: exchange `halo` rows straight into the pre-allocated extended band, filter
    it, keep the core rows."""
    band = _as_tensor(band)
    plan = halo_plan(band, halo, halo, bounds, group).load(band)
    plan.exchange()
    if plan.ext.shape[0] == 0:
        return band
    out = _as_tensor(op(plan.ext if plan.ext.is_cuda else plan.ext.numpy()))
    return out[plan.top:plan.top + plan.rows]


def gaussian_blur_banded(eng, band, h_total: int, sigma: float, exact: bool = False, group=None, bounds=None, out=None):
    """parallel_gaussian_blur (filters.rs:242-316) of a row-split image. The H pass is band-local; the
    V pass reads `r = ceil(3 sigma)` rows either side, so each rank receives r rows of *u8 input*
    from its neighbours (4 B/px instead of exchanging the 16 B/px f32 intermediate) and recomputes
    the H pass on them.  Clamp-to-edge at the true image border falls out of the extended band's own
    edges because only ranks at the border lack a halo there.  Bit-identical to the unsplit blur.

    On the GPU engine the exchange runs on a side stream while the band's own rows go through the H pass; the halo
    rows are filtered when they have landed, then the V pass produces the band's rows (pfe_dev_gaussian_band_h/_v).
    `band` may be `halo_plan(...).core` itself (the producer wrote it there): then nothing is copied at all."""
    rank, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    r = gaussian_radius(sigma)
    band = _as_tensor(band)
    if not (band.is_cuda and hasattr(eng, "gaussian_band_h")) or r <= 16 or band.shape[0] == 0:
        # CPU stand-ins, empty bands, and radii small enough for the fused H+V kernel: filter the extended band whole
        res = _windowed(eng, lambda ext: eng.gaussian_blur(ext, sigma, exact=exact), band, r, bounds, group)
        if out is not None:
            out.copy_(res)
            return out
        return res
    plan = halo_plan(band, r, r, bounds, group).load(band)
    plan.exchange_async()
    eng.gaussian_band_h(plan.ext, plan.top, plan.rows, sigma, exact=exact)  # overlaps the exchange
    plan.wait()
    if plan.top:
        eng.gaussian_band_h(plan.ext, 0, plan.top, sigma, exact=exact)
    if plan.bot:
        eng.gaussian_band_h(plan.ext, plan.top + plan.rows, plan.bot, sigma, exact=exact)
    return eng.gaussian_band_v(plan.ext, plan.top, plan.rows, sigma, exact=exact, out=out)


class PeerUnavailable(RuntimeError):
    """The neighbours' memory cannot be mapped here (no CUDA IPC between the ranks, or bands thinner than the halo)."""


class _RawDeviceMemory:
    def __init__(self, addr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (addr, False), "version": 3, "strides": None}


def _device_view(addr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(_RawDeviceMemory(addr, nbytes), device=device)


def peer_targets(infos, rank: int, up_base: Optional[int], down_base: Optional[int], row_bytes: int, header: int = 4096):
    """Where a rank's edge rows go (pure address arithmetic, tested on the CPU).  infos[r] = (handle, top, rows, bot, stride)
    of rank r's block [flags | pad to `header`][ext buffer 0][ext buffer 1], ext = [top halo | band | bottom halo] rows;
    up_base / down_base = where the neighbours' blocks are mapped here (None at the image edges).  Returns
    (rows_up, put_up, rows_down, put_down): my first rows_up rows are the upper neighbour's bottom halo, my last rows_down
    rows the lower neighbour's top halo; put_x[p] = (address of those halo rows in buffer p, address of the flag to release):
    flag u32[p * 2 + side], side 0 = written by the upper neighbour of the flag's owner, side 1 = by the lower."""
    rows_up = rows_down = 0
    put_up = put_down = None
    if up_base is not None:
        _, top_u, rows_u, bot_u, stride_u = infos[rank - 1]
        rows_up = bot_u
        put_up = [(up_base + header + p * stride_u + (top_u + rows_u) * row_bytes, up_base + 4 * (p * 2 + 1)) for p in (0, 1)]
    if down_base is not None:
        _, top_d, _rows_d, _bot_d, stride_d = infos[rank + 1]
        rows_down = top_d
        put_down = [(down_base + header + p * stride_d, down_base + 4 * (p * 2 + 0)) for p in (0, 1)]
    return rows_up, put_up, rows_down, put_down


class PeerHalo:
    """A HaloPlan's extended band, twice, in device memory the two row neighbours have mapped (CUDA IPC -> NVLink peer
    access), plus one u32 flag per buffer and side.  Nothing is sent or received: the neighbour's flatten kernel stores
    its edge rows straight into this rank's halo rows and releases the flag (pfe_dev_flatten_peer), and this rank's
    stream waits for the flag right before its first kernel that reads a halo row (pfe_dev_peer_wait).

    Two buffers, used alternately, replace an acknowledgement: a neighbour writes buffer p again at step k+2, after it
    has seen this rank's flag of step k+1, which this rank's stream released after its reads of step k.

        block:  [ flags u32[2 buffers][2 sides] | pad to 4 KB ][ ext buffer 0 ][ ext buffer 1 ]
        side 0 = written by the upper neighbour (top halo), side 1 = by the lower neighbour (bottom halo)
    """
    HEADER = 4096

    def __init__(self, eng, plan: HaloPlan, halo: int, group=None):
        self.eng, self.plan, self.group = eng, plan, group
        self.rank, self.world = plan.rank, plan.world
        self.addr = self.up = self.down = None
        if not plan.ext.is_cuda or self.world < 2:
            raise PeerUnavailable("peer halos need CUDA and more than one rank")
        if min(b - a for a, b in plan.bounds) < halo:  # same answer on every rank: decided before anything is allocated
            raise PeerUnavailable("a band is thinner than the halo: rows would come from beyond the adjacent neighbour")
        shape = tuple(plan.ext.shape)
        self.row_bytes = int(np.prod(shape[1:])) * plan.ext.element_size()
        ext_bytes = shape[0] * self.row_bytes
        self.stride = (ext_bytes + 4095) // 4096 * 4096
        err = None
        try:
            self.addr, handle = eng.peer_alloc(self.HEADER + 2 * self.stride)
        except Exception as e:  # noqa: BLE001 - whatever it is, every rank must learn of it
            err, handle = str(e), None
        infos = self._gather((handle, plan.top, plan.rows, plan.bot, self.stride))
        if err is None and all(i[0] is not None for i in infos):
            try:
                if self.rank > 0:
                    self.up = eng.peer_open(infos[self.rank - 1][0])
                if self.rank + 1 < self.world:
                    self.down = eng.peer_open(infos[self.rank + 1][0])
            except Exception as e:  # noqa: BLE001
                err = str(e)
        elif err is None:
            err = "a neighbour could not allocate"
        errs = self._gather(err)
        if any(errs):  # every rank takes this path together: unmap, wait for the others to have unmapped, free
            self._release()
            dist.barrier(group=self.group)
            if self.addr is not None:
                eng.peer_free(self.addr)
                self.addr = None
            raise PeerUnavailable("; ".join(sorted({e for e in errs if e})))
        dev = plan.ext.device
        self.ext = [_device_view(self.addr + self.HEADER + p * self.stride, ext_bytes, dev).view(plan.ext.dtype).view(shape) for p in (0, 1)]
        self.core = [e[plan.top:plan.top + plan.rows] for e in self.ext]
        self.flags = _device_view(self.addr, 16, dev).view(torch.int32)
        self.rows_up, self.put_up, self.rows_down, self.put_down = peer_targets(infos, self.rank, self.up, self.down, self.row_bytes)
        self.halo_bytes = (plan.top + plan.bot) * self.row_bytes

    def _gather(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def wait_args(self, p: int) -> Tuple[int, int]:
        """(address of the first flag, count) this rank waits on for buffer p."""
        first = 0 if self.up is not None else 1
        n = (self.up is not None) + (self.down is not None)
        return self.addr + 4 * (p * 2 + first), n

    def _release(self):
        for a in (self.up, self.down):
            if a is not None:
                self.eng.peer_close(a)
        self.up = self.down = None

    def close(self):
        """Collective: unmap the neighbours, then (once every rank has) free this rank's block."""
        if self.addr is None:
            return
        torch.cuda.synchronize()
        self.ext = self.core = self.flags = None
        self._release()
        dist.barrier(group=self.group)
        self.eng.peer_free(self.addr)
        self.addr = None


class BandedFlattenBlur:
    """composite() + parallel_gaussian_blur of ONE canvas on this rank's row band, scheduled so that the halo transfer
    hides: the band's edge rows (the ones the neighbours need) are flattened first, the interior flatten and the band's
    own H pass run while they travel; the halo rows are H-filtered when they have landed, then the V pass writes the
    band.  `layers` are this rank's band rows of every layer (device tensors in `rgba`, plus opacity / blend / ...).
    Descriptors and buffers are set up once; `step()` only enqueues.

    transport "peer": the edge rows go straight into the neighbours' halo rows over NVLink peer memory, followed by a
    flag (PeerHalo) - no send/receive, no rendezvous.  "nccl": batched isend/irecv on a side stream (HaloPlan).
    "auto" takes "peer" when every rank can map its neighbours, else "nccl" (the decision is collective)."""

    def __init__(self, eng, layers, w: int, h_total: int, sigma: float, exact: bool = False, group=None, bounds=None,
                 transport: str = "auto", timeout_ms: int = 2000, peer_put: Optional[str] = None):
        rank, world = _world(group)
        self.eng, self.sigma, self.exact = eng, float(sigma), bool(exact)
        self.bounds = bounds or band_bounds(h_total, world)
        y0, y1 = self.bounds[rank]
        self.rows, self.w, self.h_total = y1 - y0, int(w), int(h_total)
        r = gaussian_radius(sigma)
        first = next(L["rgba"] for L in layers if L.get("rgba") is not None)
        self.plan = halo_plan(first, r, r, self.bounds, group)
        self.out = torch.empty((self.rows, w, 4), dtype=torch.uint8, device=first.device)
        self.fused = r <= 16  # small radii: the fused H+V kernel on the extended band (see gaussian_blur_banded)
        self.peer, self.k, self.timeout_ms = None, 0, int(timeout_ms)
        self._side = self._ev = None
        if transport not in ("auto", "peer", "nccl"):
            raise ValueError("transport must be auto, peer or nccl")
        if transport != "nccl" and world > 1 and not self.fused and hasattr(eng, "flatten_prepared_peer"):
            try:
                self.peer = PeerHalo(eng, self.plan, r, group)
            except PeerUnavailable as e:
                self.peer_error = str(e)
                if transport == "peer":
                    raise
            except Exception as e:  # noqa: BLE001 - "auto" must end up with a working transport
                if transport == "peer":
                    raise
                self.peer_error = "%s: %s" % (type(e).__name__, e)
        elif transport == "peer":
            raise PeerUnavailable("peer transport needs the GPU engine, more than one rank and a two-pass radius")
        self.transport = "peer" if self.peer is not None else "nccl"
        # how the edge rows get into the neighbour's buffer: "copy" = a device-to-device copy to the mapped address behind
        # a plain flatten (a copy engine moves them while the SMs flatten the interior), "store" = the flatten kernel's
        # own second store (pfe_dev_flatten_peer).  Measured on 8 B200s, 8K canvas: copy 0.398 ms, store 0.404 ms, NCCL
        # 0.460 ms per step (profiles/r02_bench_n8.json).  PFE_PEER_PUT overrides the default.
        self.peer_put = peer_put or os.environ.get("PFE_PEER_PUT", "copy")
        if self.peer_put not in ("store", "copy"):
            raise ValueError("peer_put must be store or copy")
        self._put_views = {}
        e = min(r, self.rows)
        # row ranges of the band: [0, e) and [rows - e, rows) feed the neighbours; the interior is everything else
        if self.peer is not None:
            pr = self.peer
            cut_a, cut_b = pr.rows_up, self.rows - pr.rows_down
            if cut_a > cut_b:  # edges overlap (a band barely taller than the halo): the overlap is flattened twice
                self.parts = [(0, pr.rows_up, "up"), (self.rows - pr.rows_down, self.rows, "down")]
            else:
                self.parts = [(0, cut_a, "up"), (cut_b, self.rows, "down"), (cut_a, cut_b, None)]
            self.parts = [p for p in self.parts if p[1] > p[0]]
        elif self.rows > 2 * e and world > 1:
            self.parts = [(0, e, "up"), (self.rows - e, self.rows, "down"), (e, self.rows - e, None)]
        else:
            self.parts = [(0, self.rows, None)]

        def sub(a, b):
            ls = [dict(L, rgba=L["rgba"][a:b], mask=(None if L.get("mask") is None else L["mask"][a:b])) if L.get("rgba") is not None else L
                  for L in layers]
            return eng.prepare_layers(ls, w, b - a)

        self.prepared = [sub(a, b) for a, b, _ in self.parts]

    def close(self):
        self._put_views = {}
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    def _step_two_streams(self):
        """Two streams so that the small launches never leave the GPU idling behind their last wave (a 60-row edge
        flatten is 450 CTAs for 444 resident slots: alone on a stream it costs two waves):

            side (high priority)  flatten edge rows -> neighbours | wait for the neighbours' rows | H pass of the halo rows
            main                  flatten interior rows           | H pass of the band's rows     | V pass

        The edge CTAs are scheduled first and the interior CTAs fill the slots they leave; main joins side twice
        (band flattened; halo rows filtered).  Peer transport: the edge flatten's own stores carry the rows and the
        wait is a flag wait.  NCCL transport: a batched isend/irecv after the edge flatten."""
        eng, plan, pr = self.eng, self.plan, self.peer
        if pr is not None:
            self.k += 1
            p, value = self.k & 1, self.k
            ext, core = pr.ext[p], pr.core[p]
        else:
            ext, core = plan.ext, plan.core
        main = torch.cuda.current_stream(ext.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=ext.device, priority=-1)
            self._ev = [torch.cuda.Event() for _ in range(3)]
        side, (ev_prev, ev_flat, ev_halo) = self._side, self._ev
        ev_prev.record(main)  # the previous step's passes have read these rows and the H scratch
        with torch.cuda.stream(side):
            side.wait_event(ev_prev)
            for (a, b, to), prep in zip(self.parts, self.prepared):
                if to is None:
                    continue
                put = None if pr is None else pr.put_up if to == "up" else pr.put_down
                if put is None:
                    eng.flatten_prepared(prep, core[a:b])
                elif self.peer_put == "store":
                    eng.flatten_prepared_peer(prep, core[a:b], put[p][0], put[p][1], value)
                else:  # "copy": plain flatten, then a device-to-device copy into the neighbour's rows and the flag
                    eng.flatten_prepared(prep, core[a:b])
                    key = (to, p)
                    if key not in self._put_views:
                        rows_here = core[a:b]
                        self._put_views[key] = (_device_view(put[p][0], rows_here.numel() * rows_here.element_size(), rows_here.device),
                                                rows_here.reshape(-1).view(torch.uint8))
                    far, near = self._put_views[key]
                    far.copy_(near)
                    eng.peer_signal(put[p][1], value)
            ev_flat.record(side)
            if pr is None:
                plan.exchange_async()
        for (a, b, to), prep in zip(self.parts, self.prepared):
            if to is None:
                eng.flatten_prepared(prep, core[a:b])
        main.wait_event(ev_flat)
        with torch.cuda.stream(side):
            if pr is None:
                plan.wait()
            else:
                eng.peer_wait(*pr.wait_args(p), value, self.timeout_ms)
            if plan.top:
                eng.gaussian_band_h(ext, 0, plan.top, self.sigma, exact=self.exact)
            if plan.bot:
                eng.gaussian_band_h(ext, plan.top + plan.rows, plan.bot, self.sigma, exact=self.exact)
            ev_halo.record(side)
        eng.gaussian_band_h(ext, plan.top, plan.rows, self.sigma, exact=self.exact)
        main.wait_event(ev_halo)
        return eng.gaussian_band_v(ext, plan.top, plan.rows, self.sigma, exact=self.exact, out=self.out)

    def step(self):
        eng, plan = self.eng, self.plan
        if plan.ext.is_cuda and not self.fused and len(self.parts) > 1:
            return self._step_two_streams()
        for (a, b, _), prep in zip(self.parts, self.prepared):
            eng.flatten_prepared(prep, plan.core[a:b])
        plan.exchange_async()
        plan.wait()
        if self.fused or not hasattr(eng, "gaussian_band_h"):
            res = _as_tensor(eng.gaussian_blur(plan.ext if plan.ext.is_cuda else plan.ext.numpy(), self.sigma, exact=self.exact))
            self.out.copy_(res[plan.top:plan.top + plan.rows])
            return self.out
        eng.gaussian_band_h(plan.ext, 0, plan.ext.shape[0], self.sigma, exact=self.exact)
        return eng.gaussian_band_v(plan.ext, plan.top, plan.rows, self.sigma, exact=self.exact, out=self.out)


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


def adjust_banded(eng, band, h_total: int, op: int, params=(), luts=None, mask_band=None, occupancy=None, group=None, bounds=None):
    """apply_pixel_transform (adjustments.rs:21-42) on this rank's band: per pixel, so band-local - no collective.
    `occupancy` is the WHOLE canvas' chunk bitmap (ceil(h/64) x ceil(w/64), tiled_image.rs:905-933: unpopulated chunks
    are left alone); bands are 64-row aligned, so the band's chunk rows are a contiguous slice of it and a chunk never
    straddles two ranks."""
    rank, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    y0, y1 = bounds[rank]
    if y1 <= y0:
        return band
    occ = None
    if occupancy is not None:
        if y0 % CHUNK:
            raise ValueError("adjust_banded: bands must start on a chunk row when an occupancy bitmap is given")
        occ = occupancy[y0 // CHUNK:(y1 + CHUNK - 1) // CHUNK]
        occ = occ.contiguous() if isinstance(occ, torch.Tensor) else np.ascontiguousarray(occ)
    return eng.adjust(band, op, params, luts=luts, mask=mask_band, occupancy=occ)


def _warp_band(eng, band, h_total, w_out, y0, rows_out, reach, bounds, group, **warp_kw):
    """`reach` = (up, down) halo rows, python ints identical on every rank. The source rows land straight in the
    pre-allocated window; the band-form warp is stream-asynchronous (a window that turns out too small is
    reported by `eng.check_async()`)."""
    band = _as_tensor(band)
    rank, _ = _world(group)
    up, down = int(reach[0]), int(reach[1])
    plan = halo_plan(band, up, down, bounds, group).load(band)
    plan.exchange()
    window = plan.ext
    src_y0 = bounds[rank][0] - plan.top
    out = eng.warp_band(window if window.is_cuda else window.numpy(), h_total, src_y0, w_out, h_total, y0, rows_out, **warp_kw)
    return _as_tensor(out)


def displacement_reach(eng, disp_band, h_total: int, group=None, bounds=None) -> Tuple[int, int]:
    """(up, down): how many rows above / below ITS band any rank's warp_displacement taps reach, maximised over the
    ranks so that every rank sizes its halo identically (one scalar max-reduction).  It is a property of
    the FIELD: compute it once when the field changes (this call reads the result on the host) and pass it to every
    `warp_displacement_banded(..., reach=)` that uses the field."""
    rank, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    if world == 1:
        return (0, 0)
    y0, y1 = bounds[rank]
    d = _as_tensor(disp_band)
    if y1 > y0 and d.is_cuda and hasattr(eng, "disp_reach"):
        mm = eng.disp_reach(d, y0, h_total)  # [min, max] of floor(clamp(y - dy, -1, h)), on the device
        t = torch.stack([(y0 - mm[0]).clamp(min=0), (mm[1] + 2 - y1).clamp(min=0)]).to(torch.int32)
    elif y1 > y0:
        ys = torch.arange(y0, y1, dtype=torch.float32, device=d.device)[:, None]
        sy = ys - torch.nan_to_num(d[..., 1], nan=0.0, posinf=0.0, neginf=0.0)
        sy = sy.clamp(-1.0, float(h_total))
        t = torch.tensor([math.ceil(max(0.0, y0 - float(torch.floor(sy.min())))),
                          math.ceil(max(0.0, float(torch.floor(sy.max())) + 2 - y1))], dtype=torch.int32, device=d.device)
    else:
        t = torch.zeros(2, dtype=torch.int32, device=d.device)
    if t.device.type == "cpu" and dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    up, down = (int(v) for v in t.tolist())
    return up, down


def warp_displacement_banded(eng, band, disp_band, h_total: int, group=None, bounds=None, reach=None):
    """warp_displacement_full (transform.rs:1288-1345) on a row-split canvas (source and output share
    the split). `reach` = displacement_reach(...) of the field; computed here (one collective and one host read)
    when the caller does not pass it."""
    rank, world = _world(group)
    bounds = bounds or band_bounds(h_total, world)
    y0, y1 = bounds[rank]
    d = _as_tensor(disp_band)
    if reach is None:
        reach = displacement_reach(eng, d, h_total, group, bounds)
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
    return _warp_band(eng, band, h_total, w, y0, y1 - y0, (reach, reach), bounds, group,
                      original=original, deformed=deformed, cols=cols, rows=rows)
