#!/usr/bin/env python
"""Headline benchmark: Mpixels/s for a 16-layer 8K (7680x4320) flatten + Gaussian sigma=20.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One process per GPU (torchrun for N > 1; `python bench.py --gpus N` re-launches itself under
torch.distributed.run). Each rank owns its own 8K canvas (whole-image sharding, no data-path
collective: "weak" scaling).  A step = flatten(16 layers, all 25 modes cycled) followed
by Gaussian(sigma=20) with inputs resident in HBM.  Rank 0 prints ONE JSON line.

`--impl reference` times the CPU restatement of PaintFE's rayon path (oracle/, kind "port": the Rust
reference cannot be built in this image) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W8K, H8K, NLAYERS, SIGMA = 7680, 4320, 16, 20.0
METRIC = "Mpixels/sec: 16-layer 8K flatten + Gaussian sigma=20"
WORKLOAD = "8K (7680x4320) 16-layer synthetic stack cycling all 25 blend modes, flatten + Gaussian sigma=20"


def layer_meta():
    # mode = i mod 25, opacity = 0.25 + 0.05 i
    return [dict(blend=i % 25, opacity=0.25 + 0.05 * i) for i in range(NLAYERS)]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
                except ValueError:
                    continue
                for n, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ---------------------------------------------------------------------------------------------
def cpu_pass(pfo, layers_np, meta, w, h):
    t0 = time.perf_counter()
    flat = pfo.flatten([pfo.make_layer(im, **m) for im, m in zip(layers_np, meta)], w, h)
    pfo.gaussian_blur(flat, SIGMA)
    return time.perf_counter() - t0


def cpu_sample_layers(w, h):
    import numpy as np

    rng = np.random.default_rng(0x5EED)
    return [rng.integers(0, 256, (h, w, 4), dtype=np.uint8) for _ in range(NLAYERS)]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import pfo

    pfo.build()
    meta = layer_meta()
    # calibrate on 960x540, then pick the largest sample that keeps the whole run to ~2 minutes
    cal = cpu_sample_layers(960, 540)
    cpu_pass(pfo, cal, meta, 960, 540)
    t_cal = cpu_pass(pfo, cal, meta, 960, 540)
    per_px = t_cal / (960 * 540)
    total = args.steps + args.warmup
    w, h = 960, 540
    for cw, ch in ((7680, 4320), (3840, 2160), (1920, 1080)):
        if per_px * cw * ch * total <= 120.0:
            w, h = cw, ch
            break
    layers = cal if (w, h) == (960, 540) else cpu_sample_layers(w, h)
    for _ in range(args.warmup):
        cpu_pass(pfo, layers, meta, w, h)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(pfo, layers, meta, w, h)
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = (w * h) / dt / 1e6
    sample = f"{w}x{h} crop-sized canvas, {NLAYERS} layers, same modes/opacities/sigma; {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample,
                   "note": "CPU restatement of PaintFE's rayon path (OpenMP over chunks / rows); the Rust reference "
                           "cannot be built in this image (no cargo, ~400 crates)"},
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": pfo.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    _real_stdout = None
    if args.gpus > 1 and world == 1:
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the pixel engine has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its version banner straight to fd 1 at communicator creation (whatever NCCL_DEBUG says on
        # this image), and rank 0's stdout must be ONE JSON line: from here on everything any library writes to
        # stdout goes to stderr, and the result line is written to the saved descriptor at the end.
        sys.stdout.flush()
        _real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    from paintfe_b200.engine import Engine, make_layer

    eng = Engine(local)
    eng.use_torch_stream()
    w, h = W8K, H8K
    px = w * h
    meta = layer_meta()
    gen = torch.Generator(device=dev).manual_seed(0x5EED + rank)
    layers = [torch.randint(0, 256, (h, w, 4), dtype=torch.uint8, device=dev, generator=gen) for _ in range(NLAYERS)]
    dl = [make_layer(t, **m) for t, m in zip(layers, meta)]
    flat = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)
    out = torch.empty((h, w, 4), dtype=torch.uint8, device=dev)

    def step():
        eng.flatten(dl, w, h, out=flat)
        eng.gaussian_blur(flat, SIGMA, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    eng.profile(True)
    eng.profile_read()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launches - l0
    prof = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-pointer C ABI call (pinned host buffers) -----------------
    from paintfe_b200.dist import bind_to_gpu_numa

    numa = bind_to_gpu_numa(local)  # pinned buffers next to this GPU's root port (matters for N > 1)
    host_layers = [torch.empty((h, w, 4), dtype=torch.uint8, pin_memory=True) for _ in range(NLAYERS)]
    for hl, t in zip(host_layers, layers):
        hl.copy_(t)
    host_out = torch.empty((h, w, 4), dtype=torch.uint8, pin_memory=True)
    hl_np = [make_layer(t.numpy(), **m) for t, m in zip(host_layers, meta)]
    ho_np = host_out.numpy()
    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2):
        eng.flatten_gaussian(hl_np, w, h, SIGMA, out=ho_np)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(e2e_steps):
        eng.flatten_gaussian(hl_np, w, h, SIGMA, out=ho_np)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    # the host-tier result must equal the device-tier result (same inputs)
    same = bool(torch.equal(host_out.to(dev), out))

    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = world * px / (ms_step * 1e-3) / 1e6
    e2e_value = world * px / (e2e_ms / e2e_steps * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = measured_peak()
        # algorithmic bytes per launch: flatten 4L+4 per px; each Gaussian pass moves the
        # u8 image once and the f32 intermediate once (4 + 16 bytes per px)
        alg = {"flatten": (4 * NLAYERS + 4) * px, "gauss_h": 20 * px, "gauss_v": 20 * px}
        per_kernel = {}
        for k, v in prof.items():
            avg_ms = v["ms"] / max(v["launches"], 1)
            ent = {"launches": v["launches"], "avg_ms": avg_ms, "share_of_step": v["ms"] / ms_total}
            if k in alg and avg_ms > 0:
                ent["achieved_gbs"] = alg[k] / (avg_ms * 1e-3) / 1e9
                ent["frac_of_hbm"] = ent["achieved_gbs"] / peak
            per_kernel[k] = ent
        # the per-kernel event spans must add up to (just under) the step time measured around the whole loop
        span_share = sum(v["share_of_step"] for v in per_kernel.values())
        dom = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(dom)
        except Exception:
            pass
        roofline = None
        if dom in alg:
            ach = per_kernel[dom]["achieved_gbs"]
            roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": traffic, "peak_source": peak_src,
                        "note": "algorithmic bytes / CUDA-event kernel time; the kernel is FP32-issue bound when "
                                "bit-exact (DESIGN.md §4), so frac << 1 is expected"}
        # what actually bounds each kernel: warp-instruction issue for the flatten, FP32 FMA lanes
        # for the Gaussian passes. Static counts from the committed ncu capture, live kernel times.
        compute = {}
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            sm_clock_hz = (clocks or {}).get("sm_mhz", 1965.0) * 1e6
            issue_peak = 148 * 4 * sm_clock_hz  # one warp instruction per scheduler per clock
            for k, n_inst in tj.get("warp_instructions", {}).items():
                if k in per_kernel:
                    ach = n_inst / (per_kernel[k]["avg_ms"] * 1e-3)
                    compute[k] = {"bound": "warp-instruction issue", "achieved_ginst_s": ach / 1e9, "peak_ginst_s": issue_peak / 1e9,
                                  "frac": ach / issue_peak}
            for k, lanes in tj.get("fma_lanes", {}).items():
                if k in per_kernel:
                    ach = lanes / (per_kernel[k]["avg_ms"] * 1e-3)
                    compute[k].update({"fma_bound": "fp32 FMA lanes", "achieved_tfma_s": ach / 1e12,
                                       "peak_tfma_s": tj["fp32_fma_lanes_per_s_peak"] / 1e12, "fma_frac": ach / tj["fp32_fma_lanes_per_s_peak"]})
        except Exception:
            pass
        whole = {"achieved": 76 * px / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "bytes_per_px": 76}
        whole["frac"] = whole["achieved"] / peak

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import pfo

            pfo.build()
            cw, ch = 3840, 2160
            crop = [np.ascontiguousarray(t.numpy()[:ch, :cw]) for t in host_layers]
            cpu_pass(pfo, [c[:270, :480].copy() for c in crop], meta, 480, 270)
            best = min(cpu_pass(pfo, crop, meta, cw, ch) for _ in range(2))
            cpu = {"value": cw * ch / best / 1e6, "unit": "Mpixels/s", "cores": pfo.num_threads(), "kind": "port",
                   "sample": f"top-left {cw}x{ch} crop of the same 16 layers, flatten + Gaussian sigma=20, best of 2"}
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu": "one 8K canvas per rank, no collective",
                       "l2": "inputs 2.12 GB per step > 126 MB L2 (no flush needed)", "gaussian": "fast (FMA) path",
                       "timing": "CUDA events on the launching stream incl. per-kernel event pairs"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": NLAYERS * px * 4,
                    "d2h_bytes_per_step": px * 4, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "pfe_flatten_gaussian (host pointers, pinned)", "matches_device_tier": same,
                    "numa_rank0": numa},
            "gpu_launches": launches,
            "roofline": roofline, "roofline_whole_step": whole, "roofline_compute": compute, "kernels": per_kernel,
            "kernel_spans_share_of_step": span_share,
            "cpu_baseline": cpu,
        }
        if _real_stdout is not None:
            sys.stdout.flush()
            os.write(_real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
