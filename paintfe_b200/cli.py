"""Headless batch mode sharded over GPUs: the `cli::run` loop (src/cli.rs:105-215) with the serial
`for input in inputs` (cli.rs:159) replaced by "image k -> rank k mod world".

  python -m paintfe_b200.cli -i "shots/*.png" --script process.rhai --output-dir out/
  torchrun --nproc-per-node 8 -m paintfe_b200.cli -i "shots/*.png" --script process.rhai --output-dir out/

Same flags as the reference for the part of the pipeline that is in scope: -i/--input (glob patterns
or literal paths, deduplicated in order, cli.rs:315-350), -s/--script, -o/--output (single input
only), --output-dir, -f/--format, -v/--verbose.  Image decoding/encoding is harness plumbing (PIL);
`.pfe` v0/v1/v2/v3 projects load through paintfe_b200/pfe_io.py and are flattened on the GPU (--flatten).  A per-file failure is reported and the batch
continues; the exit code is 1 if any file failed (cli.rs:204-215).
"""
from __future__ import annotations

import argparse
import glob
import os
import sys
import time
from typing import List, Optional

import numpy as np


def resolve_inputs(patterns: List[str]) -> List[str]:
    """cli.rs:315-350: literal paths first, else glob; ordered, deduplicated; warn on empty matches."""
    out: List[str] = []
    for pat in patterns:
        if os.path.exists(pat):
            if pat not in out:
                out.append(pat)
            continue
        matched = False
        for p in sorted(glob.glob(pat)):
            if p not in out:
                out.append(p)
            matched = True
        if not matched:
            print(f"warning: pattern '{pat}' matched no files.", file=sys.stderr)
    return out


def build_output_path(inp: str, output: Optional[str], output_dir: Optional[str], fmt: str) -> Optional[str]:
    """cli.rs build_output_path: --output for a single file, else <output_dir>/<stem>.<ext>."""
    if output:
        return output
    stem = os.path.splitext(os.path.basename(inp))[0]
    if output_dir:
        return os.path.join(output_dir, f"{stem}.{fmt}")
    return None


RASTER_FORMATS = {"png": "PNG", "jpg": "JPEG", "jpeg": "JPEG", "bmp": "BMP", "webp": "WEBP", "tif": "TIFF", "tiff": "TIFF",
                  "tga": "TGA", "gif": "GIF", "ico": "ICO"}


def encode_and_write(img: np.ndarray, outp: str, fmt: str) -> None:
    """cli.rs encode_and_write (:277-304) for the formats PIL writes: JPEG and BMP carry no alpha, so the flat image is
    composited onto nothing and stored as RGB the way `image` does it (alpha dropped)."""
    from PIL import Image

    os.makedirs(os.path.dirname(os.path.abspath(outp)), exist_ok=True)
    pil = Image.fromarray(np.asarray(img), "RGBA")
    kind = RASTER_FORMATS[fmt]
    if kind in ("JPEG", "BMP"):
        pil = pil.convert("RGB")
    pil.save(outp, format=kind)


def run_one(eng, inp: str, outp: str, script: Optional[str], verbose: bool, flatten: bool = True, exact: bool = False,
            fmt: str = "png") -> None:
    """cli.rs:222-308: load -> script on the active layer -> flatten if several layers -> encode."""
    from PIL import Image

    from .engine import make_layer
    from .rhai_host import Interpreter, ScriptError
    from .script import apply_canvas_ops

    def run_script(pixels):  # cli.rs:247-254
        it = Interpreter(eng, pixels, exact=exact)
        try:
            res = it.run(script)
        except ScriptError as e:
            raise RuntimeError(f"script error: {e}") from None
        if verbose:
            for line in it.console:
                print(f"  [script] {line}")
        return res, it.canvas_ops

    if inp.lower().endswith(".pfe"):  # io::load_image_sync -> load_pfe (io.rs:693, :469)
        from . import pfe_io

        proj = pfe_io.load_pfe(inp)
        w, h = proj.width, proj.height
        ai = min(proj.active_layer_index, len(proj.layers) - 1)
        cxn, cyn = (w + 63) // 64, (h + 63) // 64

        def table(L):  # the project's sparse chunks as a TiledImage chunk table (row-major, None = unpopulated)
            t = [None] * (cxn * cyn)
            for (cx, cy), px in L.chunks.items():
                if cx < cxn and cy < cyn:
                    t[cy * cxn + cx] = px
            return t

        tables = [table(L) for L in proj.layers]
        if script:  # the script sees the active layer as a flat image and commits it back as tiles (cli.rs:239-260)
            res, canvas_ops = run_script(proj.layers[ai].to_flat(w, h))
            flats = [np.asarray(res) if i == ai else None for i in range(len(tables))]
            if canvas_ops:  # canvas-wide calls (resize, quarter turns, ...) are replayed on the other layers (cli.rs:262-266)
                flats = [f if i == ai else (None if L.adjustment is not None else L.to_flat(w, h))
                         for i, (f, L) in enumerate(zip(flats, proj.layers))]
                flats = [None if f is None else np.asarray(f) for f in apply_canvas_ops(eng, flats, ai, canvas_ops)]
            h, w = flats[ai].shape[:2]
            for i, f in enumerate(flats):
                if f is not None:
                    occ, tiles = eng.flat_to_tiles(f)
                    tables[i] = [tiles[k] if occ.reshape(-1)[k] else None for k in range(occ.size)]
        if flatten and len(tables) > 1:  # cli.rs:282-285 state.composite(): straight from the chunk tables
            descs = []
            for i, (t, L) in enumerate(zip(tables, proj.layers)):
                vis = proj.layer_effectively_visible(i)  # a hidden folder hides its layers (canvas_state.rs:216)
                if L.adjustment is not None:             # v3 adjustment layer (layers.rs:276-325)
                    kind, prm = L.adjustment
                    adj = (float(np.float32(2.0) ** np.float32(prm[0])),) if kind == 1 else prm  # Exposure: 2^ev on the host
                    descs.append(dict(opacity=L.opacity, blend=L.blend_mode, visible=vis, kind=kind, adj=adj))
                else:
                    descs.append(dict(tiles=t, opacity=L.opacity, blend=L.blend_mode, visible=vis))
            img = eng.flatten_tiles(descs, w, h)
        else:
            img = eng.tiles_to_flat(tables[ai], w, h)
    else:
        img = np.array(Image.open(inp).convert("RGBA"))  # a writable copy (PIL hands out read-only views)
        if script:
            # one upload, the whole script on the device (every apply_* is a pfe_dev_* call on the resident image),
            # one download - not a PCIe round trip per effect call
            import torch

            res = run_script(torch.from_numpy(img).to(f"cuda:{eng.device}"))[0]
            img = res.cpu().numpy() if isinstance(res, torch.Tensor) else np.asarray(res)
    if fmt == "pfe":  # SaveFormat::Pfe (cli.rs:277): the project itself, not a flattened raster
        from . import pfe_io

        if inp.lower().endswith(".pfe"):
            cxn = (w + 63) // 64  # the script may have resized the canvas
            for i, t in enumerate(tables):
                proj.layers[i].chunks = {(k % cxn, k // cxn): px for k, px in enumerate(t) if px is not None}
            proj.width, proj.height = w, h
        else:
            proj = pfe_io.PfeProject(img.shape[1], img.shape[0], 0, [pfe_io.layer_from_flat("Background", img)])
        os.makedirs(os.path.dirname(os.path.abspath(outp)), exist_ok=True)
        with open(outp, "wb") as f:
            f.write(pfe_io.save_pfe_v3(proj) if (proj.folders or any(L.adjustment is not None for L in proj.layers)) else pfe_io.save_pfe_v1(proj))
        return
    encode_and_write(img, outp, fmt)


def run_batch_pipelined(eng, jobs, script: str, verbose: bool, exact: bool, fmt: str, depth: int = 3, workers: int = 4):
    """Raster inputs through one script: decode (thread pool) -> ImagePipeline (upload / script / download on three
    streams, `depth` images in flight) -> encode (thread pool).  `jobs` = [(index, input path, output path)].
    Returns the indices that failed.  Replaces the body of the serial loop at src/cli.rs:159-209 for this case."""
    from concurrent.futures import ThreadPoolExecutor

    from PIL import Image

    from .pipeline import ImagePipeline
    from .rhai_host import Interpreter, ScriptError

    failed, writes = [], []
    pipe = ImagePipeline(eng, depth)

    def decode(path):
        return np.array(Image.open(path).convert("RGBA"))

    def work(dev):
        it = Interpreter(eng, dev, exact=exact)
        try:
            res = it.run(script)
        except ScriptError as e:
            raise RuntimeError(f"script error: {e}") from None
        if verbose:
            for line in it.console:
                print(f"  [script] {line}")
        if not hasattr(res, "is_cuda"):  # a script that ended in host pixel access: back to the device for the download
            import torch
            res = torch.from_numpy(np.asarray(res)).to(dev.device)
        return res

    def finish(pool):
        (idx, outp), out = pipe.collect()
        writes.append((idx, pool.submit(encode_and_write, np.array(out), outp, fmt)))  # copy: the slot is reused

    with ThreadPoolExecutor(max_workers=workers) as pool:
        ahead = 2 * depth
        decoded = {}
        for k, (idx, path, outp) in enumerate(jobs):
            for j in range(k, min(k + ahead, len(jobs))):
                if j not in decoded:
                    decoded[j] = pool.submit(decode, jobs[j][1])
            try:
                img = decoded.pop(k).result()
                if pipe.full():
                    finish(pool)
                pipe.submit(img, work, tag=(idx, outp))
            except Exception as e:  # per-file failure: report and continue (cli.rs:204-209)
                print(f"  error: {path}: {e}", file=sys.stderr)
                failed.append(idx)
        while pipe.pending:
            finish(pool)
        for idx, fut in writes:
            try:
                fut.result()
            except Exception as e:
                print(f"  error: {jobs_path(jobs, idx)}: {e}", file=sys.stderr)
                failed.append(idx)
    return failed


def jobs_path(jobs, idx):
    return next((p for i, p, _ in jobs if i == idx), "?")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="paintfe", description="PaintFE headless batch image processor (B200 engine)")
    ap.add_argument("-i", "--input", nargs="+", required=True)
    ap.add_argument("-s", "--script")
    ap.add_argument("-o", "--output")
    ap.add_argument("--output-dir")
    ap.add_argument("-f", "--format", default=None)
    ap.add_argument("--flatten", action=argparse.BooleanOptionalAction, default=True)
    ap.add_argument("--exact", action="store_true", help="bit-exact Gaussian/sharpen (reference tap order, no FMA)")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--depth", type=int, default=3, help="images in flight in the batch pipeline (upload / script / download)")
    ap.add_argument("--no-pipeline", action="store_true", help="process every file one at a time")
    args = ap.parse_args(argv)

    inputs = resolve_inputs(args.input)
    if not inputs:
        print("error: no input files.", file=sys.stderr)
        return 1
    if args.output and len(inputs) > 1:
        print("error: --output is only valid for a single input; use --output-dir.", file=sys.stderr)
        return 1
    fmt = (args.format or (os.path.splitext(args.output)[1][1:] if args.output else "") or "png").lower()
    if fmt != "pfe" and fmt not in RASTER_FORMATS:
        print(f"error: unknown output format '{fmt}' (one of: pfe, {', '.join(sorted(RASTER_FORMATS))}).", file=sys.stderr)
        return 1
    script = open(args.script).read() if args.script else None

    from .dist import shard_indices
    from .engine import Engine

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    eng = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    total, any_failure = len(inputs), False
    mine = shard_indices(total, rank, world)
    # raster inputs with a script and a raster output: the pipelined batch path; .pfe projects, script-less conversions
    # and .pfe output go one file at a time
    jobs, singles = [], []
    for idx in mine:
        outp = build_output_path(inputs[idx], args.output, args.output_dir, fmt)
        if outp is None:
            print(f"  error: cannot determine output path for '{inputs[idx]}'.", file=sys.stderr)
            any_failure = True
        elif script and fmt != "pfe" and not inputs[idx].lower().endswith(".pfe") and len(mine) > 1 and not args.no_pipeline:
            jobs.append((idx, inputs[idx], outp))
        else:
            singles.append((idx, inputs[idx], outp))
    if jobs:
        t0 = time.perf_counter()
        failed = run_batch_pipelined(eng, jobs, script, args.verbose, args.exact, fmt, depth=args.depth)
        any_failure = any_failure or bool(failed)
        if args.verbose or total > 1:
            dt = time.perf_counter() - t0
            print(f"[rank {rank}] {len(jobs) - len(failed)} of {len(jobs)} images through the pipelined path in {dt * 1000:.0f}ms "
                  f"({len(jobs) / max(dt, 1e-9):.1f} images/s)")
    for idx, path, outp in singles:
        if total > 1 or args.verbose:
            print(f"[{idx + 1}/{total}] {path}")
        t0 = time.perf_counter()
        try:
            run_one(eng, path, outp, script, args.verbose, args.flatten, args.exact, fmt)
            if args.verbose or total > 1:
                print(f"  -> {outp} ({(time.perf_counter() - t0) * 1000:.0f}ms)")
        except Exception as e:  # per-file failure: report and continue (cli.rs:204-209)
            print(f"  error: {e}", file=sys.stderr)
            any_failure = True
    eng.close()
    return 1 if any_failure else 0


if __name__ == "__main__":
    sys.exit(main())
