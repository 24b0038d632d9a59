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


def run_one(eng, inp: str, outp: str, script: Optional[str], verbose: bool, flatten: bool = True, exact: bool = False) -> None:
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
        img = np.ascontiguousarray(np.asarray(Image.open(inp).convert("RGBA")))
        if script:
            img = run_script(img)[0]
    os.makedirs(os.path.dirname(os.path.abspath(outp)), exist_ok=True)
    Image.fromarray(np.asarray(img), "RGBA").save(outp)


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
    args = ap.parse_args(argv)

    inputs = resolve_inputs(args.input)
    if not inputs:
        print("error: no input files.", file=sys.stderr)
        return 1
    if args.output and len(inputs) > 1:
        print("error: --output is only valid for a single input; use --output-dir.", file=sys.stderr)
        return 1
    fmt = (args.format or (os.path.splitext(args.output)[1][1:] if args.output else "") or "png").lower()
    script = open(args.script).read() if args.script else None

    from .dist import shard_indices
    from .engine import Engine

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    eng = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    total, any_failure = len(inputs), False
    for idx in shard_indices(total, rank, world):
        path = inputs[idx]
        if total > 1 or args.verbose:
            print(f"[{idx + 1}/{total}] {path}")
        t0 = time.perf_counter()
        outp = build_output_path(path, args.output, args.output_dir, fmt)
        if outp is None:
            print(f"  error: cannot determine output path for '{path}'.", file=sys.stderr)
            any_failure = True
            continue
        try:
            run_one(eng, path, outp, script, args.verbose, args.flatten, args.exact)
            if args.verbose or total > 1:
                print(f"  -> {outp} ({(time.perf_counter() - t0) * 1000:.0f}ms)")
        except Exception as e:  # per-file failure: report and continue (cli.rs:204-209)
            print(f"  error: {e}", file=sys.stderr)
            any_failure = True
    eng.close()
    return 1 if any_failure else 0


if __name__ == "__main__":
    sys.exit(main())
