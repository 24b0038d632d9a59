"""CLI plumbing that needs no GPU: output formats (src/cli.rs:277-304) and argument errors."""
import os

import numpy as np
import pytest
from PIL import Image

from paintfe_b200 import cli


@pytest.mark.parametrize("fmt,mode", [("png", "RGBA"), ("jpg", "RGB"), ("jpeg", "RGB"), ("bmp", "RGB"), ("tiff", "RGBA"), ("webp", "RGBA")])
def test_encode_and_write_formats(tmp_path, fmt, mode):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (24, 40, 4), dtype=np.uint8)
    out = tmp_path / "sub" / f"x.{fmt}"
    cli.encode_and_write(img, str(out), fmt)
    back = Image.open(out)
    assert back.size == (40, 24) and back.mode == mode
    if fmt in ("png", "tiff"):
        assert np.array_equal(np.asarray(back), img)
    if fmt == "bmp":
        assert np.array_equal(np.asarray(back), img[..., :3])  # alpha dropped, colours untouched


def test_unknown_format_is_rejected_before_any_work(tmp_path, capsys):
    src = tmp_path / "a.png"
    Image.fromarray(np.zeros((4, 4, 4), np.uint8), "RGBA").save(src)
    assert cli.main(["-i", str(src), "--output-dir", str(tmp_path / "out"), "-f", "xyz"]) == 1
    assert "unknown output format" in capsys.readouterr().err
    assert not os.path.exists(tmp_path / "out")


def test_output_needs_single_input(tmp_path, capsys):
    for n in "ab":
        Image.fromarray(np.zeros((4, 4, 4), np.uint8), "RGBA").save(tmp_path / f"{n}.png")
    assert cli.main(["-i", str(tmp_path / "*.png"), "-o", str(tmp_path / "o.png")]) == 1
    assert "--output is only valid for a single input" in capsys.readouterr().err
