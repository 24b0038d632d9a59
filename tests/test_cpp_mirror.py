"""The C++ host-side mirror of PaintFE's API (include/paintfe/paintfe.hpp): it must compile and link
against the C ABI on any box, and on a GPU box the ports of the reference's own tests must pass."""
import os
import struct
import subprocess

import numpy as np
import pytest

import fixtures as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mirror_binary(tmp_path_factory):
    from paintfe_b200 import build

    build.build()
    out = str(tmp_path_factory.mktemp("cpp") / "mirror_tests")
    libdir = os.path.join(ROOT, "paintfe_b200")
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror",
           "-o", out, os.path.join(ROOT, "tests", "cpp", "mirror_tests.cpp"), "-L" + libdir, "-lpfe_b200",
           "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_mirror_compiles_and_links(mirror_binary):
    assert os.path.exists(mirror_binary)
    r = subprocess.run([mirror_binary], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stdout  # no golden dir given


@pytest.mark.gpu
def test_reference_tests_ported_to_cpp_mirror(mirror_binary, tmp_path):
    for cat in sorted(os.listdir(fx.GOLDEN_DIR)):
        for f in sorted(os.listdir(os.path.join(fx.GOLDEN_DIR, cat))):
            name = f[:-4]
            img = np.ascontiguousarray(fx.golden(cat, name))
            with open(tmp_path / f"{cat}__{name}.rgba", "wb") as fh:
                fh.write(struct.pack("<II", img.shape[1], img.shape[0]))
                fh.write(img.tobytes())
    r = subprocess.run([mirror_binary, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
