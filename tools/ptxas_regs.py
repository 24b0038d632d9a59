#!/usr/bin/env python
"""Registers / spills / shared memory per kernel of one translation unit, from `ptxas -v`.

  python tools/ptxas_regs.py gaussian.cu [filter-substring ...]

Compiles paintfe_b200/csrc/<file> with the library's own flags (nothing is written into the build tree) and
prints one line per entry function whose demangled name contains every filter substring.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from paintfe_b200 import build as B  # noqa: E402


def main():
    src = sys.argv[1]
    filters = sys.argv[2:]
    path = src if os.path.exists(src) else os.path.join(B.CSRC, src)
    with tempfile.TemporaryDirectory() as td:
        cmd = [B.nvcc(), "-ccbin", "/usr/bin/g++"] + B.NVCC_FLAGS + ["-Xptxas", "-v", "-c", path, "-o", os.path.join(td, "o.o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr)
        return 1
    name = None
    spill = ""
    rows = []
    for line in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(.+)' for", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = f"stack {m.group(1)} spill {m.group(2)}/{m.group(3)}"
            continue
        m = re.search(r"Used (\d+) registers, used (\d+) barriers(.*)", line)
        if m and name:
            rows.append((name, int(m.group(1)), spill, m.group(3).strip(", ")))
            name = None
    try:
        dem = subprocess.run(["c++filt"], input="\n".join(n for n, *_ in rows), capture_output=True, text=True).stdout.splitlines()
    except Exception:
        dem = [n for n, *_ in rows]
    for (n, regs, sp, rest), d in zip(rows, dem):
        d = d.replace("(anonymous namespace)::", "")
        d = re.sub(r"\(.*", "", d)
        if all(f in d for f in filters):
            print(f"{regs:4d} regs  {sp:28s} {d}   [{rest}]")
    return 0


if __name__ == "__main__":
    sys.exit(main())
