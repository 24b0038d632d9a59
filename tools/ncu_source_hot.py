#!/usr/bin/env python
"""Where a kernel's warps wait, per SASS instruction, from an `ncu --set full --import-source on` report.
  python tools/ncu_source_hot.py gpurun_out/x.ncu-rep [kernel-substring] [top]
Prints, per kernel, the stall-sample totals by instruction class and the `top` instructions with the most samples."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            blocks.append(cur)
        elif cur is not None and row and row[0] == "Address":
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] and len(row) == len(cur["hdr"]):
            cur["rows"].append(row)
    for b in blocks:
        if want not in b["name"]:
            continue
        h = {k: i for i, k in enumerate(b["hdr"])}
        si, ni, xi = h["Warp Stall Sampling (All Samples)"], h["Warp Stall Sampling (Not-issued Samples)"], h["Instructions Executed"]
        stall_cols = [(k, i) for k, i in h.items() if k.startswith("stall_")]
        rows = b["rows"]
        total = sum(int(r[si]) for r in rows)
        print(f"== {b['name'][:110]}\n   stall samples {total}, instructions executed {sum(int(r[xi]) for r in rows)}")
        by = defaultdict(lambda: [0, 0, 0])
        for r in rows:
            op = r[h['Source']].split()[0] if not r[h['Source']].strip().startswith('@') else r[h['Source']].split()[1]
            op = op.split('.')[0]
            by[op][0] += int(r[si]); by[op][1] += int(r[ni]); by[op][2] += int(r[xi])
        print("   by opcode (samples, not-issued samples, executed):")
        for op, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:12]:
            print(f"     {op:10s} {v[0]:8d} {100.0 * v[0] / max(total, 1):5.1f}%  not-issued {v[1]:8d}  executed {v[2]}")
        if stall_cols:
            tot = {k: sum(int(r[i] or 0) for r in rows) for k, i in stall_cols}
            print("   by stall reason: " + ", ".join(f"{k[6:]}={v}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
        print(f"   top {top} instructions:")
        for idx, r in sorted(enumerate(rows), key=lambda ir: -int(ir[1][si]))[:top]:
            reasons = ""
            if stall_cols:
                rs = sorted(((int(r[i] or 0), k[6:]) for k, i in stall_cols), reverse=True)[:3]
                reasons = " ".join(f"{k}={v}" for v, k in rs if v)
            print(f"     #{idx:4d} {int(r[si]):7d} {r[h['Source']].strip()[:70]:70s} x{r[xi]:>8s}  {reasons}")


if __name__ == "__main__":
    main()
