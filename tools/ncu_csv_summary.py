#!/usr/bin/env python
"""Summarise an `ncu --csv --page raw` log (one row per profiled launch) into the table kept under profiles/.
  python tools/ncu_csv_summary.py gpurun_out/a_ops_ncu.csv [bench_ops --once stdout]  > profiles/rNN_ops_ncu.txt
"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fmacyc%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst")]


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if r]
    # skip ncu's ==PROF== preamble lines
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# one launch of every kernel (ncu --clock-control none, cold-cache serialised replays: read SHARES and counters, not absolute times)")
    print(f"{'kernel':58s} {'grid':>8s} {'blk':>4s} " + " ".join(f"{n:>9s}" for _, n in COLS) + "  top stalls (cycles per issued instruction)")
    for r in data:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = name.split("(")[0].replace("void ", "")
        vals = []
        for c, _ in COLS:
            if c in idx:
                v = r[idx[c]]
                u = units[idx[c]]
                try:
                    f = float(v.replace(",", ""))
                    if c.startswith("dram__bytes"):
                        f = f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1) / 1e6
                        vals.append(f"{f:8.1f}M")
                    elif c == "gpu__time_duration.sum":
                        f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                        vals.append(f"{f:8.1f}u")
                    elif c == "smsp__inst_executed.sum":
                        vals.append(f"{f / 1e6:8.1f}M")
                    else:
                        vals.append(f"{f:9.1f}")
                except ValueError:
                    vals.append(f"{v:>9s}")
            else:
                vals.append(f"{'-':>9s}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st = ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:4])
        print(f"{name[:58]:58s} {r[idx['launch__grid_size']]:>8s} {r[idx['launch__block_size']]:>4s} " + " ".join(vals) + "  " + st)


if __name__ == "__main__":
    main()
