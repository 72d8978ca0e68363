#!/usr/bin/env python
"""Condenses an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv
--log-file launches.csv python bench.py ...`) into the text kept under profiles/: one line per
launch (kernel names shortened) and the per-kernel totals with each kernel's share of the GPU time
of OUR kernels.  Per-launch times under ncu are serialised and cold-cache: the shares are what
must agree with bench.py, not the absolutes.

    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_launches.txt
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    if name.startswith("at::") or "at::" in name[:40]:
        m = re.search(r"(\w+_kernel)", name)
        return "torch:" + (m.group(1) if m else name[:40])
    return name.replace("(int)", "")


def main():
    path = sys.argv[1]
    cmd = sys.argv[2] if len(sys.argv) > 2 else ""
    rows = list(csv.reader(line for line in open(path) if line.startswith('"')))
    hdr, rows = rows[0], rows[1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    launches = []
    for r in rows:
        launches.append((short(r[ki]), r[gi], r[bi], float(r[vi].replace(",", "")) * scale[r[ui]]))
    print(f"# ncu launch list of: {cmd}")
    print(f"# {len(launches)} launches captured (gpu__time_duration.sum, --clock-control none)")
    agg = collections.OrderedDict()
    for name, grid, block, us in launches:
        a = agg.setdefault((name, grid), [0, 0.0, grid, block])
        a[0] += 1
        a[1] += us
    ours = sum(t for (k, _), (c, t, g, b) in agg.items() if not k.startswith("torch:"))
    print("\n== per kernel ==")
    print(f"{'launches':>8} {'mean us':>12} {'total us':>12} {'share of our GPU time':>22}  kernel  grid block")
    for (k, _), (c, t, g, b) in agg.items():
        share = "" if k.startswith("torch:") else f"{100 * t / ours:.1f} %"
        print(f"{c:8d} {t / c:12.1f} {t:12.1f} {share:>22}  {k}  {g} {b}")
    print("\n== per launch (torch fill kernels of the synthetic inputs omitted) ==")
    for i, (name, grid, block, us) in enumerate(launches):
        if not name.startswith("torch:"):
            print(f"{i:4d} {us:12.1f} us  {name}  {grid} {block}")


if __name__ == "__main__":
    main()
