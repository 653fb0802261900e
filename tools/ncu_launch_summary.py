#!/usr/bin/env python
"""Condenses an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X.csv) into the per-kernel table kept
under profiles/.  Usage: python tools/ncu_launch_summary.py gpurun_out/r01_launches.csv [skip_launches] > profiles/NAME.md"""
import csv
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    seen = 0
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        seen += 1
        if seen <= skip:
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("fpc::", "").replace("<unnamed>::", "")
        unit, val = r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
        us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + us)
    total = sum(t for _, t in agg.values())
    print(f"# ncu launch list summary (`--metrics gpu__time_duration.sum --clock-control none`) of `{path}`"
          f"{'' if not skip else f', first {skip} launches (warm-up) skipped'}\n")
    print("cold-cache, serialised launches: compare SHARES with bench.py's kernel_ms, not absolutes\n")
    print("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for name, (n, t) in agg.items():
        print(f"| {name} | {n} | {t / n:.1f} | {100 * t / total:.1f}% |")
    print(f"\ntotal {total:.1f} us over {sum(n for n, _ in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
