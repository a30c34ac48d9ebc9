#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

usage: python tools/launch_summary.py gpurun_out/launches.csv [--only cfb] > profiles/<name>.txt
Durations under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    for row in csv.DictReader(lines):
        name = row.get("Kernel Name", "")
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row.get("Metric Unit", "ns")
        v = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        if only and only not in name:
            continue
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"void |cfb::<unnamed>::|cfb::\(anonymous namespace\)::", "", short)[:64]
        n, t = agg.get(short, (0, 0.0))
        agg[short] = (n + 1, t + v)
        total += v
    print(f"# {path}: {sum(n for n, _ in agg.values())} launches, {total:.1f} us summed device time (serialised, cold)")
    print(f"{'share':>7} {'total_us':>10} {'count':>6} {'avg_us':>8}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * t / total:6.1f}% {t:10.1f} {n:6d} {t / n:8.2f}  {k}")


if __name__ == "__main__":
    main()
