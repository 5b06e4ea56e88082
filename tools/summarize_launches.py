#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py` into per-kernel totals for ONE
timed step (the launches between two L2-flush fills of the device-resident arm).  Per-launch times under ncu are
cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md).

    python tools/summarize_launches.py gpurun_out/r01a_launches.csv > profiles/r01_launches_summary.md
"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    m = re.match(r"(g4d::\w+)(<[^>]*>)?", name)
    if m:
        return m.group(1) + (m.group(2) or "")
    name = re.sub(r"\(.*", "", name)
    return name[:90]


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    for r in rows:
        r["us"] = float(r["Metric Value"].replace(",", "")) * to_us[r["Metric Unit"]]
    flush = [i for i, r in enumerate(rows) if "FillFunctor<float>" in r["Kernel Name"] and r["us"] > 20.0]
    if len(flush) < 2:
        raise SystemExit("no timed steps (L2-flush fills) found in the launch list")
    lo, hi = flush[0] + 1, flush[1]          # first timed step of the device-resident arm
    step = rows[lo:hi]
    agg = collections.OrderedDict()
    for r in step:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Block Size"], r["Grid Size"]])
        a[0] += 1
        a[1] += r["us"]
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if k.startswith("g4d::"))
    print(f"# ncu launch list, one timed step of `bench.py` (c3: 240 frames, kernel by kernel without the CUDA graph; frame groups as given on the command line): {len(step)} launches, "
          f"{total:.0f} us serialised, {ours / total * 100:.1f} % in libgarment4d_b200 kernels "
          f"({sum(a[0] for k, a in agg.items() if k.startswith('g4d::'))} launches)\n")
    print("| kernel | launches | total us | share | block | grid (first launch) |")
    print("|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / total * 100:.1f} % | {a[2]} | {a[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])
