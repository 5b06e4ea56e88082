#!/usr/bin/env python
"""Top warp-stall sampling hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv -k regex:<kernel>`.

    ncu -i gpurun_out/r01a_full.ncu-rep --page source --csv -k regex:fp_interp_mlp > /tmp/src.csv
    python tools/ncu_hotspots.py /tmp/src.csv [top_n]
"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    # several kernels may be concatenated: each starts with a "Kernel Name" row followed by a header row
    i = 0
    seen = None
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name, hdr = rows[i][1], rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) >= len(hdr) - 2:
                    body.append(rows[j])
                j += 1
            if (name, body) != seen:          # the CSV repeats each kernel (one block per source view): report it once
                report(name, hdr, body, top)
            seen = (name, body)
            i = j
        else:
            i += 1


def report(name, hdr, body, top):
    c = {h: k for k, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[c["# Samples"]] or 0) for r in body)
    print(f"== {name}: {len(body)} SASS instructions, {tot} samples")
    agg = {}
    for r in body:
        for h in stall_cols:
            agg[h] = agg.get(h, 0) + int(r[c[h]] or 0)
    print("   stall totals:", ", ".join(f"{h[6:]} {v * 100 // max(tot, 1)}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(body)), key=lambda k: -int(body[k][c["# Samples"]] or 0))[:top]
    for k in sorted(order):
        r = body[k]
        n = int(r[c["# Samples"]] or 0)
        st = sorted(((int(r[c[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print(f"   [{k:5d}] {n:6d} {n * 100.0 / max(tot, 1):5.1f}%  {r[c['Source']].strip()[:90]:90s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}  exec={r[c['Instructions Executed']]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
