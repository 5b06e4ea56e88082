#!/usr/bin/env python
"""Summarises `ncu -i X.ncu-rep --page raw --csv` (one `--set full` capture of tools/ncu_once.py) into a table and a JSON
file under profiles/: per launch duration, DRAM bytes (read + write = `roofline.traffic`), DRAM %, tensor-pipe %, issue
utilisation, occupancy, registers, shared memory and the top warp-stall reasons.

    ncu -i gpurun_out/r01a_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/summarize_ncu.py /tmp/raw.csv [/tmp/newer_raw.csv ...] profiles/r01_ncu_summary [--drop REGEX]
"""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def short(name):
    m = re.match(r"(?:void )?(g4d::\w+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else re.sub(r"\(.*", "", name)[:80]


def load(raw_csv):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name, default=None):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return default
        try:
            return float(r[i].replace(",", "")) * UNIT.get(units[i].split("/")[0], 1.0)
        except ValueError:
            return default

    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = []
    for r in data:
        stalls = sorted(((val(r, h, 0.0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall_cols),
                        reverse=True)
        rd, wr = val(r, "dram__bytes_read.sum", 0.0), val(r, "dram__bytes_write.sum", 0.0)
        out.append({
            "kernel": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
            "duration_us": val(r, "gpu__time_duration.sum"),
            "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
            "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "l2_pct": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            "l1tex_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "sm_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "tensor_pipe_pct_active": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "tensor_pipe_pct_elapsed": val(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "regs": val(r, "launch__registers_per_thread"),
            "smem_dyn": val(r, "launch__shared_mem_per_block_dynamic"), "smem_static": val(r, "launch__shared_mem_per_block_static"),
            "top_stalls": [f"{n} {v:.2f}" for v, n in stalls[:3]],
        })
    return out


def main(raw_csvs, out_prefix, drop=None):
    """raw_csvs: one or more raw pages; kernels of a later file replace the same-named kernels of the earlier ones (a newer
    capture of a kernel that changed since); `drop`: regex of kernel names to leave out (no longer launched on the default path)."""
    out = []
    for path in raw_csvs:
        new = load(path)
        names = {k["kernel"] for k in new}
        out = [k for k in out if k["kernel"] not in names] + new
    if drop:
        out = [k for k in out if not re.search(drop, k["kernel"])]
    json.dump(out, open(out_prefix + ".json", "w"), indent=1)
    f = lambda v, fmt="%.1f": "-" if v is None else fmt % v
    with open(out_prefix + ".md", "w") as md:
        md.write("# ncu --set full, one launch of every hot-path kernel at c3 (240 clouds x 8192 points; tools/ncu_once.py)\n\n"
                 "Durations are under ncu replay (cold caches, no clock control): use them for shares and counters, not as bench values.\n\n"
                 "| # | kernel | grid | block | us | DRAM rd MB | DRAM wr MB | DRAM % | L2 % | SM % | tensor pipe % (active / elapsed) | issue % | warps % | regs | smem KB | top stalls (warps per issue) |\n"
                 "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for i, k in enumerate(out):
            md.write(f"| {i} | `{k['kernel']}` | {k['grid']} | {k['block']} | {f(k['duration_us'])} | {f(k['dram_read_bytes'] / 1e6, '%.2f')} | "
                     f"{f(k['dram_write_bytes'] / 1e6, '%.2f')} | {f(k['dram_pct'])} | {f(k['l2_pct'])} | {f(k['sm_pct'])} | "
                     f"{f(k['tensor_pipe_pct_active'])} / {f(k['tensor_pipe_pct_elapsed'])} | {f(k['issue_active_pct'])} | {f(k['warps_active_pct'])} | "
                     f"{f(k['regs'], '%d')} | {f(((k['smem_dyn'] or 0) + (k['smem_static'] or 0)) / 1024)} | {'; '.join(k['top_stalls'])} |\n")
    print(open(out_prefix + ".md").read())


if __name__ == "__main__":
    args = sys.argv[1:]
    drop = None
    if "--drop" in args:
        i = args.index("--drop")
        drop = args[i + 1]
        del args[i:i + 2]
    main(args[:-1], args[-1], drop)
