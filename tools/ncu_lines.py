#!/usr/bin/env python
"""tools/ncu_lines.py REPORT.ncu-rep [--top N] [--launch I] [--ranges file:lo-hi=name ...]
Per-source-line attribution of an `ncu --set full --import-source on` capture (needs -lineinfo): warp
instructions executed and stall samples per (file, line), from `ncu --page source --print-source cuda,sass`.
Prints the top lines and, with --ranges, the totals of named line ranges (phases of a kernel)."""
import argparse
import csv
import io
import subprocess
from collections import defaultdict


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    # layout: ("File Path", f) ("Function Name", k) header rows...; a row whose first field is a number is a source
    # line with the metrics summed over its SASS instructions, rows with an empty first field are the instructions
    kernels, cur, fpath, header = {}, None, None, None
    for row in csv.reader(io.StringIO(txt)):
        if not row:
            continue
        if row[0] == "File Path":
            fpath = row[1]
        elif row[0] in ("Function Name", "Kernel Name"):
            cur = kernels.setdefault(row[1], {"name": row[1], "lines": []})
        elif row[0] == "Line No":
            header = row
        elif header and len(row) == len(header) and row[0].isdigit() and cur is not None:
            d = {h: v for h, v in zip(header[2:], row[2:])}
            cur["lines"].append((fpath, int(row[0]), row[1], d))
    kernels = list(kernels.values())
    return kernels


def num(d, k):
    try:
        return float(d.get(k, "0").replace(",", ""))
    except ValueError:
        return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--ranges", nargs="*", default=[])
    a = ap.parse_args()
    ks = load(a.rep)
    if not ks:
        raise SystemExit("no kernels with source found")
    k = ks[min(a.launch, len(ks) - 1)]
    per = defaultdict(lambda: [0.0, 0.0, ""])
    for f, ln, src, d in k["lines"]:
        key = (f.split("/")[-1], ln)
        per[key][0] += num(d, "Instructions Executed")
        per[key][1] += num(d, "Warp Stall Sampling (All Samples)")
        per[key][2] = src
    tot_i = sum(v[0] for v in per.values()) or 1.0
    tot_s = sum(v[1] for v in per.values()) or 1.0
    print(f"kernel: {k['name'][:100]}\nwarp instructions {tot_i:.0f}, stall samples {tot_s:.0f}")
    print(f"{'file:line':28s} {'inst%':>6s} {'smpl%':>6s}  source")
    for key, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print(f"{key[0] + ':' + str(key[1]):28s} {100 * v[0] / tot_i:6.2f} {100 * v[1] / tot_s:6.2f}  {v[2].strip()[:90]}")
    if a.ranges:
        print("\nranges:")
        for spec in a.ranges:
            loc, name = spec.split("=")
            f, r = loc.split(":")
            lo, hi = (int(x) for x in r.split("-"))
            i = sum(v[0] for kk, v in per.items() if kk[0] == f and lo <= kk[1] <= hi)
            s = sum(v[1] for kk, v in per.items() if kk[0] == f and lo <= kk[1] <= hi)
            print(f"  {name:34s} {loc:28s} inst {100 * i / tot_i:6.2f} %  samples {100 * s / tot_s:6.2f} %")


if __name__ == "__main__":
    main()
