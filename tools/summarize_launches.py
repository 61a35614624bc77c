#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import collections
import csv
import sys


def main(path, title=""):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(r[ki], []).append((float(r[vi].replace(",", "")), r[gi], r[bi]))
    tot = sum(v[0] for vs in agg.values() for v in vs)
    print(f"# {title or path}\n")
    print(f"{len(rows) - 1} launches, {tot / 1e6:.3f} ms total (ncu-serialised, cold cache: compare shares, not absolutes)\n")
    print("| kernel | launches | total ms | mean us | share | grid / block (last) |")
    print("|---|---:|---:|---:|---:|---|")
    for k, vs in sorted(agg.items(), key=lambda kv: -sum(v[0] for v in kv[1])):
        t = sum(v[0] for v in vs)
        name = k if len(k) < 90 else k[:87] + "..."
        print(f"| `{name}` | {len(vs)} | {t / 1e6:.3f} | {t / len(vs) / 1e3:.1f} | {100 * t / tot:.1f}% | {vs[-1][1]} / {vs[-1][2]} |")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
