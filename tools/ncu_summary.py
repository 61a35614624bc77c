#!/usr/bin/env python
"""Summarise an ncu --set full report (.ncu-rep) per captured launch: the metrics B200_PROFILING.md names."""
import csv, subprocess, sys, io

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]


def main(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {title}\n\nsource: `{path}` (ncu --set full --clock-control none; values of a profiled, serialised launch)\n")
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        print(f"## {d['Kernel Name'][1][:100]}  grid {d['Grid Size'][1]} block {d['Block Size'][1]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k][1]} | {d[k][0]} |")
        # warp-state sampling (pcsamp): share of the samples per stall reason
        samp = {h.split("issue_stalled_")[1]: float((v[1] or "0").replace(",", "")) for h, v in d.items()
                if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
        tot = sum(samp.values()) or 1.0
        top = sorted(samp.items(), key=lambda kv: -kv[1])[:6]
        print("\ntop warp-stall reasons (% of sampled warp states): " +
              ", ".join(f"{k} {100 * v / tot:.0f}" for k, v in top) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
