#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list (gpu__time_duration + a few SM metrics) per kernel name."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        per[(r[iid], r[ik].split("<")[0].split("(")[0].replace("void ", ""))][r[im]] = float(r[iv].replace(",", ""))
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for (_, k), m in per.items():
        for name, v in m.items():
            agg[k][name].append(v)
    tot = sum(sum(v["gpu__time_duration.sum"]) for v in agg.values())
    print(f"{'kernel':24s} {'n':>3s} {'total ms':>9s} {'share':>6s} {'mean us':>8s} {'no_instr':>8s} {'issue%':>7s} {'Minst':>8s} {'warps%':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        t = v["gpu__time_duration.sum"]
        g = lambda name: (sum(v[name]) / len(v[name])) if name in v else float("nan")
        print(f"{k:24s} {len(t):3d} {sum(t) / 1e6:9.3f} {100 * sum(t) / tot:5.1f}% {sum(t) / len(t) / 1e3:8.1f} "
              f"{g('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'):8.2f} "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} {g('smsp__inst_executed.sum') / 1e6:8.1f} "
              f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f}")
    print(f"sum {tot / 1e6:.3f} ms over {sum(len(v['gpu__time_duration.sum']) for v in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
