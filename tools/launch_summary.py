#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the
last `--last` fraction of launches (the timed step), optionally the launch-by-launch sequence."""
import argparse
import csv
import re


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--skip-frac", type=float, default=0.5, help="fraction of launches (warm-up) to skip")
    ap.add_argument("--seq", action="store_true")
    a = ap.parse_args()
    lines = [l for l in open(a.csv) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    rows = rows[int(len(rows) * a.skip_frac):]
    agg = {}
    tot = 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        v = float(r["Metric Value"].replace(",", ""))
        if a.seq:
            print(f"{v/1e3:9.1f} us  {name:34s} grid={r.get('Grid Size','')} block={r.get('Block Size','')}")
        x = agg.setdefault(name, [0, 0.0])
        x[0] += 1
        x[1] += v
        tot += v
    print(f"# {len(rows)} launches, sum {tot/1e6:.3f} ms")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v/tot*100:6.2f}%  {v/1e6:9.3f} ms  x{c:4d}  {k}")


if __name__ == "__main__":
    main()
