#!/usr/bin/env python
"""Per-kernel and per-stage time + DRAM traffic from an
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.
The second half of the launches is the timed step (bench.py --steps 1 --warmup 1).
Writes a markdown table to stdout and, with --json, the per-stage figures bench.py reports as `traffic`."""
import argparse
import csv
import json
import re

STAGE = [("S1", r"k_tile|k_scan_runs|k_scan_u32|k_chain$|k_chain_from|k_crc|k_blockmap"),
         ("S2", r"k_inuse|k_codemap|k_kgram|k_seg_init|k_refine|k_resolve|k_rep_|k_power|k_bwt_out"),
         ("S3", r"k_mtf|k_rle2"),
         ("S4", r"k_huff|k_bit_offsets|k_group_scan|k_pack|k_pre_copy|k_put_bits|k_concat")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--mb", type=float, default=100.0, help="input bytes of the timed step, in MB")
    ap.add_argument("--json")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    lines = [l for l in open(a.csv) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"].startswith("gpu__time"):
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r["Metric Name"]] = v
    ids = sorted(launches)
    ids = ids[len(ids) // 2:]
    agg, stage = {}, {s: [0.0, 0.0, 0.0] for s, _ in STAGE}
    tot = 0.0
    for i in ids:
        d = launches[i]
        ms, rd, wr = d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
        x = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        x[0] += 1; x[1] += ms; x[2] += rd; x[3] += wr
        tot += ms
        for s, pat in STAGE:
            if re.match(pat, d["name"]):
                stage[s][0] += ms; stage[s][1] += rd; stage[s][2] += wr
                break
    n = a.mb * 1e6
    print(f"Timed step = last {len(ids)} launches; sum of launch durations {tot:.2f} ms.\n")
    print("| stage | ms | DRAM read GB | DRAM write GB | traffic B per input B |")
    print("|---|---:|---:|---:|---:|")
    for s, _ in STAGE:
        ms, rd, wr = stage[s]
        print(f"| {s} | {ms:.2f} | {rd/1e9:.2f} | {wr/1e9:.2f} | {(rd+wr)/n:.1f} |")
    print("\n| share | ms | launches | DRAM GB (r+w) | kernel |")
    print("|---:|---:|---:|---:|---|")
    for k, (c, ms, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {ms/tot*100:.2f}% | {ms:.3f} | {c} | {(rd+wr)/1e9:.2f} | `{k}` |")
    if a.json:
        out = {"workload": f"{a.mb:.0f} MB text window, -9",
               "per_input_byte": {s: {"dram_bytes": (stage[s][1] + stage[s][2]) / n, "ms_per_100MB": stage[s][0] * 100.0 / a.mb} for s, _ in STAGE},
               "source": a.source}
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
