#!/usr/bin/env python
"""Condenses an `ncu --set full --csv --page raw` log into one line per profiled launch."""
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(suffix):
        c = [i for i, h in enumerate(hdr) if h.endswith(suffix)]
        return c[0] if c else None
    keys = [("t_us", "gpu__time_duration.sum", 1e-3), ("dram_rd_MB", "dram__bytes_read.sum", 1e-6), ("dram_wr_MB", "dram__bytes_write.sum", 1e-6),
            ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
            ("occ_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1), ("l2_hit", "lts__t_sector_hit_rate.pct", 1),
            ("inst_M", "smsp__inst_executed.sum", 1e-6), ("regs", "launch__registers_per_thread", 1),
            ("issue_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1)]
    # ncu scales the unit per column (ns/us/ms, byte/Kbyte/...): bring everything back to ns and bytes
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    stall = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), i)
             for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    ci = [(k, col(s), f) for k, s, f in keys]
    print("| kernel | grid | " + " | ".join(k for k, _, _ in ci) + " | top stalls (warps per issue) |")
    print("|---|---|" + "---:|" * len(ci) + "---|")
    for r in data:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
        vals = []
        for k, i, f in ci:
            try:
                vals.append(f"{float(r[i].replace(',', '')) * scale.get(units[i], 1.0) * f:.1f}")
            except Exception:  # noqa: BLE001
                vals.append("-")
        st = sorted(((float(r[i].replace(",", "")), n) for n, i in stall), reverse=True)[:3]
        print(f"| `{name}` | {r[8]} | " + " | ".join(vals) + " | " + ", ".join(f"{n} {v:.1f}" for v, n in st) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
