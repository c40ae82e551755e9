#!/usr/bin/env python
"""One engine, one workload, input resident: `warmup` + `steps` passes of bz2b200_compress_device and nothing else.
The command ncu profiles (bench.py runs several engines, an end-to-end leg and a CPU leg around its timed region)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="text")
    ap.add_argument("--mb", type=int, default=100)
    ap.add_argument("--level", type=int, default=9)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    import torch
    import bzip2_b200 as B
    from bench import make_input
    n = a.mb * 1_000_000
    d = torch.from_numpy(make_input(a.workload, n)).cuda()
    cap = (n + n // 50 + 24576 * (n // (100000 * a.level - 19) + 2) + 1024 + 255) & ~255
    out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    eng = B.Engine(level=a.level)
    for _ in range(a.warmup + a.steps):
        m = eng.compress_device(d.data_ptr(), n, out.data_ptr(), cap)
    s = eng.stats
    print(f"{a.workload} {a.mb} MB -> {m} B; per step: total {s.ms_total:.2f} ms  S1 {s.ms_s1:.2f}  S2 {s.ms_s2:.2f}  S3 {s.ms_s3:.2f}  S4 {s.ms_s4:.2f}; "
          f"launches {s.kernel_launches}, rounds {s.bwt_rounds}")


if __name__ == "__main__":
    main()
