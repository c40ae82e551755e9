#!/usr/bin/env python
"""Writes one SASS listing per kernel under profiles/sass/ (cuobjdump -sass of the built library)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bzip2_b200", "libbz2_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")


def main():
    os.makedirs(OUT, exist_ok=True)
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    rows = []
    for p in parts[1:]:
        name = p.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", dem).replace("void ", "").replace("bz::", "")
        short = short.replace("<", "_").replace(">", "").replace(", ", "_").replace(" ", "")
        # keep the mnemonic column only: encodings double the size and add nothing for review
        lines = []
        for ln in p.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*/\*\s*0x[0-9a-f]+\s*\*/", ln)
            if m:
                lines.append(f"/*{m.group(1)}*/ {m.group(2)}")
        open(os.path.join(OUT, short + ".sass"), "w").write(f"// {dem}\n" + "\n".join(lines) + "\n")
        ops = sorted(set(re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l.split("*/ ", 1)[1]).group(1) for l in lines
                         if re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l.split("*/ ", 1)[1])))
        notable = [o for o in ops if o in ("MATCH", "VOTE", "SHFL", "ATOMS", "ATOMG", "RED", "REDUX", "BAR", "LDS", "STS",
                                           "POPC", "FLO", "PRMT", "VIMNMX", "LDGSTS", "UBLKCP", "UTMALDG", "HMMA")]
        rows.append((short, len(lines), " ".join(notable)))
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# SASS listings (sm_100a)\n\n`python tools/dump_sass.py` -> `cuobjdump -sass bzip2_b200/libbz2_b200.so`, one file per kernel, "
                "mnemonics only.\nEvery kernel is integer/byte work: no tensor-core (`UTC*MMA`, `HMMA`) or TMA instructions are expected.\n\n"
                "| kernel | SASS instructions | warp / atomic / shared-memory ops present |\n|---|---:|---|\n")
        for r in sorted(rows):
            f.write(f"| `{r[0]}` | {r[1]} | {r[2]} |\n")
    print(f"{len(rows)} kernels -> {OUT}")


if __name__ == "__main__":
    sys.exit(main())
