#!/bin/bash
# SURVEY 8(d) C3 (periodic inputs, 1 GB at -9) and random bytes, one GPU, two engines: one bench line each
for wl in period1000 aab runs random; do
  python bench.py --workload $wl --steps 5 --warmup 2 --no-c4 --no-cpu 2>/dev/null | grep "^{" > gpurun_out/r2_final_$wl.json
  python - <<PY
import json
l = json.load(open("gpurun_out/r2_final_$wl.json"))
print("$wl", "value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], l["roofline"]["stage_ms"], "rounds", l["bwt_rounds"])
PY
done
