#!/bin/bash
# engines per GPU sweep (two or three windows in flight on one B200), 1 GB text at -9
for E in 1 2 3; do
  echo "== engines per GPU = $E"
  python bench.py --steps 3 --warmup 2 --no-c4 --no-cpu --engines-per-gpu $E 2> gpurun_out/r2_eng$E.err | tail -1 > gpurun_out/r2_eng$E.json
  python - <<PY
import json
l = json.load(open("gpurun_out/r2_eng$E.json"))
print("value", l["value"], "ms/step", l["ms_per_step"], "e2e", l["e2e"]["value"], "stage_ms(one engine)", l["roofline"]["stage_ms"], "launches", l["gpu_launches"])
PY
done
