#!/bin/bash
# S2 sub-batch size sweep (VERDICT r1 next #3 i): blocks sorted together per BWT sub-batch, 400 MB text at -9
for G in 0 4 6 8 12 16 32; do
  echo "== S2_GROUP=$G"
  BZ2_B200_S2_GROUP=$G python bench.py --mb 400 --steps 3 --warmup 2 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
l = json.loads(sys.stdin.read())
print(l['value'], l['ms_per_step'], l['roofline']['stage_ms'], l['gpu_launches'], l['bwt_rounds'])"
done
