#!/bin/bash
# A/B of one environment switch inside one box: tools/exp_ab.sh VAR A B [bench args...]
VAR=$1; A=$2; B=$3; shift 3
for rep in 1 2; do
  for V in $A $B; do
    env $VAR=$V python bench.py --steps 4 --warmup 2 --no-c4 --no-cpu "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print('$VAR=$V', 'value', l['value'], 'ms', l['ms_per_step'], 'e2e', (l['e2e'] or {}).get('value'), 's2(one engine)', l['roofline']['stage_ms']['s2_bwt'])"
  done
done
