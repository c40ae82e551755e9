#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
rm -f $O/chain.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_repeats or chains_from or periodic_segments or stage_outputs or exact_power or fuzz_small or golden_streams" > $O/pytest_chain.log 2>&1; echo "pytest rc=$?" >> $O/chain.log
run() {
  echo "== $*" >> $O/chain.log
  env "$@" timeout 300 python bench.py --mb 400 --steps 3 --warmup 2 --no-e2e --no-cpu --workload $WL 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: j = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(j['value'], j['ms_per_step'], j['roofline']['stage_ms'], 'rounds', j['bwt_rounds'])
" >> $O/chain.log 2>&1
}
for WL in text mixed period1000; do
  run WL=$WL BZ2_B200_CHAIN=1
done
