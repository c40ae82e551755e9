#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
rm -f $O/chain.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stage_outputs or golden_streams or fuzz_small or all_levels or reference_kat or long_repeats or tiny" > $O/pytest_chain.log 2>&1; echo "pytest rc=$?" >> $O/chain.log
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
for WL in text random; do
  run WL=$WL BZ2_B200_CHAIN=1
done
