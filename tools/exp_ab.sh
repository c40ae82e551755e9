#!/bin/bash
# scratch experiment driver; results land in gpurun_out/
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
rm -f $O/ab.log
python -m pytest tests -x -q -m gpu --timeout 1200 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/ab.log
for wl in text random period1000; do
  echo "== $wl" >> $O/ab.log
  python bench.py --mb 400 --steps 3 --warmup 2 --no-e2e --no-cpu --workload $wl >> $O/ab.log 2>&1
done
