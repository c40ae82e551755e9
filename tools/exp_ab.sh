#!/bin/bash
# A/B runs of the BWT-stage switches on one GPU (scratch experiment driver; results land in gpurun_out/).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
rm -f $O/ab.log
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/ab.log
for hl in 18 20 22; do
  for wl in text mixed; do
    echo "== hist_log2=$hl $wl" >> $O/ab.log
    BZ2_B200_HIST_LOG2=$hl python bench.py --mb 400 --steps 3 --warmup 2 --no-e2e --no-cpu --workload $wl >> $O/ab.log 2>&1
  done
done
BZ2_B200_HIST_LOG2=20 BZ2_B200_TRACE=1 python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/trace_text20.log 2>&1
BZ2_B200_HIST_LOG2=20 python -m pytest tests -x -q -m gpu -k "golden or fuzz or stage or level or corner or tiny" > $O/pytest_h20.log 2>&1; echo "pytest h20 rc=$?" >> $O/ab.log
