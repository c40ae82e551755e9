#!/bin/bash
# final numbers of the round: full GPU suite, smoke, default bench (both arms), the other shapes
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
rm -f $O/final.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/final.log
timeout 600 python bench.py > $O/bench_final_n1.log 2> $O/bench_final_n1.err; echo "bench rc=$?" >> $O/final.log
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_final_ref.log 2> $O/bench_final_ref.err; echo "ref rc=$?" >> $O/final.log
for wl in random period1000 aab runs mixed; do
  timeout 300 python bench.py --mb 400 --steps 3 --warmup 2 --no-cpu --workload $wl > $O/bench_final_$wl.log 2>&1; echo "$wl rc=$?" >> $O/final.log
done
