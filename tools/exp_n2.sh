#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
for m in inline proc; do
  BENCH_CLOCKS=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > $O/bench_n2_$m.log 2> $O/bench_n2_$m.err; echo "$m rc=$?"
done
