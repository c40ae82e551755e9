#!/bin/bash
# one ncu --set full capture (with source counters) of the heaviest BWT kernels, 100 MB text window
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none \
    -k regex:"k_refine_large|k_refine_medium|k_kgram|k_mtf_encode|k_rle2_emit|k_mtf_lists" -c 14 \
    -o $O/r01b_top python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_run.log 2>&1
ls -la $O/r01b_top.ncu-rep >> $O/ncu_run.log
