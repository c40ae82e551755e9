#!/bin/bash
# one ncu --set full capture of the heaviest kernels of the current build, 100 MB text window
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_plain.log 2>&1 || exit 1
ncu --set full --clock-control none \
    -k regex:"k_refine_large|k_refine_radix|k_mtf_encode$|k_refine_small<2|k_refine_medium<32|k_rle2_emit|k_tile<2>|k_mtf_lists" -c 24 \
    --csv --page raw --log-file $O/r01_v3_ncu_full_raw.csv python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_run.log 2>&1
echo "ncu rc=$?" >> $O/ncu_run.log
