#!/bin/bash
# round-1 v6 profiles: launch lists (time + DRAM bytes) of text / mixed / period-1000 windows and ncu --set full
# captures of the heaviest kernels.  Each command first runs without ncu.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
rm -f $O/v6.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for wl in text mixed period1000; do
  mb=100; [ $wl = mixed ] && mb=200
  python bench.py --mb $mb --steps 1 --warmup 1 --no-e2e --no-cpu --workload $wl > $O/v6_plain_$wl.log 2>&1 || { echo "plain $wl failed" >> $O/v6.log; continue; }
  ncu --metrics $M --clock-control none --csv --log-file $O/r01_v6_launches_dram_$wl.csv \
      python bench.py --mb $mb --steps 1 --warmup 1 --no-e2e --no-cpu --workload $wl > $O/v6_ncu_$wl.log 2>&1
  echo "$wl ncu rc=$?" >> $O/v6.log
done
python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_refine_large|k_kgram|k_refine_radix|k_refine_medium|k_refine_small|k_mtf_encode|k_rle2_emit|k_tile|k_bwt_out" -c 24 \
    -o $O/r01_v6_top python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu > $O/ncu_run.log 2>&1
echo "ncu full text rc=$?" >> $O/v6.log
ncu -i $O/r01_v6_top.ncu-rep --page raw --csv > $O/r01_v6_ncu_full_raw.csv 2>> $O/v6.log
rm -f $O/r01_v6_top.ncu-rep
python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu --workload mixed > $O/ncu_plain_mixed.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_rep_|k_resolve_periodic" -c 14 \
    -o $O/r01_v6_rep python bench.py --mb 100 --steps 1 --warmup 0 --no-e2e --no-cpu --workload mixed > $O/ncu_run_mixed.log 2>&1
echo "ncu full mixed rc=$?" >> $O/v6.log
ncu -i $O/r01_v6_rep.ncu-rep --page raw --csv > $O/r01_v6_ncu_full_rep_raw.csv 2>> $O/v6.log
rm -f $O/r01_v6_rep.ncu-rep
