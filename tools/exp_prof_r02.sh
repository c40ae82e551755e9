#!/bin/bash
# round-2 profiles (one engine, 100 MB windows at -9): launch lists with DRAM bytes for text and the C4 mix, and one
# ncu --set full capture of the heaviest kernels of the text window.  Each command first runs without ncu.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
: > $O/r02_prof.log
for wl in text mixed; do
  mb=100; [ $wl = mixed ] && mb=200
  python tools/prof_step.py --workload $wl --mb $mb > $O/r02_plain_$wl.log 2>&1 || { echo "plain $wl failed" >> $O/r02_prof.log; continue; }
  cat $O/r02_plain_$wl.log >> $O/r02_prof.log
  timeout 600 ncu --metrics $M --clock-control none --csv --log-file $O/r02_launches_dram_$wl.csv \
      python tools/prof_step.py --workload $wl --mb $mb > $O/r02_ncu_$wl.log 2>&1
  echo "$wl launch list rc=$?" >> $O/r02_prof.log
done
python tools/prof_step.py --workload text --mb 100 --warmup 0 > $O/r02_plain_full.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"k_refine_large|k_kgram|k_refine_radix|k_refine_medium|k_refine_small|k_mtf_encode|k_rle2_emit|k_tile|k_bwt_out|k_crc" -c 26 \
    -o $O/r02_top python tools/prof_step.py --workload text --mb 100 --warmup 0 > $O/r02_ncu_full.log 2>&1
echo "ncu full text rc=$?" >> $O/r02_prof.log
ncu -i $O/r02_top.ncu-rep --page raw --csv > $O/r02_ncu_full_raw.csv 2>> $O/r02_prof.log
ncu -i $O/r02_top.ncu-rep --page source --csv --kernel-name regex:k_kgram > $O/r02_ncu_source_kgram.csv 2>> $O/r02_prof.log
rm -f $O/r02_top.ncu-rep
cat $O/r02_prof.log
