#!/bin/bash
# L2 residency of BWT sub-batches (VERDICT r1 next #3 i): per-kernel time, DRAM bytes and L2 hit rate with warm caches
# (--cache-control none), 100 MB text window at -9, for several sub-batch sizes.
O=gpurun_out
for G in 0 4 8; do
  BZ2_B200_S2_GROUP=$G ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
     --clock-control none --cache-control none --csv --log-file $O/r02_group${G}_launches.csv \
     python bench.py --mb 100 --steps 1 --warmup 1 --no-e2e --no-cpu > $O/r02_group${G}_ncu.log 2>&1
  echo "G=$G rc=$?"
  python tools/traffic_summary.py $O/r02_group${G}_launches.csv --mb 100 | head -12
done
