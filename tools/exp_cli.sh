#!/bin/bash
# CLI / streaming-API throughput on tmpfs files (scratch experiment driver)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
echo start > $O/cli.log
python - <<'PY' >> $O/cli.log 2>&1
import sys
sys.path.insert(0, "tests")
import support as S
d = S.gen_text(2_000_000_000).tobytes()
open("/dev/shm/b2.txt", "wb").write(d)
with open("/dev/shm/b4.txt", "wb") as f:
    f.write(d); f.write(d)
open("/dev/shm/tiny.txt", "wb").write(d[:1000])
PY
TIMEFORMAT="%R s"
for f in tiny b2 b4 b2; do
  { echo -n "cli -9c $f  "; { time BZ2_B200_CLI_TIMING=1 bzip2_b200/bzip2-b200 -9 -c /dev/shm/$f.txt > /dev/null; } 2>&1; } >> $O/cli.log
done
{ echo -n "python one-shot 2GB (BuffToBuff, incl. import+init) "; { time python -c "
import sys,time; sys.path.insert(0,'.')
from bzip2_b200 import binding as B
import numpy as np
d=np.fromfile('/dev/shm/b2.txt',dtype=np.uint8)
t=time.time(); e=B.Engine(level=9); t1=time.time(); o=e.compress(d); t2=time.time(); o=e.compress(d); t3=time.time()
print('engine create %.2f s, first compress %.2f s, second %.2f s'%(t1-t,t2-t1,t3-t2))
"; } 2>&1; } >> $O/cli.log
rm -f /dev/shm/b2.txt /dev/shm/b4.txt /dev/shm/tiny.txt
