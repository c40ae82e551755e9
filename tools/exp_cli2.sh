#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
echo start > $O/cli2.log
python - <<'PY' >> $O/cli2.log 2>&1
import sys
sys.path.insert(0, "tests")
import support as S
d = S.gen_text(2_000_000_000).tobytes()
open("/dev/shm/b2.txt", "wb").write(d)
open("/dev/shm/tiny.txt", "wb").write(d[:1000])
PY
TIMEFORMAT="%R s"
export BZ2_B200_CLI_TIMING=1
for f in tiny b2 tiny b2; do
  { echo -n "fast exit $f  "; { time bzip2_b200/bzip2-b200 -9 -c /dev/shm/$f.txt > /dev/null; } 2>&1; } >> $O/cli2.log
  { echo -n "slow exit $f  "; { time BZ2_B200_CLI_SLOWEXIT=1 bzip2_b200/bzip2-b200 -9 -c /dev/shm/$f.txt > /dev/null; } 2>&1; } >> $O/cli2.log
done
{ echo -n "tiny window 8MB  "; { time BZ2_B200_WINDOW_MB=8 bzip2_b200/bzip2-b200 -9 -c /dev/shm/tiny.txt > /dev/null; } 2>&1; } >> $O/cli2.log
bzip2_b200/bzip2-b200 -9 -c /dev/shm/b2.txt | sha256sum >> $O/cli2.log
python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from bzip2_b200 import binding as B
import numpy as np, hashlib
d=np.fromfile('/dev/shm/b2.txt',dtype=np.uint8)
print(hashlib.sha256(B.Engine(level=9).compress(d)).hexdigest(), ' one-shot')
" >> $O/cli2.log 2>&1
rm -f /dev/shm/b2.txt /dev/shm/tiny.txt
