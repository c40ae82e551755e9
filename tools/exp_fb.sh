#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
T='tests/test_gpu_parity.py::test_sharded_stream_equals_single'
for i in 1 2; do
  python -m pytest "$T" -x -q -k fb > $O/fb_default_$i.log 2>&1; echo "default $i rc=$?" >> $O/fb.log
  BZ2_B200_S2_STREAMS=0 python -m pytest "$T" -x -q -k fb > $O/fb_nostreams_$i.log 2>&1; echo "nostreams $i rc=$?" >> $O/fb.log
done
BZ2_B200_DEBUG_SYNC=1 python -m pytest "$T" -x -q -k fb > $O/fb_dbg.log 2>&1; echo "dbg rc=$?" >> $O/fb.log
python -m pytest tests -q -m gpu --deselect "$T" -k "not test_gpu_parity or sharded or cli or scan or concat or window or roundtrip or stream" > $O/pytest_rest.log 2>&1; echo "rest rc=$?" >> $O/fb.log
