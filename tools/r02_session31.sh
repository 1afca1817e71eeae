#!/bin/bash
# Round-2 session 31 (1 GPU): smoke() on the final library
set -u
OUT=gpurun_out/r02_s31
mkdir -p $OUT
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1 ; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
