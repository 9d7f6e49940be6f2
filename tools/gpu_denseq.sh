#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/debug_denseq.py ${1:-all} > gpurun_out/denseq.log 2>&1
echo "exit $?" >> gpurun_out/denseq.log
tail -30 gpurun_out/denseq.log
