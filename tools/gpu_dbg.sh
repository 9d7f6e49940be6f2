#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/debug_qloss.py > gpurun_out/dbg_qloss.log 2>&1; tail -20 gpurun_out/dbg_qloss.log
