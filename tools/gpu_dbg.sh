#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python tools/debug_ebm_tc.py > gpurun_out/dbg_ebm.log 2>&1; tail -40 gpurun_out/dbg_ebm.log
