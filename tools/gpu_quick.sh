#!/bin/bash
# quick GPU check: pytest with the -k expression in $1 (log: gpurun_out/q_pytest.log)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider -k "$1" > gpurun_out/q_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/q_pytest.log
grep -E "rel err|passed|failed|Error|error|assert" gpurun_out/q_pytest.log | tail -40
