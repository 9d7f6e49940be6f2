#!/bin/bash
# round 2, call O: EBM tail on the tensor cores -- tests, per-config throughput with it on and off
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "ebm or shard or invariant or fullsize or statistic or mmd" > gpurun_out/o_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/o_pytest.log
PREC=bf16 timeout 900 python tools/bench_configs.py > gpurun_out/o_configs_tc.log 2>&1; tail -1 gpurun_out/o_configs_tc.log
DAMC_EBM_TC=0 PREC=bf16 timeout 900 python tools/bench_configs.py > gpurun_out/o_configs_cc.log 2>&1; tail -1 gpurun_out/o_configs_cc.log
