#!/bin/bash
# round 2, call P: DAMC denoiser after the balanced ctx mapping in den_prep_kernel: parity tests + timings + one-step launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "damc or denoiser or amortizer or toy" > gpurun_out/p_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/p_pytest.log
timeout 600 python tools/bench_denoiser.py 128 4096 16384 > gpurun_out/p_den.log 2>&1; tail -5 gpurun_out/p_den.log
