#!/bin/bash
# round 2, call C: the fused last-layer kernel -- its own tests first (bounded: a protocol bug traps instead of hanging),
# then the whole GPU suite, the bench and per-config timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_last_fused.py -m gpu -q -s -x -p no:cacheprovider > gpurun_out/c_fused.log 2>&1
rc=$?
echo "fused tests exit $rc" >> gpurun_out/c_fused.log
tail -25 gpurun_out/c_fused.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/c_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c_pytest.log
tail -8 gpurun_out/c_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/c_bench_bf16.json 2> gpurun_out/c_bench_bf16.err
cat gpurun_out/c_bench_bf16.json
PREC=bf16 timeout 900 python tools/bench_configs.py > gpurun_out/c_configs_bf16.log 2>&1
tail -1 gpurun_out/c_configs_bf16.log
