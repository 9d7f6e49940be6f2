#!/bin/bash
# round 2, call A: full GPU test suite (no -x: collect every failure), bench in bf16 and tf32
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_bf16.json 2> gpurun_out/a_bench_bf16.err
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu > gpurun_out/a_bench_tf32.json 2> gpurun_out/a_bench_tf32.err
tail -5 gpurun_out/a_pytest.log
cat gpurun_out/a_bench_bf16.json gpurun_out/a_bench_tf32.json
