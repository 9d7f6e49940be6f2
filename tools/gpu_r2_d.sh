#!/bin/bash
# round 2, call D: CTA pairs for 128-wide layers + larger EBM chain tiles: SVHN / shard-invariance tests, launch lists with
# the fused last-layer kernel, per-config timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "svhn or celeba or shard or invariant or fused or graph" > gpurun_out/d_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/d_pytest.log
tail -4 gpurun_out/d_pytest.log
PREC=bf16 timeout 900 python tools/bench_configs.py > gpurun_out/d_configs_bf16.log 2>&1
tail -1 gpurun_out/d_configs_bf16.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "svhn 16384 1 bf16" "celebaHQ 128 1 bf16" "cifar10 1024 1 bf16"; do
  set -- $cfg
  timeout 900 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/d_launches_$1_$4_B$2.csv \
    python tools/profile_config.py $1 $2 $3 $4 > gpurun_out/d_ncu_$1.log 2>&1
  echo "ncu $cfg exit $?"
done
