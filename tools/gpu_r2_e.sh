#!/bin/bash
# round 2, call E: fused last-layer kernel with pipelined x loads: parity tests, launch lists, per-config timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "fused or bf16_golden or x_hat" > gpurun_out/e_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/e_pytest.log
tail -4 gpurun_out/e_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "svhn 16384 1 bf16" "celebaHQ 128 1 bf16" "cifar10 1024 1 bf16" "mnist 4096 1 bf16"; do
  set -- $cfg
  timeout 900 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/e_launches_$1_$4_B$2.csv \
    python tools/profile_config.py $1 $2 $3 $4 > gpurun_out/e_ncu_$1.log 2>&1
  echo "ncu $cfg exit $?"
done
PREC=bf16 timeout 900 python tools/bench_configs.py > gpurun_out/e_configs_bf16.log 2>&1
tail -1 gpurun_out/e_configs_bf16.log
