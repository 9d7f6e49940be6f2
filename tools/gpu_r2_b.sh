#!/bin/bash
# round 2, call B: re-run the two tests fixed after call A; ncu launch lists (time, DRAM bytes, tensor pipe) of one
# Langevin step on SVHN, CelebA-HQ and CIFAR-10 in tf32
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_training_step.py -m gpu -q -s -p no:cacheprovider -k "fp32_golden or training_iteration" > gpurun_out/b_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/b_pytest.log
tail -3 gpurun_out/b_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "svhn 16384 1 bf16" "celebaHQ 128 1 bf16" "cifar10 1024 1 tf32" "mnist 4096 1 bf16"; do
  set -- $cfg
  timeout 600 python tools/profile_config.py $1 $2 $3 $4 > gpurun_out/b_plain_$1.log 2>&1 || { echo "plain run failed: $cfg"; continue; }
  timeout 900 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/b_launches_$1_$4_B$2.csv \
    python tools/profile_config.py $1 $2 $3 $4 > gpurun_out/b_ncu_$1.log 2>&1
  echo "ncu $cfg exit $?"
done
PREC=bf16 timeout 900 python tools/bench_configs.py svhn celebaHQ > gpurun_out/b_configs_bf16.log 2>&1
tail -1 gpurun_out/b_configs_bf16.log
