#!/bin/bash
# round 2, call F: launch lists of the fused last-layer kernel (quick A/B of a kernel change)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_last_fused.py -m gpu -q -x -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/f_pytest.log
tail -3 gpurun_out/f_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "svhn 16384 1 bf16" "celebaHQ 128 1 bf16" "cifar10 1024 1 bf16" "mnist 4096 1 bf16"; do
  set -- $cfg
  timeout 900 ncu --metrics $M --clock-control none -k regex:last_fused -c 4 --csv --log-file gpurun_out/f_launches_$1_$4_B$2.csv \
    python tools/profile_config.py $1 $2 $3 $4 > gpurun_out/f_ncu_$1.log 2>&1
  echo "ncu $cfg exit $?"
  grep last_fused gpurun_out/f_launches_$1_$4_B$2.csv | grep gpu__time | tail -1
done
