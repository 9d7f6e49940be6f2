#!/bin/bash
# round 2, call I: SVHN tuning A/B -- EBM step chain tile (DAMC_EBM_CH) and the fused last-layer kernel, ncu per-kernel times
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
M=gpu__time_duration.sum
for ch in 4 8 16; do
  DAMC_EBM_CH=$ch timeout 600 ncu --metrics $M --clock-control none -k regex:"ebm_step|last_fused" -c 4 --csv --log-file gpurun_out/i_svhn_ch$ch.csv \
    python tools/profile_config.py svhn 16384 1 bf16 > gpurun_out/i_ncu_$ch.log 2>&1
  echo "ch=$ch"; grep -E "ebm_step|last_fused" gpurun_out/i_svhn_ch$ch.csv | tail -2 | awk -F'","' '{print $5, $NF}'
done
timeout 600 ncu --metrics $M --clock-control none -k regex:"last_fused" -c 2 --csv --log-file gpurun_out/i_celeba.csv python tools/profile_config.py celebaHQ 128 1 bf16 > gpurun_out/i_ncu_c.log 2>&1
grep -E "last_fused" gpurun_out/i_celeba.csv | tail -1 | awk -F'","' '{print $5, $NF}'
timeout 600 ncu --metrics $M --clock-control none -k regex:"last_fused" -c 2 --csv --log-file gpurun_out/i_cifar.csv python tools/profile_config.py cifar10 1024 1 bf16 > gpurun_out/i_ncu_d.log 2>&1
grep -E "last_fused" gpurun_out/i_cifar.csv | tail -1 | awk -F'","' '{print $5, $NF}'
