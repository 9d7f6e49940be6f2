#!/bin/bash
# round 2, call M: EBM step chain tile 16 vs 32 at 16 384 SVHN chains
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for ch in 16 32; do
  DAMC_EBM_CH=$ch timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ebm_step" -c 2 --csv --log-file gpurun_out/m_svhn_ch$ch.csv \
    python tools/profile_config.py svhn 16384 1 bf16 > gpurun_out/m_ncu_$ch.log 2>&1
  echo "ch=$ch"; grep -E "ebm_step" gpurun_out/m_svhn_ch$ch.csv | tail -1 | awk -F'","' '{print $5, $NF}'
done
