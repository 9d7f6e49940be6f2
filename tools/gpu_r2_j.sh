#!/bin/bash
# round 2, call J: smoke(); ncu --set full of the fused last-layer kernel (CIFAR-10 1 024 chains, SVHN 16 384 chains)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/j_smoke.log 2>&1; echo "smoke exit $?"; tail -12 gpurun_out/j_smoke.log
for cfg in "cifar10 1024" "svhn 16384"; do
  set -- $cfg
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:last_fused -s 1 -c 1 -o gpurun_out/j_full_$1 -f \
    python tools/profile_config.py $1 $2 1 bf16 > gpurun_out/j_ncu_$1.log 2>&1
  echo "ncu full $cfg exit $?"
done
ls -la gpurun_out/*.ncu-rep
