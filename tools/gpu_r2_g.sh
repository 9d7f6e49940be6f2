#!/bin/bash
# round 2, call G: clock breakdown of the fused last-layer kernel's warp groups (DAMC_LAST_DEBUG=1, CTA 0)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for cfg in "svhn 16384 1 bf16" "celebaHQ 128 1 bf16" "cifar10 1024 1 bf16"; do
  set -- $cfg
  echo "== $cfg" >> gpurun_out/g_debug.log
  DAMC_LAST_DEBUG=1 timeout 300 python tools/profile_config.py $1 $2 $3 $4 >> gpurun_out/g_debug.log 2>&1
done
cat gpurun_out/g_debug.log
