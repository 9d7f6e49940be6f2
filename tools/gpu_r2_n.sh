#!/bin/bash
# round 2, call N: what bounds the first layer's forward (CT1 fwd: K = nz, N = k*k*C1, output-write heavy)?  DAMC_TC_DBG ablations:
# 1 = no global stores, 2 = no epilogue math/stores, 32 / 64 / 128 = the same stores confined to 1 / 32 / 256 MB of the output
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for dbg in 0 1 2 32 64 128; do
  DAMC_TC_DBG=$dbg timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum --clock-control none -k regex:convgemm -c 30 --csv --log-file gpurun_out/n_svhn_dbg$dbg.csv \
    python tools/profile_config.py svhn 16384 1 bf16 > gpurun_out/n_ncu.log 2>&1
  echo "dbg=$dbg"; grep -E "convgemm_tc_kernel<16" gpurun_out/n_svhn_dbg$dbg.csv | tail -3 | awk -F'","' '{print $(NF-2), $NF}'
done
