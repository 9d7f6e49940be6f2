#!/bin/bash
# round 2, call L: A/B of spinning vs sleeping accumulator waits of the epilogue warps on short-K launches (DAMC_TC_NOSPIN=1: sleep)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for ns in 0 1; do
  if [ $ns = 1 ]; then export DAMC_TC_NOSPIN=1; else unset DAMC_TC_NOSPIN; fi
  for cfg in "cifar10 1024" "svhn 16384" "cifar10 128"; do
    set -- $cfg
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:convgemm -c 40 --csv --log-file gpurun_out/l_$1_$2_ns$ns.csv \
      python tools/profile_config.py $1 $2 1 bf16 > gpurun_out/l_ncu.log 2>&1
    echo "nospin=$ns $cfg"; grep convgemm gpurun_out/l_$1_$2_ns$ns.csv | tail -7 | awk -F'","' '{printf "%s %s | ", substr($5,25,8), $NF}'; echo
  done
done
unset DAMC_TC_NOSPIN
timeout 600 python tools/bench_denoiser.py 128 4096 16384 > gpurun_out/l_den_spin.log 2>&1; tail -8 gpurun_out/l_den_spin.log
DAMC_TC_NOSPIN=1 timeout 600 python tools/bench_denoiser.py 128 4096 16384 > gpurun_out/l_den_sleep.log 2>&1; tail -8 gpurun_out/l_den_sleep.log
