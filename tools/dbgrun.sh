# A/B harness: per-kernel durations of one posterior step (ncu launch list) under experiment switches
for d in ${DBGS:-0}; do
  DAMC_TC_DBG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"convgemm|last_finish|ebm_step" --launch-skip 30 -c 10 --csv --log-file gpurun_out/dbg_${TAG:-x}_$d.csv python tools/profile_step.py 1024 2 > /dev/null 2>&1
done
echo done
