for d in ${DBGS:-0 4}; do
  DAMC_TC_DBG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"convgemm|last_finish|ebm_step" --launch-skip 30 -c 10 --csv --log-file gpurun_out/dbg_$d.csv python tools/profile_step.py 1024 2 > /dev/null 2>&1
done
echo done
