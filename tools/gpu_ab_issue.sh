#!/bin/bash
# A/B on one box: converged-warp MMA issue (shipped) vs lane-0 issue (tools/ab/libdamc_b200_lane0.so, built with -DDAMC_TC_ISSUE_LANE0)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SO=diffusion-amortized-mcmc_b200/damc_b200/libdamc_b200.so
cp $SO /tmp/new.so
for rep in 1 2; do
  cp /tmp/new.so $SO
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-eager > gpurun_out/ab_new_$rep.json 2>/dev/null
  cp tools/ab/libdamc_b200_lane0.so $SO
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-eager > gpurun_out/ab_lane0_$rep.json 2>/dev/null
done
cp /tmp/new.so $SO
python - <<'PY'
import json
for n in ("new_1","lane0_1","new_2","lane0_2"):
    d=json.loads(open(f"gpurun_out/ab_{n}.json").read().strip().splitlines()[-1])
    print(n, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac_step"],3), round(d["roofline"]["frac_gemm"],3), d["clocks"]["sm_mhz"])
PY
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "posterior or fullsize or last_fused" 2>&1 | tail -3
