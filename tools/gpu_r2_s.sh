#!/bin/bash
# round 2, call S: prior sampler on the tensor cores: test + timings (256 / 4096 / 16384 chains x 60 steps, fp32 kernel vs fp16 form)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -p no:cacheprovider -k "prior_langevin_on_tensor or ebm_tail or ebm_gradient" > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?"; grep -E "prior K=|passed|failed" gpurun_out/s_pytest.log | tail -4
timeout 600 python - <<'PY' > gpurun_out/s_prior.log 2>&1
import sys, os, json, torch
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "diffusion-amortized-mcmc_b200")]
from damc_b200 import MCMC, diffusion_net as dn
dev = torch.device("cuda:0")
torch.manual_seed(1)
E = dn._netE(128).to(dev).eval()
out = {}
for B in (256, 1024, 4096, 16384, 65536):
    z0 = torch.randn(B, 128, device=dev)
    for prec in ("fp32", "fp16"):
        fn = lambda: MCMC.sample_langevin_prior_z(z0.clone().requires_grad_(True), E, 60, 0.4, True, seed=1, precision=prec)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[f"prior_B{B}_K60_{prec}"] = {"ms": round(ms, 3), "chain_steps_per_s": round(B * 60 / ms * 1e3)}
print(json.dumps(out))
PY
tail -1 gpurun_out/s_prior.log
