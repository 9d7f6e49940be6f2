"""Benchmark of the hot path: posterior Langevin chain-steps/sec at CIFAR-10 shape (BASELINE.json:metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the reference algorithm's CPU port on the host cores)

A "step" is one call of sample_langevin_post_z_with_prior over this rank's batch: B chains x 30 Langevin steps
(reference defaults: K=30, s=0.1, sigma=0.1, noise on; train_gen_recon.py:383-386) on a CIFAR-10-shaped generator
(nz=128, ngf=128, 3x32x32) with default-init weights (seed 1) and synthetic images x = clamp(G(z*) + sigma n, -1, 1).
Chains are independent, so ranks shard them with no collective inside sampling ("scaling": "weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

NZ, NGF, NC, IMG = 128, 128, 3, 32
L_STEPS, STEP_SIZE, SIGMA = 30, 0.1, 0.1
# algorithmic work per chain-step: generator MACs 8.39M + 536.9M + 536.9M + 7.08M, forward + dgrad
GEN_MACS = 128 * 1024 * 64 + 1024 * 512 * 16 * 64 + 512 * 256 * 16 * 256 + 256 * 3 * 9 * 1024
FLOP_PER_CHAIN_STEP = 2 * 2 * GEN_MACS


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=float(d["bf16_tflops_sustained"]), hbm=float(d["hbm_gbs"]), src="measured (sustained bf16)")
    return dict(tflops=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def make_nets(device):
    import torch
    from damc_b200 import diffusion_net as dn
    torch.manual_seed(1)  # reference default seed (train_gen_recon.py:353); default nn init = synthetic weights
    G, E = dn._netG_cifar10(NZ, NGF, NC), dn._netE(NZ)
    return G.to(device).eval(), E.to(device).eval()


def make_inputs(G, B, device, seed):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    zstar = torch.randn(B, NZ, generator=g)
    xn = torch.randn(B, NC, IMG, IMG, generator=g)
    z0 = torch.randn(B, NZ, generator=g)
    with torch.no_grad():
        xs = []
        for i in range(0, B, 512):
            xs.append(torch.clamp(G(zstar[i:i + 512].to(device)) + SIGMA * xn[i:i + 512].to(device), -1.0, 1.0))
        x = torch.cat(xs, 0)
    return z0, x


def cpu_reference(B, K, warm=True):
    """The reference algorithm on the host cores: oracle port of MCMC.py:48-74 (same ATen ops, all threads)."""
    import torch
    from oracle import damc_oracle as O
    from damc_b200 import diffusion_net as dn
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(1)
    G, E = dn._netG_cifar10(NZ, NGF, NC), dn._netE(NZ)
    gen = [(G.gen[2 * i].weight.detach(), G.gen[2 * i].bias.detach(), G.gen[2 * i].stride[0], G.gen[2 * i].padding[0])
           for i in range(4)]
    ebm = [(E.ebm[2 * i].weight.detach(), E.ebm[2 * i].bias.detach()) for i in range(3)]
    z0, x = make_inputs(G, B, torch.device("cpu"), 123)
    if warm:
        O.langevin_posterior(z0[:8], x[:8], gen, ebm, 1, SIGMA, True, STEP_SIZE)
    t0 = time.perf_counter()
    O.langevin_posterior(z0, x, gen, ebm, K, SIGMA, True, STEP_SIZE)
    dt = time.perf_counter() - t0
    return B * K / dt, dt


def eager_gpu_reference(B, K, dev, reps=2):
    """The reference algorithm run the way the reference itself runs on a GPU (README: single GPU, eager PyTorch): the
    oracle port of MCMC.py:48-74 -- the same ATen / cuDNN calls and the same four .item() syncs per step -- on `dev`,
    torch defaults (TF32 allowed for the convolutions, as in the reference).  Baseline leg only; never the product path."""
    import torch
    from oracle import damc_oracle as O
    from damc_b200 import diffusion_net as dn
    torch.manual_seed(1)
    G, E = dn._netG_cifar10(NZ, NGF, NC), dn._netE(NZ)
    gen = [(G.gen[2 * i].weight.detach().to(dev), G.gen[2 * i].bias.detach().to(dev), G.gen[2 * i].stride[0],
            G.gen[2 * i].padding[0]) for i in range(4)]
    ebm = [(E.ebm[2 * i].weight.detach().to(dev), E.ebm[2 * i].bias.detach().to(dev)) for i in range(3)]
    z0, x = make_inputs(G, B, torch.device("cpu"), 123)
    z0, x = z0.to(dev), x.to(dev)
    O.langevin_posterior(z0, x, gen, ebm, 2, SIGMA, True, STEP_SIZE, trace=[])     # warm-up (cuDNN heuristics, allocator)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        O.langevin_posterior(z0, x, gen, ebm, K, SIGMA, True, STEP_SIZE, trace=[])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return B * K / (best * 1e-3), best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, K = args.cpu_chains, args.cpu_lsteps
    vals, times = [], []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_reference(B, K, warm=(i == 0))
        if i >= args.warmup:
            vals.append(v)
            times.append(dt)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": "posterior Langevin chain-steps/sec (CIFAR-10 shape)", "value": value,
            "unit": "chain-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cifar10 posterior Langevin, CPU sample of {B} chains x {K} Langevin steps per step",
                       "nz": NZ, "ngf": NGF, "image": [NC, IMG, IMG], "sigma": SIGMA, "step_size": STEP_SIZE},
            "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{B} chains x {K} Langevin steps per timed step (oracle/damc_oracle.py)"},
            "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from damc_b200 import MCMC, _lib
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    B = args.chains
    G, E = make_nets(dev)
    z0_host, x = make_inputs(G, B, dev, 1000 + rank)       # different chains on every rank
    z0_dev = z0_host.to(dev)
    chain0 = rank * B                                      # global chain index of this shard (Philox key)

    def step(z_init, seed):
        z = z_init.clone().requires_grad_(True)
        return MCMC.sample_langevin_post_z_with_prior(z, x, G, E, L_STEPS, SIGMA, True, STEP_SIZE, seed=seed,
                                                      chain0=chain0, precision=args.precision)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(z0_dev, i)
    barrier()
    # ---- timed region: inputs resident in HBM ---------------------------------------------------------------------
    # The production path: no measurement hooks, the K-step launch sequence replayed from its captured CUDA graph.
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = lib.damc_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        out = step(z0_dev, 100 + i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.damc_launch_count() - n0
    clk = clocks.stop() if rank == 0 else None
    assert torch.isfinite(out).all()
    # ---- separate pass, NOT part of `value`: per-launch CUDA events around every generator GEMM (direct launches, the
    # hooks switch graph replay off) -> time share and achieved FLOP/s of the dominant kernel --------------------------
    import ctypes
    prof_steps = max(1, min(args.steps, 3))
    lib.damc_profile_enable(1)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    pe0.record()
    for i in range(prof_steps):
        step(z0_dev, 300 + i)
    pe1.record()
    torch.cuda.synchronize()
    prof_ms = pe0.elapsed_time(pe1)
    gemm_ms, gemm_n = ctypes.c_double(), ctypes.c_longlong()
    lib.damc_profile_collect(ctypes.byref(gemm_ms), ctypes.byref(gemm_n))
    lib.damc_profile_enable(0)
    # ---- end to end: pinned host inputs -> H2D -> public API -> D2H of the chains ------------------------------------
    xh = x.cpu().pin_memory()
    zh = z0_host.pin_memory()
    res = torch.empty(B, NZ).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        xd = xh.to(dev, non_blocking=True)
        zd = zh.to(dev, non_blocking=True).requires_grad_(True)
        o = MCMC.sample_langevin_post_z_with_prior(zd, xd, G, E, L_STEPS, SIGMA, True, STEP_SIZE, seed=200 + i,
                                                   chain0=chain0, precision=args.precision)
        res.copy_(o, non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        pk = peaks()
        total_cs = world * B * L_STEPS * args.steps
        value = total_cs / (ms * 1e-3)
        gemm_flops = B * L_STEPS * prof_steps * FLOP_PER_CHAIN_STEP  # this rank's GEMM launches of the profiling pass
        achieved = gemm_flops / (gemm_ms.value * 1e-3) / 1e12 if gemm_ms.value > 0 else None
        step_tflops = (value / world) * FLOP_PER_CHAIN_STEP / 1e12   # whole step, per GPU, from the hook-free timing
        peak = pk["tflops"] * (0.5 if args.precision == "tf32" else 1.0)   # dense tf32 rate = half the bf16 rate
        cpu_v, cpu_dt = (cpu_reference(args.cpu_chains, args.cpu_lsteps) if world == 1 and not args.no_cpu else (None, None))
        eager = None
        if world == 1 and not args.no_eager:
            eager = {}
            for eb in (128, B):
                ev, ems = eager_gpu_reference(eb, L_STEPS, dev)
                eager[f"chains_{eb}"] = {"value": ev, "unit": "chain-steps/s", "ms_per_call": ems}
            eager["what"] = ("reference algorithm (oracle port of src/MCMC.py:48-74) in eager PyTorch on this B200, torch "
                             "defaults (cudnn TF32 convolutions, 4 .item() syncs per Langevin step), K=%d" % L_STEPS)
        secondary = None
        if world == 1 and not args.no_secondary:
            secondary = damc_secondary(dev)
        # DRAM traffic needs ncu counters; bench.py cannot measure it live.  The figure of the newest committed ncu capture
        # of this workload is quoted and labelled with its file; null when there is none for this precision / batch.
        traffic, traffic_note = None, "not measured in this run (needs ncu); no committed capture for this precision/batch"
        for tname in ("r02_traffic_B1024.json", "r01_traffic_B1024.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                if tj.get("chains") == B and tj.get("precision", "bf16") == args.precision:
                    traffic = tj["mean_dram_bytes_per_gemm_launch"]
                    traffic_note = ("mean dram__bytes_read+write per generator GEMM launch from the committed ncu capture "
                                    f"profiles/{tname} (same workload, earlier run) -- not measured in this run")
                    break
        line = {
            "metric": "posterior Langevin chain-steps/sec (CIFAR-10 shape)", "value": value, "unit": "chain-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "tf32": "tf32"}.get(args.precision, "f32"), "data": "synthetic",
            "config": {"workload": f"cifar10 (configs[2]): {B} chains/GPU x {L_STEPS} Langevin steps per step",
                       "nz": NZ, "ngf": NGF, "image": [NC, IMG, IMG], "sigma": SIGMA, "step_size": STEP_SIZE,
                       "noise": "philox", "parallelism": f"chains sharded x{world}, no collective in sampling",
                       "l2": "working set (activations+gradients, %.1f GB/GPU) exceeds the 126 MB L2; no flush needed"
                             % (B * 461824 * 2 * (2 if args.precision in ("bf16", "fp16") else 4) / 1e9),
                       "precision": args.precision,
                       "timing": "value: CUDA events around the hook-free production path (CUDA-graph replay); "
                                 "roofline.frac_gemm from a separate pass with per-launch events"},
            "clocks": clk,
            "e2e": {"value": world * B * L_STEPS * args.steps / (e2e_ms * 1e-3), "unit": "chain-steps/s",
                    "h2d_bytes_per_step": B * (NC * IMG * IMG + NZ) * 4, "d2h_bytes_per_step": B * NZ * 4},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         "frac_gemm": (achieved / peak) if achieved else None,
                         "achieved_step": step_tflops, "frac_step": step_tflops / peak,
                         "traffic": traffic, "traffic_note": traffic_note,
                         "peak_source": pk["src"] + ("; tf32 peak = half of it" if args.precision == "tf32" else ""),
                         "kernel": "convgemm_tc_kernel (+ last_fused_kernel, which contains the last layer's two GEMMs): generator "
                                   "implicit-GEMM launches (%d in the profiling pass of %d "
                                   "steps, %.1f%% of that pass); frac / frac_gemm = algorithmic FLOPs / summed per-launch "
                                   "event time; frac_step = the same FLOPs / whole-step time of the hook-free timed region"
                                   % (gemm_n.value, prof_steps, 100.0 * gemm_ms.value / prof_ms)},
            "eager_gpu_baseline": eager,
            "secondary": secondary,
            "cpu_baseline": None if cpu_v is None else {
                "value": cpu_v, "unit": "chain-steps/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"{args.cpu_chains} chains x {args.cpu_lsteps} Langevin steps, {cpu_dt:.1f} s "
                          "(oracle/damc_oracle.py langevin_posterior, torch CPU, all threads)"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def damc_secondary(dev):
    """The other sampler of the path, reported next to the headline (never part of it): the DAMC reverse loop
    (reference workspace/src/diffusion_net.py:597-620; T = 100, xemb and z_T resident, Philox noise, fp16 operands) on the
    hoisted-context schedule (csrc/denoiser_seq.cu), CUDA-event timed.  Any failure is reported, not raised."""
    try:
        import torch
        from damc_b200 import MCMC, diffusion_net as dn
        torch.manual_seed(1)
        T = 100
        Q = dn._netQ_U(nc=NC, nz=NZ, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=T, logsnr_min=-5.1,
                       logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev).eval()
        out = {"what": "DAMC sampler loop, T = 100 reverse steps, fp16 operands, xemb / z_T resident; 2.949 MFLOP per chain and step"}
        for Bq in (128, 16384):
            xemb = torch.randn(Bq, 1024, device=dev) * 0.5
            zT = torch.randn(Bq, NZ, device=dev)
            run = lambda: MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision="fp16")
            for _ in range(2):
                run()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                run()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / 3
            out[f"chains_{Bq}"] = {"ms": ms, "reverse_steps_per_s": Bq * T / ms * 1e3, "tflops": Bq * T * 2.949e6 / ms * 1e3 / 1e12}
        return out
    except Exception as exc:  # noqa: BLE001 -- the headline line must not depend on this leg
        return {"error": repr(exc)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("DAMC_BENCH_PRECISION", "bf16"),
                    choices=["bf16", "fp16", "tf32", "fp32"])
    ap.add_argument("--chains", type=int, default=int(os.environ.get("DAMC_BENCH_CHAINS", "1024")),
                    help="chains per GPU")
    ap.add_argument("--cpu-chains", type=int, default=128)
    ap.add_argument("--cpu-lsteps", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch-on-GPU baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the DAMC sampler timings reported under 'secondary'")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
