#!/bin/bash
# round 2, validation run of the current build: full GPU suite, bench (bf16 with CPU + eager baselines, tf32), launch lists
# (time / DRAM bytes / tensor pipe) of one Langevin step per config, per-config throughput in bf16 and tf32
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/z_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/z_pytest.log
tail -6 gpurun_out/z_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/z_bench_bf16.json 2> gpurun_out/z_bench_bf16.err
cat gpurun_out/z_bench_bf16.json
timeout 600 python bench.py --steps 5 --warmup 3 --precision tf32 --no-cpu --no-eager > gpurun_out/z_bench_tf32.json 2> gpurun_out/z_bench_tf32.err
cat gpurun_out/z_bench_tf32.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "cifar10 1024 1 bf16" "svhn 16384 1 bf16" "celebaHQ 128 1 bf16" "mnist 4096 1 bf16" "cifar10 128 1 bf16"; do
  set -- $cfg
  timeout 900 ncu --metrics $M --clock-control none -c 200 --csv --log-file gpurun_out/z_launches_$1_$4_B$2.csv \
    python tools/profile_config.py $1 $2 $3 $4 > gpurun_out/z_ncu_$1_$2.log 2>&1
  echo "ncu $cfg exit $?"
done
PREC=bf16 timeout 900 python tools/bench_configs.py > gpurun_out/z_configs_bf16.log 2>&1
tail -1 gpurun_out/z_configs_bf16.log
PREC=tf32 timeout 900 python tools/bench_configs.py svhn cifar10 celebaHQ > gpurun_out/z_configs_tf32.log 2>&1
tail -1 gpurun_out/z_configs_tf32.log
timeout 600 python tools/bench_secondary.py > gpurun_out/z_secondary.log 2>&1
timeout 600 python tools/bench_denoiser_seq.py > gpurun_out/z_denoiser_hoisted.json 2> gpurun_out/z_denoiser_hoisted.err
tail -3 gpurun_out/z_denoiser_hoisted.json
for b in 128 16384; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:den_seq --csv --log-file gpurun_out/z_launches_denoiser_hoisted_B$b.csv python tools/profile_denseq.py $b > /dev/null 2>&1
  echo "ncu denoiser $b exit $?"
done
tail -3 gpurun_out/z_secondary.log
