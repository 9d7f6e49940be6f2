#!/bin/bash
# round 2, call K (8 GPUs): 2-GPU NCCL optimiser test, training-iteration timings at 4 and 8 GPUs, sampling bench at 8 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_optimizer.py -m gpu -q -x -p no:cacheprovider > gpurun_out/k_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/k_pytest.log
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/bench_train_iter.py > gpurun_out/k_train_${n}gpu.json 2> gpurun_out/k_train_${n}gpu.err
  cat gpurun_out/k_train_${n}gpu.json
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/k_bench_8gpu.json 2> gpurun_out/k_bench_8gpu.err
cat gpurun_out/k_bench_8gpu.json
