#!/bin/bash
# round 2, call H (2 GPUs): fused optimiser tests incl. the 2-GPU NCCL one; training-iteration timings at 1 and 2 GPUs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_optimizer.py tests/test_gpu_training_step.py -m gpu -q -x -p no:cacheprovider > gpurun_out/h_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/h_pytest.log
tail -15 gpurun_out/h_pytest.log
timeout 600 python tools/bench_train_iter.py > gpurun_out/h_train_1gpu.json 2> gpurun_out/h_train_1gpu.err
cat gpurun_out/h_train_1gpu.json; tail -3 gpurun_out/h_train_1gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_train_iter.py > gpurun_out/h_train_2gpu.json 2> gpurun_out/h_train_2gpu.err
cat gpurun_out/h_train_2gpu.json; tail -3 gpurun_out/h_train_2gpu.err
