#!/bin/bash
# round 2, call Q: Q.calculate_loss on the library GEMMs: parity test + training-iteration timing with both engines
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training_step.py -m gpu -q -x -s -p no:cacheprovider > gpurun_out/q2_pytest.log 2>&1
echo "pytest exit $?"; grep -E "^\.?F?B=|passed|failed" gpurun_out/q2_pytest.log | tail -4
for e in torch library library_graphed; do
  Q_ENGINE=$e timeout 600 python tools/bench_train_iter.py > gpurun_out/q2_train_$e.json 2> gpurun_out/q2_train_$e.err
  cat gpurun_out/q2_train_$e.json; tail -2 gpurun_out/q2_train_$e.err
done
