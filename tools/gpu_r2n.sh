#!/bin/bash
# Round 2, call N: training step with split-K weight-gradient products, CUDA-graph capture and one-launch parameter / gradient copies:
# gradient tests, then the step time (tools/train_bench.py, batch 32, autocast + GradScaler + Adam, dropout on) beside stock PyTorch.
TAG=${1:-r2n}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 600 python -m pytest tests/test_training_oracle.py tests/test_gpu_parity.py tests/test_reference_trainer.py -m gpu -q -s -k "training or train or data_mutation or stale or refresh" > $OUT/${TAG}_tests_train.log 2>&1
echo "train tests rc=$? t=$SECONDS"; tail -3 $OUT/${TAG}_tests_train.log; grep -E "full-tensor|FAILED|Error" $OUT/${TAG}_tests_train.log | cut -c1-220 | head
timeout 200 python tools/train_bench.py --steps 30 --warmup 8 > $OUT/${TAG}_train_bench.json 2> $OUT/${TAG}_train_bench.err; echo "both arms rc=$?"; cat $OUT/${TAG}_train_bench.json; tail -3 $OUT/${TAG}_train_bench.err
