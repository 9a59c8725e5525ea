#!/bin/bash
# Round 2, call C: training tests after the refresh / tolerance fixes, cluster-of-8 multicast GEMM experiment, 64-window parity
# floor per weight set at K = 32 / 63, training bench.
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 300 python -m pytest tests/test_training_oracle.py -m gpu -q -s > $OUT/${TAG}_tests_train.log 2>&1
echo "train tests rc=$? t=$SECONDS"; tail -3 $OUT/${TAG}_tests_train.log; grep -E "full-tensor|fingerprint error" $OUT/${TAG}_tests_train.log | cut -c1-220
timeout 200 python tools/c8_check.py > $OUT/${TAG}_c8_check.txt 2>&1; echo "c8 check rc=$? t=$SECONDS"; cat $OUT/${TAG}_c8_check.txt
EGOEGO_GEMM_C8=1 timeout 120 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop_c8.txt 2>&1; echo "loop c8 rc=$? t=$SECONDS"; cat $OUT/${TAG}_loop_c8.txt
timeout 100 python tools/train_bench.py --steps 20 > $OUT/${TAG}_train_bench.json 2> $OUT/${TAG}_train_bench.err; echo "train bench rc=$? t=$SECONDS"; cat $OUT/${TAG}_train_bench.json; tail -2 $OUT/${TAG}_train_bench.err
for W in "" seed1 seed2 trained_like; do
    PARITY_FLOOR_WEIGHTS=$W PARITY_FLOOR_QUICK=1 timeout 120 python tools/parity_floor.py 64 32 63 > $OUT/${TAG}_parity_floor_${W:-seed0}.txt 2>&1
    echo "parity floor ${W:-seed0} rc=$? t=$SECONDS"; grep -E "tcgen05|weights" $OUT/${TAG}_parity_floor_${W:-seed0}.txt | cut -c1-200
done
