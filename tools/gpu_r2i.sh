#!/bin/bash
# Round 2, call I: "pair" steps (fp16 activations x fp16-pair weights, two passes over K) for 16 <= t < 63 instead of the 3-term split:
# parity floors per weight set (new default vs all 63 split), loop time per format, sampling/training GPU tests.
TAG=${1:-r2i}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
for W in "" seed1 seed2 trained_like; do
    PARITY_FLOOR_WEIGHTS=$W PARITY_FLOOR_QUICK=1 timeout 150 python tools/parity_floor.py 64 63::63 63::16 63::8 125::16 > $OUT/${TAG}_parity_floor_${W:-seed0}.txt 2>&1
    echo "parity floor ${W:-seed0} rc=$? t=$SECONDS"; grep -E "tcgen05" $OUT/${TAG}_parity_floor_${W:-seed0}.txt | cut -c1-200
done
timeout 150 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop.txt 2>&1; echo "loop rc=$? t=$SECONDS"; cat $OUT/${TAG}_loop.txt
timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_training_oracle.py -m gpu -q -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log | tee $OUT/${TAG}_tests.log
grep -E "FAILED|^weight set|^zero-valued|^full size|full-tensor" $OUT/${TAG}_tests_full.log | cut -c1-260 | head -20
