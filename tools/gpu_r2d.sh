#!/bin/bash
# Round 2, call D: full (fp64-arbitrated) parity floors on the new weight sets, training tests.
TAG=${1:-r2d}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 300 python -m pytest tests/test_training_oracle.py -m gpu -q -s > $OUT/${TAG}_tests_train.log 2>&1
echo "train tests rc=$? t=$SECONDS"; tail -2 $OUT/${TAG}_tests_train.log
for W in trained_like seed2 seed1; do
    PARITY_FLOOR_WEIGHTS=$W timeout 400 python tools/parity_floor.py 64 1000 63 32 > $OUT/${TAG}_parity_floor_full_$W.txt 2>&1
    echo "parity floor full $W rc=$? t=$SECONDS"; grep -vE "^torch fp32 done" $OUT/${TAG}_parity_floor_full_$W.txt | cut -c1-330
done
