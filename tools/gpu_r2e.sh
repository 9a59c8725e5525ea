#!/bin/bash
# Round 2, call E: dual-accumulator split GEMMs / attention S -- primitive errors, parity floors, split-step cost, parity tests.
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 100 python tools/split_probe.py > $OUT/${TAG}_split_probe.txt 2>&1; echo "split probe rc=$? t=$SECONDS"; cat $OUT/${TAG}_split_probe.txt
PARITY_FLOOR_WEIGHTS=trained_like timeout 400 python tools/parity_floor.py 64 1000 63 > $OUT/${TAG}_parity_floor_full_trained_like.txt 2>&1
echo "parity floor full trained_like rc=$? t=$SECONDS"; grep -E "tcgen05|simt" $OUT/${TAG}_parity_floor_full_trained_like.txt | cut -c1-330
for W in "" seed1 seed2; do
    PARITY_FLOOR_WEIGHTS=$W PARITY_FLOOR_QUICK=1 timeout 120 python tools/parity_floor.py 64 1000 63 > $OUT/${TAG}_parity_floor_${W:-seed0}.txt 2>&1
    echo "parity floor ${W:-seed0} rc=$? t=$SECONDS"; grep -E "tcgen05" $OUT/${TAG}_parity_floor_${W:-seed0}.txt | cut -c1-200
done
timeout 120 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop.txt 2>&1; echo "loop rc=$? t=$SECONDS"; cat $OUT/${TAG}_loop.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py tests/test_training_oracle.py -m gpu -q -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log | tee $OUT/${TAG}_tests.log
grep -E "FAILED|^weight set|^zero-valued|^full size|full-tensor" $OUT/${TAG}_tests_full.log | cut -c1-260 | head -20
