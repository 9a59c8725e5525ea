#!/bin/bash
# Round 2, call J: streamed attention (QKV projection and attention concurrently, per-window counters): bit-identity test, loop time
# for several SM partitions / balance ratios, small-batch step times, then the sampling tests.
TAG=${1:-r2j}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "streamed_attention" > $OUT/${TAG}_stream_test.log 2>&1
echo "stream test rc=$? t=$SECONDS"; tail -5 $OUT/${TAG}_stream_test.log
for PR in 0:0 58:0.96 62:0.96 62:0.85 62:1.1 66:0.96; do
    P=${PR%%:*}; R=${PR##*:}
    if [ $P = 0 ]; then export EGOEGO_STREAM_ATT=0; else unset EGOEGO_STREAM_ATT; export EGOEGO_STREAM_ATT_PAIRS=$P EGOEGO_STREAM_ATT_RATIO=$R; fi
    LOOP_ONLY=fp16 timeout 120 python tools/loop_time.py 256 400 > $OUT/${TAG}_loop_p${P}_r$R.txt 2>&1; echo "loop P=$P R=$R rc=$? t=$SECONDS"; grep -E "^fp16|Error|error" $OUT/${TAG}_loop_p${P}_r$R.txt | head -3
done
for B in 1 32; do
    for SA in 0 1; do
        unset EGOEGO_STREAM_ATT_PAIRS EGOEGO_STREAM_ATT_RATIO; export EGOEGO_STREAM_ATT=$SA
        LOOP_ONLY=fp16 timeout 120 python tools/loop_time.py $B 400 > $OUT/${TAG}_loop_b${B}_sa$SA.txt 2>&1; echo "B=$B stream=$SA"; grep -E "^fp16|rror" $OUT/${TAG}_loop_b${B}_sa$SA.txt | head -2
    done
done
unset EGOEGO_STREAM_ATT EGOEGO_STREAM_ATT_PAIRS EGOEGO_STREAM_ATT_RATIO
if [ "$2" = "tests" ]; then
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log | tee $OUT/${TAG}_tests.log
fi
