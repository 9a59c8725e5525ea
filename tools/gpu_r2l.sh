#!/bin/bash
# Round 2, call L: launch lists (ncu gpu__time_duration) of the fp16 loop at B = 1 and B = 32: how much of a small-batch step is
# kernel time and how much is gaps between the 24 graph nodes.
TAG=${1:-r2l}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
for B in 1 32; do
    LOOP_ONLY=fp16 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 96 --csv --log-file $OUT/${TAG}_launches_b$B.csv \
        python tools/loop_time.py $B 30 > $OUT/${TAG}_ncu_b$B.log 2>&1
    echo "B=$B rc=$? t=$SECONDS"
    python tools/summarize_launches.py $OUT/${TAG}_launches_b$B.csv "fp16 loop B=$B" | head -20
done
