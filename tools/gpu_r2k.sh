#!/bin/bash
# Round 2, call K: streamed attention without proxy fences; per-CTA timelines; loop times
TAG=${1:-r2k}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
for PR in 0:0 62:0.96 58:0.96; do
    P=${PR%%:*}; R=${PR##*:}
    if [ $P = 0 ]; then export EGOEGO_STREAM_ATT=0; else unset EGOEGO_STREAM_ATT; export EGOEGO_STREAM_ATT_PAIRS=$P EGOEGO_STREAM_ATT_RATIO=$R; fi
    timeout 120 python tools/stream_timeline.py 256 24 > $OUT/${TAG}_timeline_p${P}.txt 2>&1; echo "timeline rc=$?"; tail -6 $OUT/${TAG}_timeline_p${P}.txt | cut -c1-1400
done
