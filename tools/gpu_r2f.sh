#!/bin/bash
# Round 2, call F: dual accumulators on the last 16 steps only -- parity floors per weight set, loop time; full GPU tests; bench.
TAG=${1:-r2f}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
for W in "" seed1 seed2 trained_like; do
    PARITY_FLOOR_WEIGHTS=$W PARITY_FLOOR_QUICK=1 timeout 120 python tools/parity_floor.py 64 63 48 > $OUT/${TAG}_parity_floor_${W:-seed0}.txt 2>&1
    echo "parity floor ${W:-seed0} rc=$? t=$SECONDS"; grep -E "tcgen05" $OUT/${TAG}_parity_floor_${W:-seed0}.txt | cut -c1-200
done
timeout 120 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop.txt 2>&1; echo "loop rc=$? t=$SECONDS"; cat $OUT/${TAG}_loop.txt
timeout 700 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log | tee $OUT/${TAG}_tests.log
grep -E "FAILED|^weight set|^zero-valued|^full size|full-tensor" $OUT/${TAG}_tests_full.log | cut -c1-260 | head -20
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$? t=$SECONDS"; tail -3 $OUT/${TAG}_bench.err; python - <<PYEOF
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "all_split", d.get("all_split_windows_per_s"), "clocks", d["clocks"], "roofline", d["roofline"]["frac"])
    print({k: round(v["ms_per_launch"] * 1e3, 1) for k, v in d.get("kernels", {}).items()})
    nr = d.get("next_rows", {})
    print("train", nr.get("train_step")); print("stage1", nr.get("stage1"))
except Exception as e:
    print("bench parse failed", e)
PYEOF
