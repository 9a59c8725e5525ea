#!/bin/bash
# Round 2, call M: mid-point validation after the pair steps / ddpm-advance fold: full GPU tests, smoke, bench line, loop time,
# launch list of the training step (where its 7 ms go).
TAG=${1:-r2m}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 120 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop.txt 2>&1; echo "loop rc=$? t=$SECONDS"; cat $OUT/${TAG}_loop.txt
for B in 1 32; do LOOP_ONLY=fp16,default timeout 120 python tools/loop_time.py $B 1000 > $OUT/${TAG}_loop_b$B.txt 2>&1; cat $OUT/${TAG}_loop_b$B.txt; done
timeout 800 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log | tee $OUT/${TAG}_tests.log
grep -E "FAILED|^weight set|^zero-valued|^full size|full-tensor" $OUT/${TAG}_tests_full.log | cut -c1-260 | head -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$? t=$SECONDS"; tail -3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$? t=$SECONDS"; tail -3 $OUT/${TAG}_bench.err; python - <<PYEOF
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "all_split", d.get("all_split_windows_per_s"), "clocks", d["clocks"], "roofline", d["roofline"]["frac"])
    print(d["dtype"]); print(d["engine_info"])
    print({k: round(v["ms_per_launch"] * 1e3, 1) for k, v in d.get("kernels", {}).items()})
    print("latency", d.get("single_window_latency")); print("pipeline", d.get("pipeline_config4"))
except Exception as e:
    print("bench parse failed", e)
PYEOF
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 800 --csv --log-file $OUT/${TAG}_train_launches.csv \
    python tools/train_bench.py --steps 4 --warmup 5 --arms ours > $OUT/${TAG}_ncu_train.log 2>&1
echo "train launch list rc=$? t=$SECONDS"
python tools/summarize_launches.py $OUT/${TAG}_train_launches.csv "training step launch list (B=32)" > $OUT/${TAG}_train_launches.md; head -45 $OUT/${TAG}_train_launches.md
