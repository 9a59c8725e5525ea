#!/bin/bash
# One short gpurun call at the end of a session: precision check (dithered weight sets on/off), GPU tests, bench line,
# and -- only if the call is still young -- the ncu launch list.   gpurun --timeout 400 -- 'bash tools/gpu_final.sh r1be'
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
PARITY_FLOOR_QUICK=1 timeout 100 python tools/parity_floor.py 64 63:8 63:1 63:16 32:8 125:8 > $OUT/${TAG}_parity_floor.txt 2>&1
echo "parity_floor rc=$? t=$SECONDS"; grep tcgen05 $OUT/${TAG}_parity_floor.txt | cut -c1-200
timeout 170 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log > $OUT/${TAG}_tests.log; cat $OUT/${TAG}_tests.log
grep -E "^full size|FAILED|Error" $OUT/${TAG}_tests_full.log | head -10
timeout 240 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$? t=$SECONDS"; python - <<PYEOF
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "roofline", d["roofline"]["frac"])
    print("parity", d.get("parity_vs_reference"))
    print({k: round(v["ms_per_launch"] * 1e3, 1) for k, v in d.get("kernels", {}).items()})
except Exception as e:
    print("bench parse failed", e)
PYEOF
if [ $SECONDS -lt ${NCU_DEADLINE:-175} ]; then
    EGOEGO_BENCH_SKIP_TORCH=1 timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
        python bench.py --steps 1 --warmup 1 --diffusion-steps 100 --cpu-seconds 1 > $OUT/${TAG}_ncu_bench.log 2>&1
    echo "launch list rc=$? t=$SECONDS"
fi
