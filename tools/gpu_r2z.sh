#!/bin/bash
# Round 2, call Z (final validation): full GPU tests, smoke, bench line, launch list of the loop, ncu --set full of the dominant kernel
# and of the attention / fused-LN kernels in the zig-zag loop, stage-1 timings.
TAG=${1:-r2z}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 700 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log | tee $OUT/${TAG}_tests.log
grep -E "FAILED|^weight set|^zero-valued|^full size|full-tensor" $OUT/${TAG}_tests_full.log | cut -c1-260 | head -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$? t=$SECONDS"; tail -3 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$? t=$SECONDS"; tail -3 $OUT/${TAG}_bench.err; python - <<PYEOF
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "all_split", d.get("all_split_windows_per_s"), "clocks", d["clocks"], "roofline", d["roofline"]["frac"])
    print({k: round(v["ms_per_launch"] * 1e3, 1) for k, v in d.get("kernels", {}).items()})
    print("stage1", d.get("next_rows", {}).get("stage1"))
except Exception as e:
    print("bench parse failed", e)
PYEOF
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; echo "reference arm rc=$? t=$SECONDS"
# launch list of one whole 100-step loop of the bench workload: 37 fp16 steps, 47 pair steps, 16 split steps (dual accumulators)
EGOEGO_BENCH_SKIP_TORCH=1 EGOEGO_BENCH_LIGHT=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2433 -c 2433 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --diffusion-steps 100 --cpu-seconds 1 > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$? t=$SECONDS"
timeout 200 python tools/train_bench.py --steps 30 --warmup 8 > $OUT/${TAG}_train_bench.json 2> $OUT/${TAG}_train_bench.err; echo "train bench rc=$? t=$SECONDS"; cat $OUT/${TAG}_train_bench.json
timeout 120 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop.txt 2>&1; cat $OUT/${TAG}_loop.txt
for K in gemm_half_tma_2cta_kernel; do
    PROF_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 2 -f \
        -o $OUT/${TAG}_prof_$K python tools/time_kernels.py 256 > $OUT/${TAG}_ncu_$K.log 2>&1
    echo "ncu full $K rc=$? t=$SECONDS"
done
ls -la $OUT | grep ${TAG} | tail -20
