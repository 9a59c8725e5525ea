#!/bin/bash
# Round 2, call B: GPU tests (dropout training step, floor height, tensor-core ResNet, cleaned GEMM header), full bench line,
# ResNet timing, policy margin scan, remaining ncu captures.
TAG=${1:-r2b}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 700 python -m pytest tests -m gpu -q -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log | tee $OUT/${TAG}_tests.log
grep -E "FAILED|Error|^\[.*drop\]|ResNet-18 features|raw-flow|fused-LN" $OUT/${TAG}_tests_full.log | head -30
timeout 100 python tools/time_resnet.py > $OUT/${TAG}_resnet.txt 2>&1; echo "resnet rc=$? t=$SECONDS"; cat $OUT/${TAG}_resnet.txt
timeout 120 python tools/precise_scan_ws.py 32 48 63 > $OUT/${TAG}_precise_scan_ws.txt 2>&1; echo "scan rc=$? t=$SECONDS"; cat $OUT/${TAG}_precise_scan_ws.txt
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$? t=$SECONDS"; tail -3 $OUT/${TAG}_bench.err; python - <<PYEOF
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "all_split", d.get("all_split_windows_per_s"), "clocks", d["clocks"], "roofline", d["roofline"]["frac"])
    print({k: round(v["ms_per_launch"] * 1e3, 1) for k, v in d.get("kernels", {}).items()})
    nr = d.get("next_rows", {})
    print("train", nr.get("train_step")); print("resnet", nr.get("resnet18_encoder")); print("stage1", nr.get("stage1"))
    print("pipeline", d.get("pipeline_config4")); print("once", d.get("once_per_sample_kernels")); print("torch", d.get("torch_gpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PYEOF
timeout 200 ncu --set full --clock-control none --import-source on -k regex:init_sample_kernel -c 1 -f \
    -o $OUT/${TAG}_prof_init_sample_kernel python tools/prof_hbm_kernels.py 256 > $OUT/${TAG}_ncu_init_sample.log 2>&1; echo "ncu init rc=$? t=$SECONDS"
PROF_STEPS=4 PROF_SPLIT_STEPS=4 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_split3_2cta_kernel -s 1 -c 1 -f \
    -o $OUT/${TAG}_prof_split_start python tools/prof_hbm_kernels.py 256 > $OUT/${TAG}_ncu_split_start.log 2>&1; echo "ncu split start rc=$? t=$SECONDS"
PROF_STEPS=4 PROF_SPLIT_STEPS=4 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_split3_2cta_kernel -s 18 -c 1 -f \
    -o $OUT/${TAG}_prof_split_out python tools/prof_hbm_kernels.py 256 > $OUT/${TAG}_ncu_split_out.log 2>&1; echo "ncu split out rc=$? t=$SECONDS"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 2 -c 3 -f \
    -o $OUT/${TAG}_prof_conv_tc python tools/time_resnet.py > $OUT/${TAG}_ncu_conv_tc.log 2>&1; echo "ncu conv_tc rc=$? t=$SECONDS"
ls -la $OUT | grep ${TAG} | tail -30
