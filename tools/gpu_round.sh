#!/bin/bash
# One gpurun call = GPU tests + bench line + ncu launch list (+ optional full captures).
#   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh r1q [full]'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
TAG=${1:-run}
FULL=${2:-}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; python - <<EOF
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"], "roofline", d["roofline"]["frac"], d["roofline"]["kernel"])
    print({k: round(v["ms_per_launch"] * 1e3, 1) for k, v in d.get("kernels", {}).items()})
except Exception as e:
    print("bench parse failed", e)
EOF
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
# launch list of the same command (short run: 60-step schedule so both formats appear; shares, not absolutes)
EGOEGO_BENCH_SKIP_TORCH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --diffusion-steps 100 --cpu-seconds 1 > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
# precision of the default policy at scale (64 windows x 1000 steps vs PyTorch fp32 of the reference op sequence on one noise tape)
PARITY_FLOOR_QUICK=1 timeout 120 python tools/parity_floor.py 64 1000 63 > $OUT/${TAG}_parity_floor.txt 2>&1
echo "parity floor rc=$?"; grep tcgen05 $OUT/${TAG}_parity_floor.txt | cut -c1-170
if [ -n "$FULL" ]; then
    for K in gemm_half_tma_2cta_kernel attention_half_kernel gemm_ln_half_c4_kernel; do
        PROF_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 8 -c 2 -f \
            -o $OUT/${TAG}_prof_$K python tools/time_kernels.py 256 > $OUT/${TAG}_ncu_$K.log 2>&1
        echo "ncu full $K rc=$?"
    done
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_metrics -s 3 -c 1 -f \
        -o $OUT/${TAG}_prof_eval_metrics python tools/time_metrics.py > $OUT/${TAG}_ncu_eval_metrics.log 2>&1
    echo "ncu full eval_metrics rc=$?"
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 24 -c 3 -f \
        -o $OUT/${TAG}_prof_conv_igemm python tools/time_resnet.py > $OUT/${TAG}_ncu_conv_igemm.log 2>&1
    echo "ncu full conv_igemm rc=$?"
fi
ls -la $OUT | tail -20
