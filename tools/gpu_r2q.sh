#!/bin/bash
# Round 2, call Q: ncu --set full of the x-path kernels of an fp16 step inside the loop (start, linear_out, DDPM update)
TAG=${1:-r2q}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
for K in "gemm_split3_2cta_kernel"; do
    N=$(echo $K | tr -cd 'A-Za-z_')
    PROF_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -s 9 -c 2 -f \
        -o $OUT/${TAG}_prof_$N python tools/time_kernels.py 256 > $OUT/${TAG}_ncu_$N.log 2>&1
    echo "ncu full $K rc=$? t=$SECONDS"
done
ls -la $OUT | grep ${TAG}
