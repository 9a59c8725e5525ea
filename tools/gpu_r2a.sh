#!/bin/bash
# Round 2, call A: GPU tests (precision-policy / checksum / weight-set goldens), zig-zag on/off in the loop, A-resident QKV,
# and ncu --set full captures of the HBM-bound kernels.
TAG=${1:-r2a}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -q -x -s > $OUT/${TAG}_tests_full.log 2>&1
echo "tests rc=$? t=$SECONDS"; tail -4 $OUT/${TAG}_tests_full.log | tee $OUT/${TAG}_tests.log
grep -E "^full size|^weight set|^zero-valued|FAILED|Error" $OUT/${TAG}_tests_full.log | head -20
timeout 120 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop_zigzag.txt 2>&1; echo "loop zigzag rc=$? t=$SECONDS"; cat $OUT/${TAG}_loop_zigzag.txt
EGOEGO_ZIGZAG=0 timeout 120 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop_nozigzag.txt 2>&1; echo "loop no-zigzag rc=$? t=$SECONDS"; cat $OUT/${TAG}_loop_nozigzag.txt
timeout 100 python tools/time_kernels.py 256 > $OUT/${TAG}_kernels.txt 2>&1; echo "kernels rc=$? t=$SECONDS"; cat $OUT/${TAG}_kernels.txt
EGOEGO_QKV_ARES=1 timeout 100 python tools/time_kernels.py 256 > $OUT/${TAG}_kernels_ares.txt 2>&1; echo "kernels ares rc=$? t=$SECONDS"; cat $OUT/${TAG}_kernels_ares.txt
for K in ddpm_update_kernel layernorm512_kernel postprocess_kernel fk_smpl_kernel canonicalize_head_kernel init_sample_kernel tail_condition_kernel; do
    timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 2 -f \
        -o $OUT/${TAG}_prof_$K python tools/prof_hbm_kernels.py 256 > $OUT/${TAG}_ncu_$K.log 2>&1
    echo "ncu full $K rc=$? t=$SECONDS"
done
# split-format start / out GEMMs (gemm_split3_2cta_kernel with TcEpiStart / TcEpiOut): launches of the split steps
PROF_STEPS=4 PROF_SPLIT_STEPS=4 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_split3_2cta_kernel.*TcEpiStart -s 1 -c 1 -f \
    -o $OUT/${TAG}_prof_split_start python tools/prof_hbm_kernels.py 256 > $OUT/${TAG}_ncu_split_start.log 2>&1; echo "ncu split start rc=$? t=$SECONDS"
PROF_STEPS=4 PROF_SPLIT_STEPS=4 timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_split3_2cta_kernel.*TcEpiOut -s 1 -c 1 -f \
    -o $OUT/${TAG}_prof_split_out python tools/prof_hbm_kernels.py 256 > $OUT/${TAG}_ncu_split_out.log 2>&1; echo "ncu split out rc=$? t=$SECONDS"
EGOEGO_QKV_ARES=1 PROF_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ares_tma_2cta_kernel -s 8 -c 1 -f \
    -o $OUT/${TAG}_prof_ares python tools/time_kernels.py 256 > $OUT/${TAG}_ncu_ares.log 2>&1; echo "ncu ares rc=$? t=$SECONDS"
ls -la $OUT | tail -30
