#!/bin/bash
# Round 2, call R: L2 prefetch of the next tile's epilogue operand in the CTA-pair GEMM (start_conv's `base`, the residual of the split steps)
TAG=${1:-r2r}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 200 python tools/time_kernels.py 256 > $OUT/${TAG}_kernels.txt 2>&1; cat $OUT/${TAG}_kernels.txt
timeout 150 python tools/loop_time.py 256 1000 > $OUT/${TAG}_loop.txt 2>&1; cat $OUT/${TAG}_loop.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py tests/test_training_oracle.py -m gpu -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
