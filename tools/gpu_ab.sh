#!/bin/bash
# A/B timing of opt-in/opt-out kernel variants in one gpurun call:  bash tools/gpu_ab.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...
# (first argument = tag; every further argument is one environment setting string, "" = defaults)
TAG=${1:-ab}; shift
OUT=gpurun_out; mkdir -p $OUT
export TQDM_DISABLE=1 EGOEGO_BENCH_SKIP_TORCH=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
for CFG in "$@"; do
    echo "=== [$CFG] kernels"
    env $CFG timeout 180 python tools/time_kernels.py 256 20 2>&1 | tail -3
done
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -4 $OUT/${TAG}_tests.log
for CFG in "$@"; do
    echo "=== [$CFG] bench"
    env $CFG timeout 300 python bench.py --steps 2 --warmup 3 --cpu-seconds 1 2> $OUT/${TAG}_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'sm_mhz', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'], 'jerr_mm', d['parity_vs_reference'].get('joint_max_abs_mm'))
print({k: round(v['ms_per_launch']*1e3,1) for k,v in d['kernels'].items()})"
done
