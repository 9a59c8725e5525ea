#!/bin/bash
# Round 2, 8-GPU call: configs[2] (sharded sampling, one all-gather of SMPL parameters, shard-invariance check) and configs[4]
# (training step under DDP).
TAG=${1:-r2g8}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 \
    > $OUT/${TAG}_bench_8gpu.json 2> $OUT/${TAG}_bench_8gpu.err
echo "bench 8 rc=$? t=$SECONDS"; tail -2 $OUT/${TAG}_bench_8gpu.err | cut -c1-300; python - <<PYEOF
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_8gpu.json").read().strip().splitlines()[-1])
    print("value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"], "shard_check", d.get("shard_check"), "clocks", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
PYEOF
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/train_bench.py --steps 20 \
    > $OUT/${TAG}_train_ddp_8gpu.json 2> $OUT/${TAG}_train_ddp_8gpu.err
echo "train ddp rc=$? t=$SECONDS"; tail -2 $OUT/${TAG}_train_ddp_8gpu.err | cut -c1-300; cat $OUT/${TAG}_train_ddp_8gpu.json
