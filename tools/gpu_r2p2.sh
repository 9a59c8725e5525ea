#!/bin/bash
# Round 2, 2-GPU call: gradient averaging check (DDP and set_grad_sync), graph-replay test.
TAG=${1:-r2p2}
OUT=gpurun_out
mkdir -p $OUT
export TQDM_DISABLE=1
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || { tail -20 $OUT/${TAG}_build.log; exit 1; }
timeout 200 python -m pytest tests/test_training_oracle.py -m gpu -q -s -k "graph_replay" > $OUT/${TAG}_replay_test.log 2>&1; echo "replay test rc=$?"; tail -4 $OUT/${TAG}_replay_test.log | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/train_ddp_check.py > $OUT/${TAG}_ddp_check.txt 2>&1
echo "check rc=$? t=$SECONDS"; grep -E "^rank|Error|error" $OUT/${TAG}_ddp_check.txt | head
