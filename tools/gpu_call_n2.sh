#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 tools/sharded_check.py > gpurun_out/sharded_$N.log 2>&1; echo "sharded_check rc=$?"
grep -E "mode|SHARDED|Error|error" gpurun_out/sharded_$N.log | head -20
timeout 600 $TR --master-port 29715 tools/sharded_reduce_time.py > gpurun_out/sharded_reduce_$N.log 2>&1; grep -E "mode|exchange" gpurun_out/sharded_reduce_$N.log
