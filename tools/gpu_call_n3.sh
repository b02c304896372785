#!/bin/bash
# slim multi-GPU call: sharded parity (small shards: the small-cloud tile rule is active on the ranks) + our bench arm
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29711 tools/sharded_check.py > gpurun_out/sharded_$N.log 2>&1; echo "sharded_check rc=$?"
grep -E "mode|SHARDED|Error|error" gpurun_out/sharded_$N.log | head -20
timeout 400 $TR --master-port 29713 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['pose_err'])
for k in ('sharded_icp','reduction','trials'):
    if k in d: print('  ',k, json.dumps(d[k])[:500])
"
