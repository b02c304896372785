#!/bin/bash
# multi-GPU call: sharded parity in both exchange modes, bench at N ranks (both arms)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29711 tools/sharded_check.py > gpurun_out/sharded_$N.log 2>&1; echo "sharded_check rc=$?"
grep -E "mode|SHARDED|Error|error" gpurun_out/sharded_$N.log | head -20
SHARDED_CHECK_POINTS=1000000 timeout 600 $TR --master-port 29712 tools/sharded_check.py > gpurun_out/sharded_1m_$N.log 2>&1; echo "sharded_check 1M rc=$?"
grep -E "mode|SHARDED|Error|error" gpurun_out/sharded_1m_$N.log | head -20
timeout 600 $TR --master-port 29715 tools/sharded_reduce_time.py > gpurun_out/sharded_reduce_$N.log 2>&1; grep -E "mode|exchange" gpurun_out/sharded_reduce_$N.log
timeout 600 $TR --master-port 29713 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_n$N.err
timeout 300 $TR --master-port 29714 bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/ref_n$N.json 2> gpurun_out/ref_n$N.err; echo "ref rc=$?"
python -c "
import json
for f in ('ref_n$N','bench_n$N'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d.get('e2e',{}).get('value'), (d.get('cpu_baseline') or {}).get('cores'))
        for k in ('sharded_icp','reduction','trials'):
            if k in d: print('  ',k, json.dumps(d[k])[:600])
    except Exception as e: print(f, 'ERR', e)
"
