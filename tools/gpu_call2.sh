#!/bin/bash
# GPU test suite, both bench arms, the per-iteration timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 300 python tools/timeline.py > gpurun_out/timeline.log 2>&1; cat gpurun_out/timeline.log | cut -c1-330
for v in default NO_GRAPH; do
  if [ $v = default ]; then VARIANT=$v timeout 200 python tools/loop_variants.py; else env VARIANT=$v DCREG_$v=1 timeout 200 python tools/loop_variants.py; fi
done > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log | cut -c1-700
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/ref_a.json 2> gpurun_out/ref_a.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_a.err
python -c "
import json
for f in ('ref_a','bench_a'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}) and d['cpu_baseline'].get('cores'), d.get('roofline',{}).get('frac'), d.get('trials',{}).get('cpu_port'))
    except Exception as e: print(f, 'ERR', e)
"
