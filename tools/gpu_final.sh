#!/bin/bash
# final 1-GPU verification: build check, smoke, full GPU test suite, both bench arms as the driver runs them
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/ref_final.json 2> gpurun_out/ref_final.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_final.err
python -c "
import json
r=json.loads(open('gpurun_out/ref_final.json').read().strip().splitlines()[-1]); b=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['cores'], r['config']==b['config'])
print('ours', b['value'], b['e2e']['value'], b['roofline']['frac'], b['roofline']['batch_ms'], b['pose_err'], b['launches_per_step']['kernels'])
print('ratio e2e', b['e2e']['value']/r['value'])
"
