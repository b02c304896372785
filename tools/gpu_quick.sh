#!/bin/bash
# quick loop check: parity of the benchmarked step + loop timing + timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "c2_as_benchmarked or lattice or reused" > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lattice or reused or fixed_iterations" >> gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
VARIANT=default timeout 200 python tools/loop_variants.py > gpurun_out/variants.log 2>&1; head -3 gpurun_out/variants.log | cut -c1-500
timeout 300 python tools/timeline.py > gpurun_out/timeline.log 2>&1; grep -E "^iteration|per-phase|last block|solve step" gpurun_out/timeline.log | cut -c1-250
