#!/bin/bash
# round-2 profile capture (one GPU): launch list of the bench command + ncu --set full of the loop kernel (lean / reuse) and K1
mkdir -p gpurun_out
export DCREG_NO_GRAPH=1
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_r2.json 2> gpurun_out/bench_under_ncu_r2.err; echo "launch list rc=$?"
ICP_ITERS=30 timeout 600 $NCU --set full --import-source on -k regex:icp_iter2_kernel -s 3 -c 1 -o gpurun_out/iter2_lean_r2 -f python tools/prof_icp.py > gpurun_out/prof_lean_r2.log 2>&1; echo "lean rc=$?"
ICP_ITERS=30 timeout 600 $NCU --set full --import-source on -k regex:icp_iter2_kernel -s 27 -c 1 -o gpurun_out/iter2_late_r2 -f python tools/prof_icp.py > gpurun_out/prof_late_r2.log 2>&1; echo "late rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:reduce_stream -s 8 -c 2 -o gpurun_out/k1_r2 -f python tools/prof_k1.py > gpurun_out/prof_k1_r2.log 2>&1; echo "k1 rc=$?"
unset DCREG_NO_GRAPH
K1_REPS=20 timeout 300 python tools/prof_k1.py > gpurun_out/k1_r2_times.log 2>&1; tail -8 gpurun_out/k1_r2_times.log
ls -la gpurun_out/*.ncu-rep | tail -5
