#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -m gpu -q -x --timeout=150 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_gpr4096.csv python tools/prof_gpr.py --what gpr --n 4096 > gpurun_out/prof_gpr.log 2>&1
echo "ncu gpr exit $?"
timeout 300 python tools/measure_fp64.py > gpurun_out/measure.log 2>&1; grep -E "ours_potrf|cusolver_potrf_.*_ms|ours_trsm" gpurun_out/measure.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.log 2>&1
echo "bench exit $?"; tail -2 gpurun_out/bench1.log
