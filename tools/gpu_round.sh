#!/bin/bash
# One gpurun call: smoke, GPU parity tests, FP64 measurements, bench (both arms), ncu launch list.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" | tee -a gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout ${PYTEST_LIMIT:-900} python -m pytest tests -m gpu -q -rA --timeout=200 -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|Error|worst" gpurun_out/pytest_gpu.log | tail -40
if [ "${SKIP_MEASURE}" != "1" ]; then
timeout 400 python tools/measure_fp64.py > gpurun_out/measure.log 2>&1
echo "measure exit: $?" >> gpurun_out/measure.log
grep -E "ours_|cusolver_potrf_.*_ms|exit" gpurun_out/measure.log | tail -30
fi
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench1.log 2>&1
echo "bench exit $?"; tail -2 gpurun_out/bench1.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "bench ref exit $?"; tail -1 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_gpr8192.csv python tools/prof_gpr.py --what gpr --n 8192 > gpurun_out/prof_gpr.log 2>&1
echo "ncu gpr exit $?"
