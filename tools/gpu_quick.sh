#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
timeout 240 python tools/gemm_check.py > gpurun_out/gemm_check.log 2>&1
echo "gemm_check exit $?"; tail -9 gpurun_out/gemm_check.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout=200 -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?"; tail -3 gpurun_out/pytest_gpu.log
if [ "$FULL" = "1" ]; then
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench1.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench1.log
fi
