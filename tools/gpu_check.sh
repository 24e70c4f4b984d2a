#!/bin/bash
# One gpurun call: smoke first, then GPU tests (per-test timeout), then vendor/own FP64
# measurements.  Logs are written incrementally into gpurun_out/ so a hang still leaves evidence.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 240 python -X faulthandler -c "import faulthandler, sys; faulthandler.dump_traceback_later(150, exit=True); import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
rc=$?; echo "smoke exit: $rc" >> gpurun_out/smoke.log
tail -15 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed; running kernel tests only for diagnostics"; fi
timeout ${PYTEST_LIMIT:-700} python -m pytest tests -m gpu -q -rA --timeout=150 -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|PASSED|FAILED|ERROR|Error|error|worst|exit" gpurun_out/pytest_gpu.log | tail -60
if [ "${SKIP_MEASURE}" != "1" ]; then
timeout 400 python tools/measure_fp64.py > gpurun_out/measure.log 2>&1
echo "measure exit: $?" >> gpurun_out/measure.log
tail -45 gpurun_out/measure.log
fi
