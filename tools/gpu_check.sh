#!/bin/bash
# One gpurun call: GPU tests, smoke, vendor/own FP64 measurements.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -rA 2>&1 | tail -80 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python tools/measure_fp64.py > gpurun_out/measure.log 2>&1
echo "measure exit: $?" >> gpurun_out/measure.log
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -5; tail -40 gpurun_out/measure.log
