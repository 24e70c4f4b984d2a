#!/bin/bash
# Session-3 call A (1 GPU): smoke, GPU parity tests, world-1 distributed phases, bench, ncu.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" | tee -a gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -rA --timeout=200 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|ERROR|Error" gpurun_out/pytest_gpu.log | tail -20
timeout 300 python tools/dist_time.py --size 16384 > gpurun_out/dist_time_w1_16k.log 2>&1
echo "dist_time exit $?"; tail -5 gpurun_out/dist_time_w1_16k.log
timeout 900 python bench.py > gpurun_out/bench1.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench1.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_nt_dmma -c 1 -f -o gpurun_out/gemm8192 python tools/prof_gpr.py --what gemm --n 8192 > gpurun_out/prof_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu bench exit $?"
python tools/summarise_launches.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.txt 2>&1; head -12 gpurun_out/launches_bench_summary.txt
