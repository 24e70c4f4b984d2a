#!/bin/bash
# 1-GPU evidence call: full GPU tests, bench (+cpu baseline), reference arm, ncu --set full of the
# TMA GEMM, ncu launch list of bench.py, secondary workloads.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?"; tail -1 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q --timeout=200 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench1.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench1.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
echo "bench ref exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_nt_tma -c 1 -f -o gpurun_out/gemm8192_tma python tools/prof_gpr.py --what gemm --n 8192 > gpurun_out/prof_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu bench exit $?"
python tools/summarise_launches.py gpurun_out/launches_bench.csv > gpurun_out/launches_bench_summary.txt 2>&1; head -10 gpurun_out/launches_bench_summary.txt
bash tools/gpu_secondary.sh
