#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k gram_backward --timeout=150 -p no:cacheprovider > gpurun_out/pytest_gram.log 2>&1
tail -5 gpurun_out/pytest_gram.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_gpr4096.csv python tools/prof_gpr.py --what gpr --n 4096 > gpurun_out/prof_gpr.log 2>&1
echo "ncu gpr exit $?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_nt_dmma -c 1 -f -o gpurun_out/gemm8192 python tools/prof_gpr.py --what gemm --n 8192 > gpurun_out/prof_gemm.log 2>&1
echo "ncu gemm exit $?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:potrf_base -c 1 -f -o gpurun_out/base python tools/prof_gpr.py --what potrf --n 1024 > gpurun_out/prof_base.log 2>&1
echo "ncu base exit $?"
timeout 600 python bench.py > gpurun_out/bench1.log 2>&1
echo "bench exit $?"; tail -3 gpurun_out/bench1.log
