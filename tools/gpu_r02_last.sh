#!/bin/bash
# Last GPU call of round 2: the shipped (one tile per CTA) TMA kernel again -- suite, ncu capture, bench.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02n_gpu_tests.log 2>&1; tail -2 gpurun_out/r02n_gpu_tests.log
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_nt_tma -c 1 -f -o gpurun_out/r02_gemm8192_tma_full python tools/prof_gpr.py --what gemm --n 8192 > gpurun_out/r02n_prof_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 400 python bench.py > gpurun_out/r02_bench_n32768_1gpu.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02_bench_n32768_1gpu.json') if l.startswith('{')][-1])
    print('1 GPU: %.4f evals/s e2e %.4f roofline %.3f parity %.1e potrf %.1f TF' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_rel_err'], d['potrf']['tflops']))
    for k, v in d['secondary'].items(): print(' ', k, round(v['value'], 2), v['metric'], 'frac', round(v['roofline']['frac'], 3))
except Exception as e: print('no line', e)
PY
