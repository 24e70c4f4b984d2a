#!/bin/bash
# Round-2 last call (1 GPU, 2 GPU-minutes left): the GPU suite on the final code, the bench line (secondary
# configs included, CPU baseline skipped for time: the complete line of the same day is r02p), and -- if the
# budget still allows -- one ncu --set full capture of the tuned NKN backward kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 100 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02q_gpu_tests.log 2>&1; tail -2 gpurun_out/r02q_gpu_tests.log; grep "gram_impl=" gpurun_out/r02q_gpu_tests.log
timeout 100 python bench.py --no-cpu-baseline > gpurun_out/r02q_bench_n32768_1gpu.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02q_bench_n32768_1gpu.json') if l.startswith('{')][-1])
    print('1 GPU: %.4f evals/s e2e %.4f roofline %.3f parity %.1e potrf %.1f TF' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_rel_err'], d['potrf']['tflops']))
    for k, v in d['secondary'].items(): print(' ', k, round(v['value'], 2), v['metric'], 'frac', round(v['roofline']['frac'], 3))
except Exception as e: print('no line', e)
PY
timeout 40 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gram_bwd_nkn -c 1 -f -o gpurun_out/r02_nkn_bwd_tuned_full python tools/prof_gpr.py --what nkn --n 4096 > gpurun_out/r02q_ncu_nkn.log 2>&1; echo "ncu rc=$?"
