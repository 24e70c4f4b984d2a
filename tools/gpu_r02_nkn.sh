#!/bin/bash
# Round-2 closing call (1 GPU): whole GPU suite (prints the NKN Gram timings of the three implementations),
# the bench line with every secondary config, one ncu --set full capture of the NKN backward kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r02p_gpu_tests.log 2>&1; tail -2 gpurun_out/r02p_gpu_tests.log; grep "gram_impl=" gpurun_out/r02p_gpu_tests.log
timeout 300 python bench.py > gpurun_out/r02p_bench_n32768_1gpu.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02p_bench_n32768_1gpu.json') if l.startswith('{')][-1])
    print('1 GPU: %.4f evals/s e2e %.4f roofline %.3f parity %.1e potrf %.1f TF' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_rel_err'], d['potrf']['tflops']))
    for k, v in d['secondary'].items(): print(' ', k, round(v['value'], 2), v['metric'], 'frac', round(v['roofline']['frac'], 3))
except Exception as e: print('no line', e)
PY
timeout 100 ncu --set full --clock-control none --import-source on -k regex:gram_bwd_nkn -c 1 -f -o gpurun_out/r02_nkn_bwd_full python -m pytest tests/test_gpu_switches.py -m gpu -q -k speed_nkn -p no:cacheprovider > gpurun_out/r02p_ncu_nkn.log 2>&1; echo "ncu rc=$?"
