#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-secondary --no-cpu-baseline > gpurun_out/r02o_bench_1gpu.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02o_bench_1gpu.json') if l.startswith('{')][-1])
    print('1 GPU: %.4f evals/s e2e %.4f roofline %.3f achieved %.2f parity %.1e potrf %.2f TF launches %d' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['achieved'], d['parity_rel_err'], d['potrf']['tflops'], d['gpu_launches']))
except Exception as e: print('no line', e)
PY
