#!/bin/bash
# 8-GPU A/B of the factorisation schedules + per-panel timeline.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561"
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    ph = d.get('phases_ms') or {}
    print('%-28s %.3f evals/s  %.1f ms  factor %.1f  U %.1f  Kinv %.1f  parity %.1e' % (
        sys.argv[2], d['value'], d['ms_per_step'], ph.get('factor(lookahead)', 0), ph.get('rows_of_U', 0),
        ph.get('rows_of_Kinv', 0), d.get('parity_rel_err') or 0))
except Exception as e:
    print(sys.argv[2], 'no line', e)
PY
}
timeout 300 $TR tools/dist_trace.py --size 32768 > gpurun_out/r02g_trace_v2.txt 2>&1
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02g_trace_v2.txt | head -75
timeout 300 $TR bench.py --gpus 8 --steps 4 --warmup 3 --no-secondary > gpurun_out/r02g_bench_v2_5s.json 2> gpurun_out/r02g_bench_v2_5s.err
show gpurun_out/r02g_bench_v2_5s.json "v2 narrow-stream"
GPSLIM_DIST_NARROW_STREAM=0 timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-secondary > gpurun_out/r02g_bench_v2_4s.json 2> gpurun_out/r02g_bench_v2_4s.err
show gpurun_out/r02g_bench_v2_4s.json "v2 four streams"
timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-secondary --schedule v1 > gpurun_out/r02g_bench_v1.json 2> gpurun_out/r02g_bench_v1.err
show gpurun_out/r02g_bench_v1.json "v1 (round 1 schedule)"
tail -2 gpurun_out/r02g_bench_v1.err
