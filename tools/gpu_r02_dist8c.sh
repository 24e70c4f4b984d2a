#!/bin/bash
# 8-GPU A/B #2: one hardware queue per stream (CUDA_DEVICE_MAX_CONNECTIONS=32, now the default),
# persistent bulk GEMM with reserved SMs, v1 vs v2.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571"
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    ph = d.get('phases_ms') or {}
    print('%-34s %.3f evals/s  %.1f ms  factor %.1f  U %.1f  Kinv %.1f  parity %.1e' % (
        sys.argv[2], d['value'], d['ms_per_step'], ph.get('factor(lookahead)', 0), ph.get('rows_of_U', 0),
        ph.get('rows_of_Kinv', 0), d.get('parity_rel_err') or 0))
except Exception as e:
    print(sys.argv[2], 'no line', e)
PY
}
timeout 300 $TR tools/dist_trace.py --size 32768 --out gpurun_out/r02i_trace > gpurun_out/r02i_trace_v2_maxconn32.txt 2>&1
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r02i_trace_v2_maxconn32.txt | awk 'NR<=3 || NR%3==0' | head -30
timeout 300 $TR bench.py --gpus 8 --steps 4 --warmup 3 --no-secondary > gpurun_out/r02i_bench_v2.json 2> gpurun_out/r02i_bench_v2.err
show gpurun_out/r02i_bench_v2.json "v2 maxconn32"
GPSLIM_GEMM_RESERVE_SMS=12 timeout 300 $TR bench.py --gpus 8 --steps 4 --warmup 3 --no-secondary > gpurun_out/r02i_bench_v2_res12.json 2> gpurun_out/r02i_bench_v2_res12.err
show gpurun_out/r02i_bench_v2_res12.json "v2 maxconn32 reserve 12 SMs"
GPSLIM_GEMM_RESERVE_SMS=24 timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-secondary > gpurun_out/r02i_bench_v2_res24.json 2> gpurun_out/r02i_bench_v2_res24.err
show gpurun_out/r02i_bench_v2_res24.json "v2 maxconn32 reserve 24 SMs"
timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-secondary --schedule v1 > gpurun_out/r02i_bench_v1.json 2> gpurun_out/r02i_bench_v1.err
show gpurun_out/r02i_bench_v1.json "v1 maxconn32"
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 300 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-secondary --schedule v1 > gpurun_out/r02i_bench_v1_mc8.json 2> gpurun_out/r02i_bench_v1_mc8.err
show gpurun_out/r02i_bench_v1_mc8.json "v1 maxconn8"
tail -2 gpurun_out/r02i_bench_v1.err
