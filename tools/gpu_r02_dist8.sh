#!/bin/bash
# 8-GPU call: correctness at world 8, the headline at block 512 and 1024, N = 65536.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
timeout 400 $TR tools/dist_check.py --size 6000 --block 256 --repeat 3 > gpurun_out/r02f_dist_check_world8.txt 2>&1
grep -h "rank 0\|DIST_CHECK\|Error\|assert" gpurun_out/r02f_dist_check_world8.txt | tail -8
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02f_bench_8gpu.json 2> gpurun_out/r02f_bench_8gpu.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02f_bench_8gpu.json') if l.startswith('{')][-1])
    print('8 GPUs block 512:', d['value'], 'evals/s', d['ms_per_step'], 'ms', 'parity', d.get('parity_rel_err'))
    print(d.get('phases_ms')); print(d.get('potrf')); print(d.get('secondary'))
except Exception as e: print('no line', e)
PY
tail -3 gpurun_out/r02f_bench_8gpu.err
timeout 400 $TR bench.py --gpus 8 --steps 4 --warmup 2 --block 1024 --no-secondary > gpurun_out/r02f_bench_8gpu_block1024.json 2> gpurun_out/r02f_bench_8gpu_block1024.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02f_bench_8gpu_block1024.json') if l.startswith('{')][-1])
    print('8 GPUs block 1024:', d['value'], 'evals/s', d['ms_per_step'], 'ms', 'parity', d.get('parity_rel_err'))
    print(d.get('phases_ms'))
except Exception as e: print('no line', e)
PY
timeout 600 $TR bench.py --gpus 8 --size 65536 --steps 2 --warmup 1 --no-secondary > gpurun_out/r02f_bench_8gpu_n65536.json 2> gpurun_out/r02f_bench_8gpu_n65536.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02f_bench_8gpu_n65536.json') if l.startswith('{')][-1])
    print('8 GPUs N=65536:', d['value'], 'evals/s', d['ms_per_step'], 'ms')
    print(d.get('phases_ms')); print(d.get('potrf'))
except Exception as e: print('no line', e)
PY
tail -3 gpurun_out/r02f_bench_8gpu_n65536.err
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
