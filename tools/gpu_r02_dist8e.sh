#!/bin/bash
# 8-GPU A/B #4 (short): v1 on one communicator (the round-1 arrangement) vs three.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601"
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
GPSLIM_NCCL_COMMS=1 timeout 120 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-secondary --schedule v1 > gpurun_out/r02l_v1_1comm.json 2> gpurun_out/r02l_v1_1comm.err
show gpurun_out/r02l_v1_1comm.json "v1 one communicator"
GPSLIM_NCCL_COMMS=1 CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 120 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-secondary --schedule v1 > gpurun_out/r02l_v1_1comm_mc8.json 2> gpurun_out/r02l_v1_1comm_mc8.err
show gpurun_out/r02l_v1_1comm_mc8.json "v1 one communicator maxconn8"
tail -2 gpurun_out/r02l_v1_1comm.err
