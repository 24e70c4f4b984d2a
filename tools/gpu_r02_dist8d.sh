#!/bin/bash
# 8-GPU A/B #3: NCCL CTA budgets per communicator (a spinning collective pins its CTAs).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591"
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
run() { # name, env..., extra args
  name=$1; shift
  env "$@" timeout 200 $TR bench.py --gpus 8 --steps 3 --warmup 2 --no-secondary $EXTRA > gpurun_out/r02k_$name.json 2> gpurun_out/r02k_$name.err
  show gpurun_out/r02k_$name.json "$name"
}
EXTRA="" run v2_ctas_2_2_8 GPSLIM_NCCL_CTAS=2,2,8
EXTRA="--schedule v1" run v1_ctas_2_2_8 GPSLIM_NCCL_CTAS=2,2,8
EXTRA="" run v2_ctas_1_1_4 GPSLIM_NCCL_CTAS=1,1,4
EXTRA="--schedule v1" run v1_ctas_4_4_16 GPSLIM_NCCL_CTAS=4,4,16
tail -2 gpurun_out/r02k_v2_ctas_2_2_8.err
