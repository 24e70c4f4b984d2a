#!/bin/bash
# Lean multi-GPU call: parity check, phase timing for the block sizes in $BLOCKS, bench.
NG=${NG:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 tools/dist_check.py --size 3000 > gpurun_out/dist_check_$NG.log 2>&1
echo "dist_check exit $?"; grep -E "DIST_CHECK_OK|Error|error" gpurun_out/dist_check_$NG.log | tail -3
for B in ${BLOCKS:-512}; do
timeout 200 $TR --master-port 29543 tools/dist_time.py --size 32768 --fused 0 --reps 2 --block $B > gpurun_out/dist_time_${NG}_32k_b$B.log 2>&1
echo "dist_time 32k block $B exit $?"; grep -E "^distributed|^phases" gpurun_out/dist_time_${NG}_32k_b$B.log
done
timeout 300 $TR --master-port 29544 bench.py --gpus $NG --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$NG.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench_$NG.log
