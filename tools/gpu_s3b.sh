#!/bin/bash
# Multi-GPU call (NG ranks): NCCL parity check, phase timing (with / without look-ahead), bench.
NG=${NG:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/nvsmi_b.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
if [ "$QUICK" != "1" ]; then
timeout 300 $TR --master-port 29541 tools/dist_check.py --size 3000 > gpurun_out/dist_check_$NG.log 2>&1
echo "dist_check exit $?"; grep -E "rel err|DIST_CHECK_OK|Error|error" gpurun_out/dist_check_$NG.log | tail -4
timeout 300 $TR --master-port 29545 tools/dist_time.py --size 32768 --fused 0 --lookahead 0 --block ${BLOCK:-512} > gpurun_out/dist_time_${NG}_32k_nola.log 2>&1
echo "dist_time 32k no-lookahead exit $?"; grep -E "^distributed|^phases" gpurun_out/dist_time_${NG}_32k_nola.log
fi
for B in ${BLOCKS:-512}; do
timeout 300 $TR --master-port 29543 tools/dist_time.py --size 32768 --fused 0 --block $B > gpurun_out/dist_time_${NG}_32k_b$B.log 2>&1
echo "dist_time 32k block $B exit $?"; grep -E "^distributed|^phases" gpurun_out/dist_time_${NG}_32k_b$B.log
done
timeout 400 $TR --master-port 29544 bench.py --gpus $NG --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$NG.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench_$NG.log
