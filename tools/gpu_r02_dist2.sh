#!/bin/bash
# 2-GPU call: the whole GPU suite (incl. the 2-rank NCCL check of the new four-stream schedule,
# distributed predict / SGPR / NKN), phases at N=32768 on 2 ranks, bench.py --gpus 2.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02d_gpu_tests.log 2>&1
tail -4 gpurun_out/r02d_gpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR tools/dist_check.py --size 4000 --block 512 --repeat 4 > gpurun_out/r02d_dist_check_world2.txt 2>&1
grep -h "rank 0\|DIST_CHECK\|Error\|assert" gpurun_out/r02d_dist_check_world2.txt | tail -8
timeout 600 $TR tools/dist_time.py --size 32768 --fused 0 > gpurun_out/r02d_dist_phases_world2.txt 2>&1
grep -h "distributed\|phases" gpurun_out/r02d_dist_phases_world2.txt
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r02d_bench_2gpu.json 2> gpurun_out/r02d_bench_2gpu.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02d_bench_2gpu.json; tail -3 gpurun_out/r02d_bench_2gpu.err
