#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581"
timeout 300 $TR tools/dist_hostprof.py --size 32768 > gpurun_out/r02j_hostprof_world2.txt 2>&1
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02j_hostprof_world2.txt | head -45
timeout 300 $TR tools/dist_trace.py --size 32768 --out gpurun_out/r02j_trace > gpurun_out/r02j_trace_world2.txt 2>&1
grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/r02j_trace_world2.txt | awk 'NR<=14' 
