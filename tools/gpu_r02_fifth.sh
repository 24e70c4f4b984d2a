#!/bin/bash
# 1-GPU check: persistent TMA GEMM, register-resident 8x8 diagonal factorisation of the leaf.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02h_gpu_tests.log 2>&1
tail -3 gpurun_out/r02h_gpu_tests.log
python tools/leaf_probe.py > gpurun_out/r02h_leaf_probe.txt 2>&1; cat gpurun_out/r02h_leaf_probe.txt
python tools/bench_secondary.py --what c2 > gpurun_out/r02h_c2.jsonl 2>&1; grep '^{' gpurun_out/r02h_c2.jsonl | cut -c1-300
python tools/bench_secondary.py --what potrf > gpurun_out/r02h_potrf.jsonl 2>&1; grep '^{' gpurun_out/r02h_potrf.jsonl | cut -c1-200
python tools/bench_secondary.py --what c4 > gpurun_out/r02h_c4.jsonl 2>&1; grep '^{' gpurun_out/r02h_c4.jsonl | cut -c1-330
python tools/rank_share.py --size 32768 --world 8 --rank 0 --what chain > gpurun_out/r02h_rank_share_chain.json 2>&1; tail -1 gpurun_out/r02h_rank_share_chain.json
