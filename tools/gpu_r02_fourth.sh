#!/bin/bash
# 1-GPU check of the generalised split-K (triangular operands, small slices) and the level-batched
# triangular inverse before the 8-GPU run.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02e_gpu_tests.log 2>&1
tail -3 gpurun_out/r02e_gpu_tests.log
python tools/leaf_probe.py > gpurun_out/r02e_leaf_probe.txt 2>&1; cat gpurun_out/r02e_leaf_probe.txt
python tools/bench_secondary.py --what c4 > gpurun_out/r02e_c4.jsonl 2>&1; grep '^{' gpurun_out/r02e_c4.jsonl | cut -c1-400
python tools/bench_secondary.py --what c2 > gpurun_out/r02e_c2.jsonl 2>&1; grep '^{' gpurun_out/r02e_c2.jsonl | cut -c1-300
python tools/bench_secondary.py --what potrf > gpurun_out/r02e_potrf.jsonl 2>&1; grep '^{' gpurun_out/r02e_potrf.jsonl | cut -c1-200
python tools/rank_share.py --size 32768 --world 8 --rank 0 --what chain,inverse > gpurun_out/r02e_rank_share_w8.json 2> gpurun_out/r02e_rank_share_w8.err
tail -1 gpurun_out/r02e_rank_share_w8.json; tail -2 gpurun_out/r02e_rank_share_w8.err
