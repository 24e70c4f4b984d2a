#!/bin/bash
# Second GPU call of round 2 (1 GPU): split-K in the TMA kernel + new adjoint defaults, C4 again,
# the per-rank share of an 8-rank evaluation on one GPU, leaf-kernel ncu capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -s -p no:cacheprovider > gpurun_out/r02b_gpu_tests.log 2>&1
tail -5 gpurun_out/r02b_gpu_tests.log
grep -h "split-K\|graphed\|gram_impl" gpurun_out/r02b_gpu_tests.log
python tools/leaf_probe.py > gpurun_out/r02b_leaf_probe.txt 2>&1; cat gpurun_out/r02b_leaf_probe.txt
python tools/bench_secondary.py --what c4 > gpurun_out/r02b_c4.jsonl 2>&1; tail -1 gpurun_out/r02b_c4.jsonl
python tools/bench_secondary.py --what c2 > gpurun_out/r02b_c2.jsonl 2>&1; tail -2 gpurun_out/r02b_c2.jsonl
python tools/rank_share.py --size 32768 --world 8 --rank 0 > gpurun_out/r02b_rank_share_w8.json 2> gpurun_out/r02b_rank_share_w8.err
tail -1 gpurun_out/r02b_rank_share_w8.json; tail -3 gpurun_out/r02b_rank_share_w8.err
timeout 300 ncu --set full --import-source on --clock-control none -k regex:potrf_leaf_kernel -c 2 -o gpurun_out/r02b_leaf_full \
    python tools/leaf_probe.py 128 > gpurun_out/r02b_leaf_ncu.log 2>&1
tail -2 gpurun_out/r02b_leaf_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches_rank_share_inverse.csv \
    python tools/rank_share.py --size 32768 --world 8 --rank 0 --what inverse > gpurun_out/r02b_rank_share_under_ncu.log 2>&1
python tools/summarise_launches.py gpurun_out/r02b_launches_rank_share_inverse.csv | head -12
