#!/bin/bash
# Third GPU call of round 2 (1 GPU): everything new since the second call on the regular suite
# (big-leaf prefix solves, fast-path Gram input gradient, full-size parity), the new bench line,
# the per-rank share again, C4 launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02c_gpu_tests.log 2>&1
tail -4 gpurun_out/r02c_gpu_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench_1gpu.json 2> gpurun_out/r02c_bench_1gpu.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r02c_bench_1gpu.json; tail -3 gpurun_out/r02c_bench_1gpu.err
python tools/rank_share.py --size 32768 --world 8 --rank 0 --what inverse,bulk > gpurun_out/r02c_rank_share_w8.json 2> gpurun_out/r02c_rank_share_w8.err
tail -1 gpurun_out/r02c_rank_share_w8.json; tail -3 gpurun_out/r02c_rank_share_w8.err
python tools/rank_share.py --size 32768 --world 1 --rank 0 --what inverse > gpurun_out/r02c_rank_share_w1.json 2>&1
tail -1 gpurun_out/r02c_rank_share_w1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02c_launches_c4.csv \
    python tools/bench_secondary.py --what c4 --reps 2 > gpurun_out/r02c_c4_under_ncu.log 2>&1
python tools/summarise_launches.py gpurun_out/r02c_launches_c4.csv > gpurun_out/r02c_launches_c4_summary.txt 2>&1
head -16 gpurun_out/r02c_launches_c4_summary.txt
