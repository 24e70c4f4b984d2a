#!/bin/bash
# First GPU call of round 2: validate what the last (GPU-less) session of round 1 left behind.
#   gpurun --timeout 900 -- bash tools/gpu_round2_first.sh
# 1. the whole GPU suite (incl. tests/test_gpu_zz_widened.py: mc_models / likelihoods_extra have
#    not run on a B200 yet);
# 2. the experimental shared-memory-accumulator Gram backward (gram_impl = 2) against the default
#    interpreter, on the kernel zoo and the NKN-GPR golden case, with a timing line;
# 3. C3 (NKN GPR N=16384) with gram_impl 0 and 2, to decide whether 2 becomes the default.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r02_gpu_tests.log 2>&1
tail -3 gpurun_out/r02_gpu_tests.log
python -m pytest tests/test_gpu_switches.py -q -s -p no:cacheprovider \
    > gpurun_out/r02_experimental.log 2>&1
tail -8 gpurun_out/r02_experimental.log
python tools/bench_secondary.py --what c3 > gpurun_out/r02_c3_gram_impl0.jsonl 2>&1
GPSLIM_GRAM_IMPL=2 python tools/bench_secondary.py --what c3 > gpurun_out/r02_c3_gram_impl2.jsonl 2>&1
tail -2 gpurun_out/r02_c3_gram_impl0.jsonl gpurun_out/r02_c3_gram_impl2.jsonl
# 4. where the 12 ms of a C4 (SVGP) step go: ncu launch list + per-kernel totals
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv \
    python tools/bench_secondary.py --what c4 --reps 2 > gpurun_out/r02_c4_under_ncu.log 2>&1
python tools/summarise_launches.py gpurun_out/r02_launches_c4.csv > gpurun_out/r02_launches_c4_summary.txt 2>&1
head -14 gpurun_out/r02_launches_c4_summary.txt
# 5. the 8-GPU headline after the host-overhead cuts needs its own call:
#    gpurun --gpus 8 --timeout 600 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3'

