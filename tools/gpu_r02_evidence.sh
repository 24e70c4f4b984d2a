#!/bin/bash
# Round-2 evidence call (1 GPU): suite, smoke, ncu --set full of the dominant kernel, launch list
# of one bench step, the bench line itself (with secondary configs + CPU baseline), reference arm.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02m_smoke.log 2>&1; tail -1 gpurun_out/r02m_smoke.log
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02m_gpu_tests.log 2>&1; tail -2 gpurun_out/r02m_gpu_tests.log
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_nt_tma -c 1 -f -o gpurun_out/r02_gemm8192_tma_full python tools/prof_gpr.py --what gemm --n 8192 > gpurun_out/r02m_prof_gemm.log 2>&1
echo "ncu gemm rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_n32768.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02m_bench_under_ncu.log 2>&1
python tools/summarise_launches.py gpurun_out/r02_launches_bench_n32768.csv > gpurun_out/r02_launches_bench_n32768_summary.txt 2>&1; head -8 gpurun_out/r02_launches_bench_n32768_summary.txt
timeout 400 python bench.py > gpurun_out/r02_bench_n32768_1gpu.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02_bench_n32768_1gpu.json') if l.startswith('{')][-1])
    print('1 GPU: %.4f evals/s e2e %.4f roofline %.3f parity %.1e potrf %.1f TF' % (d['value'], d['e2e']['value'], d['roofline']['frac'], d['parity_rel_err'], d['potrf']['tflops']))
    for k, v in d['secondary'].items(): print(' ', k, round(v['value'], 2), v['metric'], 'frac', round(v['roofline']['frac'], 3))
    print(' cpu', d['cpu_baseline']['value'], d['cpu_baseline']['sample'][:120])
except Exception as e: print('no line', e)
PY
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-250 gpurun_out/r02_bench_reference_arm.json
timeout 200 python tools/rank_share.py --size 32768 --world 8 --rank 0 --what bulk > gpurun_out/r02m_rank_share_bulk.json 2>&1; tail -1 gpurun_out/r02m_rank_share_bulk.json
