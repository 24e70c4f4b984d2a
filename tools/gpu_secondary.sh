#!/bin/bash
# Secondary workloads: C2, C3, C4, POTRF sweep -> gpurun_out/secondary.jsonl
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
: > gpurun_out/secondary.jsonl
for w in potrf c2 c3 c4; do
  timeout 300 python tools/bench_secondary.py --what $w > gpurun_out/sec_$w.log 2>&1
  echo "$w exit $?"; grep '^{' gpurun_out/sec_$w.log | tee -a gpurun_out/secondary.jsonl | cut -c1-400
  grep -E "Error|error|Traceback" gpurun_out/sec_$w.log | head -5
done
