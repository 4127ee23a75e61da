#!/bin/bash
# One GPU call: the -m gpu suite, the default bench line, then (each only after its command exited 0
# without a profiler) the ncu launch list of one eager step with per-launch DRAM bytes.
# usage (from the repo root, on the GPU box): bash tools/final_profile.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_tests_$tag.log
cat gpurun_out/r2_tests_$tag.log
timeout 400 python bench.py > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err || exit 1
python -c "import json;d=json.load(open('gpurun_out/r2_bench_$tag.json'));print(d['value'],d['ms_per_step'],d['e2e']['ms_per_step']);print({k:(round(v['avg_ms'],4),round(v['frac'],3)) for k,v in d['kernels'].items()});print(d['config'].get('producers'))"
timeout 300 python tools/profile_step.py bf16 1 > gpurun_out/r2_profile_step_$tag.log 2>&1 || exit 1
tail -1 gpurun_out/r2_profile_step_$tag.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file gpurun_out/r2_launches_$tag.csv python tools/profile_step.py bf16 1 \
  > gpurun_out/r2_ncu_step_$tag.log 2>&1
wc -l gpurun_out/r2_launches_$tag.csv
