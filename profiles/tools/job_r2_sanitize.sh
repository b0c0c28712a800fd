#!/bin/bash
# compute-sanitizer over every kernel family at small sizes:  gpurun --timeout 1800 -- bash profiles/tools/job_r2_sanitize.sh
set -u
O=gpurun_out
mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool python profiles/tools/sanitize.py all > $O/r2f_sanitizer_$tool.log 2>&1
  tail -n 4 $O/r2f_sanitizer_$tool.log
done
