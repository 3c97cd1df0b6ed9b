#!/bin/bash
# compute-sanitizer over the smallest shape of every sta_* kernel (run on a GPU box: gpurun -- bash tools/sanitize.sh).
#   memcheck   out-of-bounds / misaligned global, shared and TMEM-adjacent accesses, leaked allocations
#   racecheck  shared-memory hazards between the TMA / MMA / math warps (the generic-proxy side of the mbarrier protocols)
#   synccheck  illegal barrier use (named barriers, __syncwarp masks)
# Output: gpurun_out/sanitize_<tool>.log; exit code != 0 if any tool reports an error.
set -u
mkdir -p gpurun_out
rc=0
for tool in memcheck synccheck racecheck; do
  timeout ${SANITIZE_TIMEOUT:-600} compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
      python tools/sanitize_cases.py "$@" > gpurun_out/sanitize_$tool.log 2>&1
  code=$?
  echo "== $tool: exit $code; $(grep -c 'ok' gpurun_out/sanitize_$tool.log) case lines; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  [ $code -ne 0 ] && rc=1
done
exit $rc
