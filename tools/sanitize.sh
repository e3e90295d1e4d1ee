#!/bin/bash
# compute-sanitizer pass over the hot kernels on small configurations:
#   memcheck + racecheck + synccheck over em_team_kernel (HOT: copy warp + mbarrier ring, and generic), em_kernel (cp.async
#   and TMA-bulk variants), em_group_kernel, the parallel-in-time scan kernels, quad / ckf / rollout kernels.
# Usage: tools/sanitize.sh <out_dir>      (logs: <out_dir>/sanitize_<tool>.log; exit code != 0 if a tool reports errors)
out=${1:-gpurun_out}
mkdir -p "$out"
rc=0
for tool in memcheck racecheck synccheck; do
  log="$out/sanitize_$tool.log"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_driver.py > "$log" 2>&1
  r=$?
  echo "$tool: exit $r; $(grep -c 'ERROR SUMMARY' "$log") summary line(s): $(grep 'ERROR SUMMARY' "$log" | tail -1)"
  [ $r -ne 0 ] && rc=1
done
exit $rc
