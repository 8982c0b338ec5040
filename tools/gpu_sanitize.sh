#!/bin/bash
# compute-sanitizer over smoke() (memcheck, racecheck, synccheck); logs in gpurun_out/
set -u
O=gpurun_out; mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file $O/sanitize_$tool.log python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitize_$tool.out 2>&1
  echo "$tool rc=$?"; tail -2 $O/sanitize_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $O/sanitize_$tool.log | tail -3
done
