#!/bin/bash
# compute-sanitizer passes over the smoke invocation of the hot path: memcheck, racecheck, initcheck,
# synccheck.  usage: bash profiles/sanitize.sh > profiles/r2/sanitizer.txt   (under gpurun, one GPU)
cd "$(dirname "$0")/.."
for tool in memcheck racecheck initcheck synccheck; do
  echo "=== compute-sanitizer --tool $tool python -c '__graft_entry__.smoke()'"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" > /tmp/san_$tool.log 2>&1
  echo "exit code $?"
  grep -E "ERROR SUMMARY|smoke ok|RACECHECK SUMMARY|hazard|Invalid|Uninitialized|Barrier error" /tmp/san_$tool.log | head -12
  tail -3 /tmp/san_$tool.log
done
