#!/bin/bash
# compute-sanitizer passes over the smoke invocation of the hot path: memcheck, racecheck, initcheck,
# synccheck.  usage: bash profiles/sanitize.sh > profiles/r2/sanitizer.txt   (under gpurun, one GPU)
cd "$(dirname "$0")/.."
for tool in memcheck racecheck initcheck synccheck; do
  echo "=== compute-sanitizer --tool $tool python -c '__graft_entry__.smoke()'"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "ERROR SUMMARY|smoke ok|Error|error|RACECHECK SUMMARY|hazard" | head -12
done
