#!/bin/bash
# Wall time of `pRIblast_b200 db` on cfg2 in full (100,000 transcripts): the default (GPUs recruited on demand) and
# fixed worker counts (PRIB_NUM_GPUS = 1, 2, 4, 8: all workers started at once).  One worker process per GPU.
# usage: bash profiles/db_scaling.sh [n_transcripts]   (run under gpurun --gpus 8)
cd "$(dirname "$0")/.."
N=${1:-100000}
NG=$(nvidia-smi -L | wc -l)
echo "=== default: demand-driven recruitment ($NG GPUs visible)"
python profiles/db_e2e.py "$N" 2>&1 | grep -E "recruit|used|chunks|context|seq/.ind|wall"
for g in 1 2 4 8; do
  if [ "$g" -le "$NG" ]; then
    echo "=== PRIB_NUM_GPUS=$g (all started at once)"
    PRIB_NUM_GPUS=$g python profiles/db_e2e.py "$N" 2>&1 | grep -E "chunks|context|seq/.ind|wall"
  fi
done
