#!/bin/bash
# Wall time of `pRIblast_b200 db` on cfg2 in full (100,000 transcripts) with 1, 2, 4 and all GPUs of the box
# (one worker process per GPU).  usage: bash profiles/db_scaling.sh [n_transcripts]   (run under gpurun --gpus 8)
cd "$(dirname "$0")/.."
N=${1:-100000}
NG=$(nvidia-smi -L | wc -l)
for g in 1 2 4 8; do
  if [ "$g" -le "$NG" ]; then
    echo "=== PRIB_NUM_GPUS=$g"
    PRIB_NUM_GPUS=$g python profiles/db_e2e.py "$N" 2>&1 | grep -E "using|worker: prib_acc_run|worker: context|seq/.ind|wall" | sort | uniq -c | sort -k2 | head -12
  fi
done
