#!/bin/bash
# Staged GPU check: each pytest file in its own process (a device trap in one stage must not poison the next),
# every stage under its own timeout, logs merged back through gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
rc=0
for t in "$@"; do
  name=$(echo "$t" | tr '/:[]' '____')
  timeout 900 python -m pytest "$t" -m gpu -q --no-header -rA -p no:cacheprovider > "gpurun_out/stage_${name}.log" 2>&1
  r=$?
  echo "stage $t rc=$r"
  grep -E "PARITY|passed|failed|Error" "gpurun_out/stage_${name}.log" | tail -n 40
  [ $r -ne 0 ] && rc=$r
done
exit $rc
