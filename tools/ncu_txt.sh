#!/bin/bash
# usage (on the GPU box): tools/ncu_txt.sh TAG "regex1" ...  — one --set full capture per regex, reduced on the box to the
# text summary + per-line stall table (tools/ncu_summary.py, tools/ncu_lines.py); the .ncu-rep files are not kept
tag=$1; shift
mkdir -p gpurun_out
for re in "$@"; do
  name=$(echo "$re" | tr -c 'A-Za-z0-9_' '_' | cut -c1-40)
  ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:"$re" -c 1 -f -o /tmp/${tag}_${name} python tools/profile_step.py ${PROFILE_ARGS} > gpurun_out/${tag}_${name}.log 2>&1
  { python tools/ncu_summary.py /tmp/${tag}_${name}.ncu-rep; python tools/ncu_lines.py /tmp/${tag}_${name}.ncu-rep | head -60; } > gpurun_out/${tag}_ncu_${name}.txt 2>&1
  head -14 gpurun_out/${tag}_ncu_${name}.txt | tail -12
done
