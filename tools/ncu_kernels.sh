#!/bin/bash
# usage: tools/ncu_kernels.sh TAG "regex1" "regex2" ...   (on the GPU box) — one --set full capture (first launch) per regex
tag=$1; shift
mkdir -p gpurun_out
for re in "$@"; do
  name=$(echo "$re" | tr -c 'A-Za-z0-9_' '_' | cut -c1-40)
  ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k regex:"$re" -c 1 -f -o gpurun_out/${tag}_${name} python tools/profile_step.py ${PROFILE_ARGS} > gpurun_out/${tag}_${name}.log 2>&1
  tail -1 gpurun_out/${tag}_${name}.log
done
ls -la gpurun_out
