#!/bin/bash
# usage (on the GPU box): tools/variants.sh TAG NAME...   — bench line + stage table of each variants/NAME.so
tag=$1; shift
mkdir -p gpurun_out
cp pace_b200/libfv3b200.so /tmp/base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so pace_b200/libfv3b200.so; else cp variants/$v.so pace_b200/libfv3b200.so; fi
  timeout 600 python bench.py --stage-table --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/${tag}_${v}.json 2> gpurun_out/${tag}_${v}.err
  echo "== $v: $(python -c "import json,sys; d=json.load(open('gpurun_out/${tag}_${v}.json')); print(d['ms_per_step'], d['parity']['status'], d['state_digest'])")"
  grep calls gpurun_out/${tag}_${v}.err | head -${VARIANT_LINES:-9}
done
cp /tmp/base.so pace_b200/libfv3b200.so
