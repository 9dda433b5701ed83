#!/bin/bash
# usage (on the GPU box): tools/gpu_quick.sh TAG [pytest args...]  — gpu tests + one bench line with the per-stage table
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt
if [ "$1" != "nobench" ]; then
timeout 900 python bench.py --stage-table --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json; tail -40 gpurun_out/${tag}_bench.err
else shift; fi
timeout 1500 python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
