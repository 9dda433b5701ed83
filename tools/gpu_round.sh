#!/bin/bash
# usage (on the GPU box): tools/gpu_round.sh TAG ["kernel regex" ...]
# gpu tests, one bench line (+ per-stage table), the ncu launch list of one C128 k1n1 step, optional --set full captures
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --stage-table > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cat gpurun_out/${tag}_bench.json; tail -45 gpurun_out/${tag}_bench.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv \
    --log-file gpurun_out/${tag}_launches_k1n1.csv python tools/profile_step.py > gpurun_out/${tag}_launches.log 2>&1
tail -2 gpurun_out/${tag}_launches.log

# DRAM bytes per launch of the same step (profiles/traffic.json via tools/stage_traffic.py)
timeout 900 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    --kernel-name-base demangled --csv --log-file gpurun_out/${tag}_dram_k1n1.csv python tools/profile_step.py > gpurun_out/${tag}_dram.log 2>&1
tail -1 gpurun_out/${tag}_dram.log
if [ $# -gt 0 ]; then PROFILE_ARGS="" timeout 1500 tools/ncu_kernels.sh $tag "$@"; fi
