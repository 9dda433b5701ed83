#!/bin/bash
# usage (on the GPU box): tools/gpu_round.sh TAG ["kernel regex" ...]
# gpu tests, one full bench line (+ per-stage table), the ncu launch list of one C128 k1n1 step with DRAM bytes and fp64
# operation counts per launch (tools/stage_metrics.py -> profiles/traffic.json), optional --set full captures
tag=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --stage-table > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
cut -c1-600 gpurun_out/${tag}_bench.json; grep calls gpurun_out/${tag}_bench.err > gpurun_out/${tag}_stage_table.txt; head -12 gpurun_out/${tag}_stage_table.txt
timeout 900 ncu --profile-from-start off --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/${tag}_metrics_k1n1.csv \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum \
    python tools/profile_step.py > gpurun_out/${tag}_metrics.log 2>&1
tail -1 gpurun_out/${tag}_metrics.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --csv \
    --log-file gpurun_out/${tag}_launches_k1n1.csv python tools/profile_step.py > gpurun_out/${tag}_launches.log 2>&1
tail -1 gpurun_out/${tag}_launches.log
if [ $# -gt 0 ]; then PROFILE_ARGS="" timeout 1500 tools/ncu_kernels.sh $tag "$@"; fi
