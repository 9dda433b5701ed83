#!/usr/bin/env python
"""Per-stage DRAM traffic and fp64 work from ONE ncu pass over a k_split=1, n_split=1 C128 step (tools/profile_step.py):

    ncu --profile-from-start off --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/TAG_metrics_k1n1.csv \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum \
        python tools/profile_step.py
    python tools/stage_metrics.py gpurun_out/TAG_metrics_k1n1.csv profiles/traffic.json

The d_sw stage is the contiguous run of launches from the flux-prep kernel to the last kernel carrying fv3_d_sw in its
name (the stage is called once in that step); the other stages are recognised by their kernel names.  Writes
{"fv3_d_sw": DRAM bytes per call, ..., "_fp64_flops": {stage: DADD + DMUL + 2 * DFMA thread-instructions per call},
"_launches": ..., "_split": per-kernel record} — bench.py reads `roofline.traffic` and the fp64 roof from it.
"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "inst": 1}
BY_NAME = {"fv3_tracer_subcycle": "fv3_tracer_subcycle", "fv3_riem_solver3": "fv3_riem_solver3", "fv3_riem_solver_c": "fv3_riem_solver_c",
           "kmap": "fv3_map_multi", "fv3_c_sw": "fv3_c_sw", "fv3_nh_p_grad": "fv3_nh_p_grad"}


def main(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        d = per.setdefault(r[0], {"name": r[4], "m": {}})
        d["m"][r[-3]] = float(r[-1].replace(",", "")) * UNIT.get(r[-2], 1)
    launches = list(per.values())
    names = [l["name"] for l in launches]

    def bytes_of(l):
        return l["m"].get("dram__bytes_read.sum", 0) + l["m"].get("dram__bytes_write.sum", 0)

    def flops_of(l):
        m = l["m"]
        return (m.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0) + m.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0)
                + 2 * m.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0))

    first = next(i for i, n in enumerate(names) if "fv_prep" in n)
    last = max(i for i, n in enumerate(names) if "fv3_d_sw" in n)
    stages = {"fv3_d_sw": launches[first:last + 1]}
    for key, stage in BY_NAME.items():
        sel = [l for l in launches if key in l["name"] and l not in stages["fv3_d_sw"]]
        if sel:
            calls = 3 if stage == "fv3_tracer_subcycle" else (2 if stage == "fv3_map_multi" else 1)
            stages[stage] = sel
            stages[stage + "#calls"] = calls
    out = {"_source": src.split("/")[-1], "_fp64_flops": {}, "_launches": {}, "_time_ns": {}}
    for stage, sel in stages.items():
        if stage.endswith("#calls"):
            continue
        calls = stages.get(stage + "#calls", 1)
        out[stage] = sum(bytes_of(l) for l in sel) / calls
        out["_fp64_flops"][stage] = sum(flops_of(l) for l in sel) / calls
        out["_launches"][stage] = len(sel) / calls
        out["_time_ns"][stage] = sum(l["m"].get("gpu__time_duration.sum", 0) for l in sel) / calls
        print(f"{stage:22s} {len(sel) / calls:5.1f} launches/call  {out[stage] / 1e9:7.3f} GB DRAM  {out['_fp64_flops'][stage] / 1e9:8.2f} Gflop fp64  "
              f"{out['_time_ns'][stage] / 1e3:9.1f} us (cold, serialised)")
    split = collections.OrderedDict()
    for l in stages["fv3_d_sw"]:
        key = re.sub(r"\(.*", "", l["name"])
        key = re.sub(r"void fv3::|<unnamed>::|\[lambda", "", key)[:60]
        a = split.setdefault(key, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += bytes_of(l)
        a[2] += l["m"].get("gpu__time_duration.sum", 0)
        a[3] += flops_of(l)
    out["_split"] = {k: {"launches": v[0], "dram_bytes": v[1], "time_ns": v[2], "fp64_flops": v[3]} for k, v in split.items()}
    for k, v in split.items():
        print(f"   d_sw {v[0]:3d} x {k:60s} {v[1] / 1e6:9.1f} MB  {v[2] / 1e3:9.1f} us")
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
