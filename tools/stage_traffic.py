#!/usr/bin/env python
"""DRAM traffic of one fv3_d_sw call from an ncu launch list of tools/profile_step.py (k_split=1, n_split=1):

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/TAG_dram_k1n1.csv python tools/profile_step.py
    python tools/stage_traffic.py gpurun_out/TAG_dram_k1n1.csv profiles/traffic.json

The launches of the d_sw stage are the contiguous run from the first fv3_fv_prep kernel to the last kernel whose name
carries fv3_d_sw (the stage is called once in that step).  Writes {"fv3_d_sw": bytes per call, ...} for bench.py's
`roofline.traffic`, plus the per-kernel split for the record.
"""
import collections
import csv
import json
import re
import sys


def main(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    per = collections.OrderedDict()   # launch id -> {name, metrics}
    for r in rows:
        d = per.setdefault(r[0], {"name": r[4], "m": {}})
        val = float(r[-1].replace(",", ""))
        unit = r[-2]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
        d["m"][r[-3]] = val * scale
    launches = list(per.values())
    names = [l["name"] for l in launches]
    first = next(i for i, n in enumerate(names) if "fv3_fv_prep" in n)
    last = max(i for i, n in enumerate(names) if "fv3_d_sw" in n)
    sel = launches[first:last + 1]
    tot = sum(l["m"].get("dram__bytes_read.sum", 0) + l["m"].get("dram__bytes_write.sum", 0) for l in sel)
    split = collections.OrderedDict()
    for l in sel:
        key = re.sub(r"\(.*", "", l["name"])
        key = re.sub(r"void fv3::|<unnamed>::|\[lambda", "", key)[:60]
        a = split.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += l["m"].get("dram__bytes_read.sum", 0) + l["m"].get("dram__bytes_write.sum", 0)
        a[2] += l["m"].get("gpu__time_duration.sum", 0)
    out = {"fv3_d_sw": tot, "_source": src.split("/")[-1], "_launches": len(sel),
           "_split": {k: {"launches": v[0], "dram_bytes": v[1], "time_ns": v[2]} for k, v in split.items()}}
    json.dump(out, open(dst, "w"), indent=1)
    print(f"fv3_d_sw: {len(sel)} launches, {tot / 1e9:.3f} GB DRAM traffic per call")
    for k, v in split.items():
        print(f"  {v[0]:3d} x {k:60s} {v[1] / 1e6:9.1f} MB  {v[2] / 1e3:9.1f} us")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
