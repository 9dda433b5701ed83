#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of an .ncu-rep (needs --import-source on, -lineinfo):
   python tools/ncu_lines.py FILE.ncu-rep [top]"""
import csv
import io
import subprocess
import sys


def main(path, top=45):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fpath, hdr, lines = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            d = dict(zip(hdr[4:], r[4:]))
            lines.append((fpath, int(r[0]), r[1].strip(), d))
    tot_i = sum(float(d["Instructions Executed"] or 0) for _, _, _, d in lines)
    tot_s = sum(float(d["# Samples"] or 0) for _, _, _, d in lines)
    print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
    for f, n, src, d in sorted(lines, key=lambda x: -float(x[3]["# Samples"] or 0))[:top]:
        st = sorted(((float(d[k] or 0), k[6:]) for k in d if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:3]
        sts = " ".join(f"{k}:{100 * v / max(1.0, float(d['# Samples'] or 1)):.0f}%" for v, k in st)
        print(f"{100 * float(d['# Samples'] or 0) / tot_s:5.1f}%s {100 * float(d['Instructions Executed'] or 0) / tot_i:5.1f}%i {f}:{n:<4d} {src[:90]:90s} | {sts}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
