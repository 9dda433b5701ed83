#!/usr/bin/env python
"""Where the executed instructions of a kernel go: per source file, per source line (top N) and per SASS opcode.
   python tools/ncu_instr.py FILE.ncu-rep [top]      (needs --import-source on, -lineinfo)"""
import collections
import csv
import io
import re
import subprocess
import sys


def rows_of(path, what):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", what], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(path, top=40):
    rows = rows_of(path, "cuda,sass")
    fpath, hdr, lines = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            lines.append((fpath, int(r[0]), r[1].strip(), dict(zip(hdr[4:], r[4:]))))
    tot = sum(float(d["Instructions Executed"] or 0) for *_, d in lines)
    byfile = collections.Counter()
    for f, n, src, d in lines:
        byfile[f] += float(d["Instructions Executed"] or 0)
    print(f"total warp instructions {tot:.3e}")
    print("per file:", {k: round(100 * v / tot, 1) for k, v in byfile.most_common()})
    for f, n, src, d in sorted(lines, key=lambda x: -float(x[3]["Instructions Executed"] or 0))[:top]:
        print(f"{100 * float(d['Instructions Executed'] or 0) / tot:5.1f}%i {f}:{n:<4d} {src[:120]}")
    rows = rows_of(path, "sass")
    hdr, ops, t2 = None, collections.Counter(), 0.0
    for r in rows:
        if hdr is None:
            if r and "Source" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            n = float(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", d.get("Source", ""))
        if m:
            ops[m.group(2)] += n
            t2 += n
    print("per opcode:", " ".join(f"{k} {100 * v / t2:.1f}%" for k, v in ops.most_common(24)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
