#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel: python tools/agg_launches.py FILE"""
import collections
import csv
import re
import sys


def main(path, top=60):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name, t = r[4], float(r[-1])
        key = re.sub(r"\(.*", "", name)
        key = re.sub(r"void fv3::|<unnamed>::|\[lambda", "", key)[:110]
        grid, blk = r[8], r[7]
        a = agg.setdefault(key, [0, 0.0, grid, blk])
        a[0] += 1
        a[1] += t
    tot = sum(a[1] for a in agg.values())
    print(f"{len(rows)} launches, {tot / 1e6:.3f} ms total (serialised, cold cache)")
    for k, (n, t, grid, blk) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t / 1e3:10.1f} us {100 * t / tot:5.1f}% {n:5d} x {t / n / 1e3:8.1f} us  {k}  grid {grid} block {blk}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
