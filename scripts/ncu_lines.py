#!/usr/bin/env python
"""Per-source-line summary (samples, executed warp-instructions) of one kernel from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > f.csv`.
usage: python scripts/ncu_lines.py f.csv [top]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = []
    fpath = ""
    ix = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            ix = {h: i for i, h in enumerate(r)}
            S = r.index("# Samples")
            E = r.index("Instructions Executed")
            continue
        if ix is None or not r[0].isdigit():
            continue
        try:
            out.append((int(r[S]), int(r[E]), fpath, int(r[0]), r[1].strip()))
        except ValueError:
            pass
    agg = {}
    for o in out:                                   # several launches of the kernel: one entry per source line
        k = (o[2], o[3], o[4])
        a = agg.setdefault(k, [0, 0])
        a[0] += o[0]
        a[1] += o[1]
    out = [(v[0], v[1], k[0], k[1], k[2]) for (k, v) in agg.items()]
    ts = sum(o[0] for o in out)
    te = sum(o[1] for o in out)
    print("samples %d  executed warp-instructions %d" % (ts, te))
    for o in sorted(out, key=lambda o: -o[1])[:top]:
        print("%5.1f%% smp %5.1f%% exe  %s:%d  %s" % (100.0 * o[0] / max(ts, 1), 100.0 * o[1] / max(te, 1), o[2], o[3], o[4][:110]))


if __name__ == "__main__":
    main()
