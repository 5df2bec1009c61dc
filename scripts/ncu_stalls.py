#!/usr/bin/env python
"""Per-opcode / per-instruction stall summary of one kernel from `ncu --page source --csv`.
usage: ncu -i X.ncu-rep --page source --csv --kernel-name regex:K > src.csv; python scripts/ncu_stalls.py src.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    # the file may hold several launches of the kernel: use the first block
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[start + 1:]:
        if not r or r[0] in ("Address", "Kernel Name"):
            break
        if len(r) >= len(hdr):
            data.append(r)
    S, E = ix["# Samples"], ix["Instructions Executed"]
    tot = sum(int(r[S]) for r in data)
    te = sum(int(r[E]) for r in data)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    print("total samples", tot, "static instructions", len(data), "executed warp-instructions", te)
    stalls = collections.Counter()
    byop, execs = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ix["Source"]].strip().split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        byop[op] += int(r[S])
        execs[op] += int(r[E])
        for h in stall_cols:
            stalls[h] += int(r[ix[h]])
    print("stall totals:", [(h, stalls[h]) for h in sorted(stall_cols, key=lambda h: -stalls[h])[:8]])
    for op, c in byop.most_common(16):
        print("%-8s samples %6d (%4.1f%%)  executed %10d (%4.1f%%)" % (op, c, 100.0 * c / tot, execs[op], 100.0 * execs[op] / te))
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    print("--- top instructions by samples")
    for r in sorted(data, key=lambda r: -int(r[S]))[:n]:
        st = {h: int(r[ix[h]]) for h in stall_cols if int(r[ix[h]]) > 0}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(r[S].rjust(6), r[ix["Source"]].strip()[:64].ljust(64), top)


if __name__ == "__main__":
    main()
