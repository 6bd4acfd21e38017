"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by function of solver_core.cuh."""
import collections
import csv
import re
import sys


def main(path, src_path="cave_b200/csrc/solver_core.cuh"):
    rows = list(csv.reader(open(path)))
    cur = hdr = None
    S = collections.Counter(); I = collections.Counter()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
        if hdr and r[0] not in ("", "-") and len(r) > ii:
            try:
                s, n = int(r[si]), int(r[ii])
            except ValueError:
                continue
            S[(cur, int(r[0]))] += s; I[(cur, int(r[0]))] += n
    tot = sum(S.values()); ti = sum(I.values())
    bounds = []
    for i, l in enumerate(open(src_path).read().split("\n"), 1):
        m = re.match(r"^CAVE_DEV\s+[\w<>:\*& ]+?\s+(\w+)\(", l)
        if m:
            bounds.append((i, m.group(1)))

    def fn(line):
        name = "?"
        for b, n in bounds:
            if b <= line:
                name = n
            else:
                break
        return name
    agg = collections.Counter(); aggi = collections.Counter()
    for (f, l), s in S.items():
        key = fn(l) if f == src_path.split("/")[-1] else f
        agg[key] += s; aggi[key] += I[(f, l)]
    print(f"total samples {tot}, instructions {ti}")
    for k, s in agg.most_common(20):
        print(f"{100 * s / tot:5.1f}% smp {100 * aggi[k] / ti:5.1f}% inst  {k}")


if __name__ == "__main__":
    main(*sys.argv[1:])
