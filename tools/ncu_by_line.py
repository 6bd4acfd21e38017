"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import collections
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur, hdr = None, None
    samples = collections.Counter(); insts = collections.Counter(); text = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r; si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); continue
        if hdr and r[0] not in ("", "-") and len(r) > ii:      # a source line row (aggregated over its SASS)
            try:
                s, n = int(r[si]), int(r[ii])
            except ValueError:
                continue
            samples[(cur, int(r[0]))] += s; insts[(cur, int(r[0]))] += n; text[(cur, int(r[0]))] = r[1]
    tot = sum(samples.values()); ti = sum(insts.values())
    print(f"total samples {tot}, instructions {ti}")
    for k, s in samples.most_common(top):
        print(f"{100 * s / tot:5.1f}% smp  {100 * insts[k] / ti:5.1f}% inst  {k[0]}:{k[1]}: {text[k].strip()[:105]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
