"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line (samples and instructions)."""
import collections
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main(path, top=24):
    rows = list(csv.reader(open(path)))
    kern = collections.OrderedDict(); fname = None; func = None; hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split('/')[-1]; continue
        if r[0] == "Function Name":
            func = r[1]; kern.setdefault(func, collections.defaultdict(lambda: [0, 0, ""])); continue
        if r[0] == "Line No":
            hdr = r; ixS = hdr.index("# Samples"); ixI = hdr.index("Instructions Executed"); continue
        if r[0] == "" or func is None:
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue
        e = kern[func][(fname, ln)]; e[0] += num(r[ixS]); e[1] += num(r[ixI]); e[2] = r[1]
    for name, lines in kern.items():
        tot_s = sum(v[0] for v in lines.values()); tot_i = sum(v[1] for v in lines.values())
        print("=====", name[:60], "samples", tot_s, "inst", tot_i)
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            print("  %5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100 * v[0] / max(1, tot_s), 100 * v[1] / max(1, tot_i), f[:16], ln, v[2].strip()[:84]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
