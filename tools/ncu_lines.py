#!/usr/bin/env python
"""Per-CUDA-source-line share of warp-stall samples and executed instructions from an ncu report
(ncu -i REP --page source --csv --print-source cuda,sass).  Usage: ncu_lines.py REP [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
kf = (["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else [])
txt = subprocess.run(["ncu", "-i", rep] + kf + ["--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur = r[1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No":
        hdr = r; n = len(hdr); iS = hdr.index("# Samples") - n; iI = hdr.index("Instructions Executed") - n; continue
    if hdr is None or not r or not r[0].isdigit(): continue
    try: smp = int(r[iS]); ins = int(r[iI])
    except Exception: continue
    out.append((smp, ins, cur.split("/")[-1], int(r[0]), ",".join(r[1:len(r) - n + 2])[:100]))
tot = sum(o[0] for o in out) or 1; toti = sum(o[1] for o in out) or 1
print("total samples", tot, "warp instructions", toti)
pf = collections.Counter(); pi = collections.Counter()
for o in out: pf[o[2]] += o[0]; pi[o[2]] += o[1]
for k, v in pf.most_common(): print("  %-28s samples %5.1f%%  instructions %5.1f%%" % (k, 100 * v / tot, 100 * pi[k] / toti))
for o in sorted(out, reverse=True)[:top]: print("%5.2f%% %5.2f%%i %s:%d %s" % (100 * o[0] / tot, 100 * o[1] / toti, o[2], o[3], o[4]))
