#!/usr/bin/env python
"""Executed-instruction and stall-sample shares per SASS opcode of one kernel of an ncu report.
Usage: ncu_opcodes.py REP KERNEL_REGEX [top]"""
import csv, subprocess, collections, sys
rep, kre = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 22
txt = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kre, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "Address": hdr = r; start = i + 1; break
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); byop = collections.defaultdict(collections.Counter); cnt = collections.Counter()
for r in rows[start:]:
    if len(r) < len(hdr): continue
    w = r[ix["Source"]].split()
    if not w: continue
    op = (w[1] if w[0].startswith("@") and len(w) > 1 else w[0]).split(".")[0]
    try: n = int(r[ix["Instructions Executed"]])
    except Exception: continue
    cnt[op] += n
    for s_ in stalls:
        try: v = int(r[ix[s_]])
        except Exception: v = 0
        tot[s_] += v; byop[op][s_] += v
T = sum(tot.values()) or 1; N = sum(cnt.values()) or 1
print("samples", T, "warp instructions", N)
print("stall mix:", {k[6:]: round(100 * v / T, 1) for k, v in tot.most_common(9)})
for op, c in cnt.most_common(top):
    sm = sum(byop[op].values())
    print("%-10s instr %5.1f%%  samples %5.1f%%  top: %s" % (op, 100 * c / N, 100 * sm / T, ", ".join("%s %.1f" % (k[6:], 100 * v / T) for k, v in byop[op].most_common(3))))
