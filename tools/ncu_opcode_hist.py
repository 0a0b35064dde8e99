#!/usr/bin/env python
"""Opcode histogram and hottest stall lines of one kernel out of an ncu report (needs --import-source on):
tools/ncu_opcode_hist.py report.ncu-rep kernel-regex [launch-index]"""
import collections, csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip-before-match", "0"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
# several kernels: blocks start with a "Kernel Name" row
blocks = txt.split('"Kernel Name"')[1:]
blk = blocks[min(skip, len(blocks) - 1)]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
print(rows[0][1][:100], "| blocks:", len(blocks))
hdr = rows[1]
ie, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
ops, st, tot, tst = collections.Counter(), collections.Counter(), 0, 0
body = []
for r in rows[2:]:
    try:
        n = int(r[ie])
    except (ValueError, IndexError):
        continue
    w = r[isrc].split()
    op = w[1] if w[0].startswith("@") else w[0]
    op = op.split(".")[0]
    s = int(r[ist] or 0)
    ops[op] += n; tot += n; st[op] += s; tst += s
    body.append((s, n, r[isrc]))
print("warp instructions executed:", tot, " stall samples:", tst)
for op, n in ops.most_common(22):
    print("%-10s %12d %5.1f %%   stall %5.1f %%" % (op, n, 100.0 * n / tot, 100.0 * st[op] / max(tst, 1)))
print("-- hottest lines (stall samples, executed, SASS)")
for s, n, src in sorted(body, key=lambda x: -x[0])[:18]:
    print("%6d %12d  %s" % (s, n, src[:100]))
