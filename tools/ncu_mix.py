#!/usr/bin/env python
"""Instruction mix (by SASS opcode) and stall samples of one kernel of an ncu report:  python tools/ncu_mix.py REPORT KERNEL_SUBSTRING"""
import collections, csv, io, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, h, idx, done = None, None, None, set()
cat, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name":
        cur = r[1]; continue
    if r[0] == "Address":
        h = r; idx = {k: i for i, k in enumerate(h)}
        if cur in done: cur = None   # first launch of a kernel only
        elif cur and want in cur: done.add(cur)
        continue
    if cur is None or want not in cur or len(r) < len(h): continue
    t = r[idx["Source"]].strip().split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    n, s = int(r[idx["Instructions Executed"]]), int(r[idx["# Samples"]])
    cat[op] += n; samp[op] += s; tot += n
print("kernel(s):", done); print("total warp instructions", tot, "samples", sum(samp.values()))
for k, v in cat.most_common(22): print("%-10s %10d %5.1f%%  samples %6d" % (k, v, 100 * v / max(tot, 1), samp[k]))
