#!/usr/bin/env python
"""Join the per-instruction stall samples of an ncu report with the source lines of the kernel.

  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [--top N] [--nth N] [--by-inst]

ncu's CSV source page is SASS-only; nvdisasm -g prints the same instructions with `//## File ..., line N`
markers.  Both list the kernel's instructions in address order, so they are matched by position.
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so_path, kernel, n_inst=None):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so_path)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    for cub in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
        txt = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout
        out, cur, inside, cands = [], None, False, []
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                if out:
                    cands.append(out)
                    out = []
                inside = kernel in m.group(1)
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                out.append(cur)
        if out:
            cands.append(out)
        # several template instantiations may match the name: take the one with the report's instruction count
        for c in cands:
            if n_inst is None or len(c) == n_inst:
                return c
        if cands:
            return cands[0]
    return []


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    by_inst = "--by-inst" in sys.argv   # rank the source lines by executed warp instructions instead of stall samples
    so = os.environ.get("UVS_SO", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "uv-slam_b200", "csrc", "libuvs_b200.so"))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    # the report may hold several launches / instantiations of the kernel: --nth N picks the N-th block (default 0)
    nth = int(sys.argv[sys.argv.index("--nth") + 1]) if "--nth" in sys.argv else 0
    hdr_i = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][nth]
    print(" ".join(rows[hdr_i - 1][:2]) if hdr_i > 0 else "")
    hdr = rows[hdr_i]
    ws, si = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source")
    ie = hdr.index("Instructions Executed")
    inst = []
    for r in rows[hdr_i + 1:]:
        if len(r) != len(hdr) or r[0] == "Address":
            break
        inst.append((int(r[ws] or 0), r[si].strip(), int(r[ie] or 0)))
    lines = sass_lines(so, re.sub(r'[^A-Za-z0-9_].*', '', kernel), len(inst))
    if len(lines) != len(inst):
        print("warning: %d SASS instructions in the report, %d in the cubin (stale build?)" % (len(inst), len(lines)))
    per, ops = {}, {}
    for k, (s, src, n) in enumerate(inst):
        key = lines[k] if k < len(lines) else None
        s = n if by_inst else s
        per[key] = per.get(key, 0) + s
        ops.setdefault(key, {})
        op = src.split()[0] if src else "?"
        if op.startswith("@"):
            op = src.split()[1]
        ops[key][op] = ops[key].get(op, 0) + s
    tot = sum(per.values()) or 1
    srcs = {}
    print("total %s %d over %d instructions" % ("warp instructions executed" if by_inst else "samples", tot, len(inst)))
    for key, s in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
        text = ""
        if key:
            f = key[0]
            if f not in srcs:
                p = os.path.join(os.path.dirname(so), f)
                srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
            if 0 < key[1] <= len(srcs[f]):
                text = srcs[f][key[1] - 1].strip()
        hot = ", ".join("%s %d" % kv for kv in sorted(ops[key].items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%%  %-22s %-90s [%s]" % (100.0 * s / tot, "%s:%d" % key if key else "?", text[:90], hot))


if __name__ == "__main__":
    main()
