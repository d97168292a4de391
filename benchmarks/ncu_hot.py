#!/usr/bin/env python
"""Hottest SASS instructions of a kernel in an .ncu-rep: python benchmarks/ncu_hot.py file.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
body = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        body = []
    elif hdr and len(r) == len(hdr):
        body.append(r)
i_src, i_smp, i_ins = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot_ins = sum(int(r[i_ins]) for r in body)
tot_smp = sum(int(r[i_smp]) for r in body)
print("total warp instructions %d, samples %d, SASS lines %d" % (tot_ins, tot_smp, len(body)))
ops = {}
for r in body:
    op = r[i_src].split()[0] if not r[i_src].strip().startswith("@") else r[i_src].split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[i_ins])
print("by opcode:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_ins) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]))
print("top by stall samples:")
for n, r in sorted(enumerate(body), key=lambda nr: -int(nr[1][i_smp]))[:top]:
    print("  line %4d  smp %5.1f%%  inst %5.1f%%  %s" % (n, 100.0 * int(r[i_smp]) / max(tot_smp, 1), 100.0 * int(r[i_ins]) / tot_ins, r[i_src].strip()[:90]))
