#!/usr/bin/env python
"""Per-kernel SASS evidence from the built library (no GPU needed):

    python benchmarks/sass_excerpts.py > profiles/r02_sass_excerpts.txt

For every kernel in libfrcnn_b200.so: instruction count, the Blackwell / Hopper+-specific mnemonics it contains
(packed fp32 FADD2 / FMUL2 / FFMA2, cp.async LDGSTS + LDGDEPBAR, 1-D bulk copy UBLKCP, mbarrier SYNCS, cluster
barriers UCGABAR, distributed-shared-memory stores, REDUX, VIMNMX / VIADDMNMX, MATCH, ...) and the first lines that use them."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "faster_rcnn_b200", "libfrcnn_b200.so")
WATCH = ["FFMA2", "FADD2", "FMUL2", "LDGSTS", "LDGDEPBAR", "DEPBAR", "UBLKCP", "SYNCS", "UCGABAR", "ST.E.64.STRONG", "MAPA", "REDUX",
         "VIMNMX", "VIADDMNMX", "MATCH", "ATOMS", "ATOMG", "RED.E", "R2P", "STG.E.EF", "LDG.E.128.CONSTANT", "BAR.SYNC",
         "SHFL", "VOTE", "POPC", "FLO", "MUFU", "DFMA", "DMUL", "DADD"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = re.split(r"\n\s*Function : ", out)[1:]
    print("SASS summary of %s (cuobjdump -sass, sm_100a)\n" % os.path.relpath(LIB, ROOT))
    merged = collections.OrderedDict()
    for k in kernels:
        name = k.split("\n", 1)[0].strip()
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", demangled).replace("frcnn::", "").replace("void ", "")
        base = re.sub(r"<.*", "", short)
        lines = [re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip() for ln in k.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
        lines = [ln.rstrip(";").strip() for ln in lines if ln]
        merged.setdefault(base, []).append((short, lines))
    for base, variants in merged.items():
        short, lines = max(variants, key=lambda v: len(v[1]))
        ops = collections.Counter()
        first = {}
        for ln in lines:
            body = re.sub(r"^@!?U?P\d+\s+", "", ln)
            for w in WATCH:
                if body.startswith(w) or (" " + w) in (" " + body.split(" ")[0]):
                    ops[w] += 1
                    first.setdefault(w, ln)
        print("== %s   (%d template instance%s; largest: %s, %d SASS instructions)" %
              (base, len(variants), "" if len(variants) == 1 else "s", short, len(lines)))
        print("   " + ", ".join("%s x%d" % (w, ops[w]) for w in WATCH if ops[w]))
        for w in ("FFMA2", "FADD2", "FMUL2", "LDGSTS", "UBLKCP", "SYNCS", "UCGABAR", "MAPA", "REDUX", "VIMNMX", "MATCH", "R2P", "STG.E.EF"):
            if w in first:
                print("     %-10s %s" % (w, first[w][:120]))
        print()


if __name__ == "__main__":
    sys.exit(main())
