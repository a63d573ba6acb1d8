#!/usr/bin/env python3
"""Shared-memory wavefronts per SASS opcode class and per CUDA line from an ncu report.

    python tools/ncu_smem.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, cur, cur_file = None, None, None
    by_op, by_line = defaultdict(lambda: [0, 0, 0]), defaultdict(lambda: [0, 0, 0])
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r
            iw, ii, ie = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("Instructions Executed")
            continue
        if hdr is None or r[0] == "Function Name":
            continue
        if r[0] != "":
            cur = (cur_file, int(r[0]), r[1].strip()[:80])
        elif cur is not None and len(r) > iw:
            try:
                w, i, e = int(r[iw]), int(r[ii]), int(r[ie])
            except ValueError:
                continue
            if w == 0:
                continue
            toks = [t for t in r[1].split() if not t.startswith("@")]
            op = toks[0] if toks else "?"
            for d, k in ((by_op, op), (by_line, cur)):
                d[k][0] += w; d[k][1] += i; d[k][2] += e
    tot = sum(v[0] for v in by_op.values()) or 1
    print("total shared wavefronts %d" % tot)
    for k, (w, i, e) in sorted(by_op.items(), key=lambda kv: -kv[1][0]):
        print("%6.2f%%  %-14s wave %11d ideal %11d exec %10d  wave/exec %.2f" % (100.0 * w / tot, k, w, i, e, w / max(e, 1)))
    print()
    for k, (w, i, e) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f%%  wave/exec %5.2f  %s:%d  %s" % (100.0 * w / tot, w / max(e, 1), k[0], k[1], k[2]))


if __name__ == "__main__":
    main()
