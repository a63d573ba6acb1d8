#!/usr/bin/env python3
"""Aggregate an ncu report's executed warp-instructions per CUDA source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]

Needs a report captured with --import-source on from a -lineinfo build."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, per_line, cur = None, None, {}, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_exec = hdr.index("Instructions Executed")
            i_samp = hdr.index("# Samples")
            continue
        if hdr is None or r[0] in ("Function Name",):
            continue
        if r[0] != "":                       # a CUDA source line header
            cur = (cur_file, int(r[0]), r[1].strip())
            per_line.setdefault(cur, [0, 0, 0])
        elif cur is not None and len(r) > i_exec:
            try:
                per_line[cur][0] += int(r[i_exec]); per_line[cur][1] += int(r[i_samp]); per_line[cur][2] += 1
            except ValueError:
                pass
    total = sum(v[0] for v in per_line.values()) or 1
    tsamp = sum(v[1] for v in per_line.values()) or 1
    print("total executed warp-instructions: %d, samples %d" % (total, tsamp))
    for (f, ln, src), (ex, sm, n) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%6.2f%% exec %5.2f%% samp %4d sass  %s:%d  %s" % (100.0 * ex / total, 100.0 * sm / tsamp, n, f, ln, src[:90]))


if __name__ == "__main__":
    main()
