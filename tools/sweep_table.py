"""Tabulate a tools/shape_sweep.py log:  python tools/sweep_table.py gpurun_out/x_sweep.jsonl"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
tab, shas = {}, {}
for r in rows:
    if "error" in r:
        print("ERROR", r["library"].split("/")[-1], r["interp"], r["fr"], r["ctas"], r["error"][:80])
        continue
    tab.setdefault((r["interp"], r.get("dtype", "u8"), r["library"]), {})[(r["fr"], r["ctas"] * 10 + r.get("teams", 0), r["pct"])] = r
    shas.setdefault((r["interp"], r.get("dtype", "u8")), set()).add(r["sha"])
for (i, dt, l), d in sorted(tab.items()):
    print(i, dt, l)
    for fr in sorted({k[0] for k in d}):
        cells = []
        for (f, c, p), r in sorted(d.items()):
            if f != fr:
                continue
            st = r.get("stats")
            cells.append("c%d/t%d/p%d:%6.1f%s" % (c // 10, c % 10, p, r["Gpix_per_s"],
                                              "" if not st else " (cw %.2f pw %.2f multi %d/%d)" % (
                                                  st["consumer_wait_frac"], st["producer_wait_frac"], st["multi_slots"], st["slots"])))
        print("   fr", fr, "  ".join(cells))
print({k: sorted(v) for k, v in shas.items()})
