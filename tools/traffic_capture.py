"""DRAM traffic of one launch of the tiled kernel for every workload of bench.py's table (B200, under gpurun).

For each workload: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` around a
three-step bench.py run of that workload alone, last launch of remap_tiled_kernel.  Writes
profiles/r02_dram_traffic.json with the digest of csrc/ so that bench.py can tell a stale capture from a current one.

    python tools/traffic_capture.py [workload ...]"""
import csv
import json
import pathlib
import subprocess
import sys
import tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

METRICS = "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"


def main():
    class NS:
        preset, size, frames, interp = "full360coverage", 1600, 16, "cubic"
    table = bench.workload_table(NS)
    names = sys.argv[1:] or list(table)
    out = {"csrc_digest": bench.csrc_digest(), "metrics": METRICS, "tool": "ncu --clock-control none, last remap_tiled_kernel launch of "
           "`bench.py --workload NAME --steps 1 --warmup 2 --no-variants --no-e2e --no-cpu-baseline`", "workloads": {}}
    for name in names:
        w = table[name]
        with tempfile.NamedTemporaryFile(suffix=".csv") as tmp:
            cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "-k", "regex:remap_tiled", "--csv", "--log-file", tmp.name,
                   sys.executable, str(ROOT / "bench.py"), "--workload", name, "--steps", "1", "--warmup", "2", "--no-variants", "--no-e2e",
                   "--no-cpu-baseline"]
            rc = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600).returncode
            rows = [r for r in csv.reader(l for l in open(tmp.name) if not l.startswith("=="))]
        if rc != 0 or len(rows) < 2:
            out["workloads"][name] = {"error": "ncu run failed (rc %d)" % rc, "interp": w["interp"], "frames": w["frames"],
                                      "dram_bytes_read": 0, "dram_bytes_write": 0}
            continue
        head = rows[0]
        recs = [dict(zip(head, r)) for r in rows[1:]]
        # the dominant kernel of a step: the main walk (a short large-patch pass may follow it) -- the last launch
        # whose duration is at least half the longest one's
        dur = {r["ID"]: float(r["Metric Value"].replace(",", "")) for r in recs if r["Metric Name"] == "gpu__time_duration.sum"}
        last_id = [i for i in dur if dur[i] >= 0.5 * max(dur.values())][-1]
        vals = {r["Metric Name"]: float(r["Metric Value"].replace(",", "")) for r in recs if r["ID"] == last_id}
        units = {r["Metric Name"]: r["Metric Unit"] for r in recs if r["ID"] == last_id}
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = vals["dram__bytes_read.sum"] * scale.get(units["dram__bytes_read.sum"], 1)
        wr = vals["dram__bytes_write.sum"] * scale.get(units["dram__bytes_write.sum"], 1)
        per_frame, _u = bench.algorithmic_bytes(w, len(bench.workload_views(w)))
        alg = per_frame * w["frames"] if per_frame else None
        out["workloads"][name] = {"interp": w["interp"], "frames": w["frames"], "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
                                  "gpu_time_ns_under_ncu": vals["gpu__time_duration.sum"], "launches_seen": len({r["ID"] for r in recs}),
                                  "kernel": next(r["Kernel Name"] for r in recs if r["ID"] == last_id)[:120], "algorithmic_bytes": alg,
                                  "traffic_over_algorithmic": round((rd + wr) / alg, 3) if alg else None}
        print(name, json.dumps(out["workloads"][name]), flush=True)
    dest = ROOT / "gpurun_out" / "r02_dram_traffic.json"
    dest.parent.mkdir(exist_ok=True)
    # a partial run (workload names given) updates the committed file when the sources are still the same
    if sys.argv[1:] and bench.TRAFFIC_FILE.exists():
        old = json.loads(bench.TRAFFIC_FILE.read_text())
        if old.get("csrc_digest") == out["csrc_digest"]:
            old["workloads"].update(out["workloads"])
            out = old
    dest.write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
