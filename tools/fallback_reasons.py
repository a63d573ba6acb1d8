"""Why tiles are left to the float64 direct path: histogram of the reason the plan kernel records in every fallback
tile's record (TilePlan.pad[0]), per view, for the bench's seam / pole view set at 8K (B200, under gpurun).

    python tools/fallback_reasons.py [cubic|linear] [u8|u16]"""
import collections
import json
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200"))
sys.path.insert(0, str(ROOT))
import remap360  # noqa: E402
from remap360 import api  # noqa: E402
import bench  # noqa: E402

REASONS = {0: "-", 1: "polynomial does not fit", 2: "patch larger than the ring budget", 3: "columns outside one period, no seam path",
           4: "rows outside the sensor / invalid pixels", 5: "tile spans >= 2048 source px", 6: "map pool full"}


def main():
    interp = sys.argv[1] if len(sys.argv) > 1 else "cubic"
    dtype = {"u8": torch.uint8, "u16": torch.uint16}[sys.argv[2] if len(sys.argv) > 2 else "u8"]

    class NS:
        preset, size, frames, interp = "full360coverage", 1600, 4, "cubic"
    w = dict(bench.workload_table(NS)["cfg2_seam_pole_views"], interp=interp)
    views = bench.workload_views(w)
    pviews = [remap360.PerspectiveView(y, p, hf, vf, view_id=vid) for vid, y, p, hf, vf, _ in views]
    src = torch.zeros((4, 3840, 7680, 3), dtype=dtype, device="cuda")
    out = remap360.remap_erp(src, pviews, (1600, 1600), interp=interp)
    torch.cuda.synchronize()
    del out
    plan = list(api._PLAN_CACHE.values())[-1]
    ws = plan.workspace.cpu().numpy().view(np.uint8)
    n_tiles = plan.tiles_per_view
    # workspace layout (remap360.cu, workspace_layout): header | views | plan records | fallback list | order list |
    # pool of per-pixel maps (16 KB each, one tile in sixteen + 32)
    al = lambda n: (n + 255) // 256 * 256
    nt = n_tiles * len(pviews)
    pool = min(nt // 16 + 32, 1 << 20) * 16384
    off = ws.size - pool - 2 * al(8 * nt) - al(368 * nt)
    rec = ws[off:off + 368 * n_tiles * len(pviews)].reshape(len(pviews), n_tiles, 368)
    ints = rec[:, :, 336:368].copy().view(np.int32)          # py0 rows xb0 row_bytes pitch mode_slot pad0 pad1
    mode, reason, cmap = ints[..., 5] & 0xff, ints[..., 6], ints[..., 7]
    MODES = {0: "fallback", 1: "boxes", 2: "fill", 3: "rows", 4: "seam"}
    rows, pitch = ints[..., 1], ints[..., 4]
    staged = np.where(mode == 1, (rows // 32) * 32 + ((rows % 32) + 7) // 8 * 8, rows) * pitch
    for v, (vid, yaw, pch, *_rest) in enumerate(views):
        print(json.dumps({"view": vid, "yaw": yaw, "pitch": pch, "tiles_by_mode": {MODES[m]: int((mode[v] == m).sum()) for m in MODES if (mode[v] == m).any()},
                          "per_pixel_map_tiles": int((cmap[v] > 0).sum()),
                          "patch_KB_mean": round(float(staged[v][mode[v] != 0].mean()) / 1024, 1),
                          "patch_KB_max": round(float(staged[v][mode[v] != 0].max()) / 1024, 1),
                          "patch_KB_by_mode": {MODES[m]: round(float(staged[v][mode[v] == m].mean()) / 1024, 1) for m in (1, 3, 4) if (mode[v] == m).any()}}))
    total = collections.Counter()
    for v, (vid, yaw, pitch, *_rest) in enumerate(views):
        fb = mode[v] == 0
        c = collections.Counter(int(r) for r in reason[v][fb])
        total.update(c)
        if fb.any():
            print(json.dumps({"view": vid, "yaw": yaw, "pitch": pitch, "fallback_tiles": int(fb.sum()), "of": n_tiles,
                              "reasons": {REASONS.get(k, str(k)): n for k, n in sorted(c.items())}}))
    print(json.dumps({"interp": interp, "all_views": len(views), "fallback_tiles": int((mode == 0).sum()), "tiles": int(mode.size),
                      "per_pixel_map_tiles": int((cmap > 0).sum()), "plan_n_map_tiles": plan.n_map_tiles, "plan_n_fallback": plan.n_fallback,
                      "reasons": {REASONS.get(k, str(k)): n for k, n in sorted(total.items())},
                      "patch_budget_note": "ring budget of the single-frame plan; multi-frame launches use the same plan"}))


if __name__ == "__main__":
    main()
