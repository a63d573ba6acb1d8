"""Patch statistics of a tile plan (bench workload: 8K -> full360coverage 12 x 1600^2): staged bytes per tile by view,
against the shared-memory ring of the remap kernel.  python tools/plan_stats.py [cubic|linear]"""
import sys, pathlib, json
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200")); sys.path.insert(0, str(ROOT))
import remap360
from remap360 import api
from bench import preset_views

interp = sys.argv[1] if len(sys.argv) > 1 else "cubic"
W, H, size = 7680, 3840, 1600
src = torch.zeros((1, H, W, 3), dtype=torch.uint8, device="cuda")
pv = preset_views("full360coverage", size)
views = [remap360.PerspectiveView(y, p, hf, vf) for _, y, p, hf, vf in pv]
out = torch.empty((1, len(views), size, size, 3), dtype=torch.uint8, device="cuda")
srcd = api._describe(src, "frames"); dstd = api._describe(out.view(len(views), size, size, 3), "out")
plan = api.get_plan(srcd, dstd, views, api._options(interp, path="tiled"), src.device)
ws = plan.workspace.cpu().numpy()
off = 256 + (len(views) * 112 + 255) // 256 * 256
n_tiles = plan.tiles_per_view
rec = np.frombuffer(ws[off:off + len(views) * n_tiles * 368].tobytes(), dtype=np.int32).reshape(len(views), n_tiles, 92)
rows, row_bytes, pitch, mode = rec[..., 85], rec[..., 87], rec[..., 88], rec[..., 89] & 0xff
staged = np.where(mode == 1, ((rows // 32) * 32 + (((rows % 32) + 7) // 8) * 8) * pitch, rows * pitch)
staged = np.where((mode == 1) | (mode == 3) | (mode == 4), staged, 0)
need = (staged + 127) // 128 * 128
print("interp", interp, "tiles/view", n_tiles, "fallback", plan.n_fallback)
for k, (vid, y, p, _, _) in enumerate(pv):
    m = mode[k]
    ok = need[k][need[k] > 0]
    print("view %-4s yaw %6.1f pitch %5.1f  modes %s  patch bytes mean %6.0f p50 %6.0f p90 %6.0f max %6.0f  rows mean %.1f pitch mean %.0f row_bytes mean %.0f"
          % (vid, y, p, np.bincount(m, minlength=5).tolist(), ok.mean(), np.percentile(ok, 50), np.percentile(ok, 90), ok.max(),
             rows[k][need[k] > 0].mean(), pitch[k][need[k] > 0].mean(), row_bytes[k][need[k] > 0].mean()))
allok = need[need > 0]
print("all views: mean %.0f B per tile = %.1f B per output pixel; useful bytes (rows x row_bytes) mean %.0f"
      % (allok.mean(), allok.mean() / 1024.0, (rows * row_bytes)[need > 0].mean()))
