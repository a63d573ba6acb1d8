"""Debug helper: compare tiled vs direct output for one view and list mismatching tiles with their plan records."""
import sys, pathlib
import numpy as np, torch
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200"))
import remap360
from remap360 import api

yaw, pitch = float(sys.argv[1]), float(sys.argv[2])
interp = sys.argv[3] if len(sys.argv) > 3 else "linear"
W, H, size = 7680, 3840, 1600
g = torch.Generator(device="cuda"); g.manual_seed(1)
src = torch.randint(0, 256, (1, H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
v = [remap360.PerspectiveView(yaw, pitch, 104.2500326978036, 104.2500326978036)]
a = remap360.remap_erp(src, v, (size, size), interp=interp, path="direct")
b = remap360.remap_erp(src, v, (size, size), interp=interp, path="tiled")
d = (a.to(torch.int16) - b.to(torch.int16)).abs().amax(dim=-1)[0, 0]
bad = (d > 1)
print("bad pixels", int(bad.sum()), "of", bad.numel())
tiles = bad.view(50, 32, 50, 32).any(dim=3).any(dim=1)
idx = tiles.nonzero().cpu().numpy()
print("bad tiles", len(idx))
# plan records
srcd = api._describe(src, "frames"); dstd = api._describe(b.view(1, size, size, 3), "out")
plan = api.get_plan(srcd, dstd, v, api._options(interp, path="tiled"), src.device)
ws = plan.workspace.cpu().numpy()
off_plans = 256 + 256            # header + 1 view (80 B -> 256)
rec = np.frombuffer(ws[off_plans:off_plans + 2500 * 368].tobytes(), dtype=np.int32).reshape(2500, 92)
modes = rec[:, 89] & 0xff
print("mode histogram", np.bincount(modes, minlength=4), "n_fallback", plan.n_fallback)
order = sorted(idx.tolist(), key=lambda t: -int(bad[t[0]*32:(t[0]+1)*32, t[1]*32:(t[1]+1)*32].sum()))
for tj, ti in order[:24]:
    r = rec[tj * 50 + ti]
    cnt = int(bad[tj*32:(tj+1)*32, ti*32:(ti+1)*32].sum())
    aff = r[72:84].view(np.float64)
    print("tile", (int(ti), int(tj)), "bad px", cnt, "py0,rows", r[84], r[85], "xb0,row_bytes,pitch", r[86], r[87], r[88],
          "mode", r[89] & 0xff, "wbox", r[89] >> 16, "ax", np.round(aff[:3] / 32, 3), "ay", np.round(aff[3:] / 32, 3))
