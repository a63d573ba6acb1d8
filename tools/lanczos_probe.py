import sys, json, pathlib, torch, os
ROOT = pathlib.Path("/root/repo")
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200")); sys.path.insert(0, str(ROOT))
import remap360
from bench import preset_views
views = [remap360.PerspectiveView(y, p, hf, vf) for _, y, p, hf, vf in preset_views("full360coverage", 1600)]
src = torch.randint(0, 256, (2, 3840, 7680, 3), dtype=torch.uint8, device="cuda")
out = torch.empty((2, 12, 1600, 1600, 3), dtype=torch.uint8, device="cuda")
fn = lambda: remap360.remap_erp(src, views, (1600, 1600), interp="lanczos4", out=out)
for _ in range(3): fn()
torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): fn()
e1.record(); torch.cuda.synchronize(); ms=e0.elapsed_time(e1)/5
print(json.dumps({"ctas": os.environ.get("R360_TILED_CTAS_PER_SM"), "ring_kb": os.environ.get("R360_RING_KB"), "ms": ms, "Gpix_per_s": out.numel()/3/ms/1e6}))
