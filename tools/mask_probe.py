import sys, json, pathlib, torch
ROOT = pathlib.Path("/root/repo")
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200")); sys.path.insert(0, str(ROOT))
import remap360
from remap360 import dualfisheye as dfh
sensors, _ = dfh.load_metashape_calibration(ROOT / "360cam-pgm-3dgs-tools_b200/remap360/templates/Osmo360-Fisheye-Distortion.xml")
cal = sensors["0"]
specs = dfh.build_sfm10_specs(1750, 14.0, "36 36", 40.0, 40.0)
views, cals, info = dfh.choose_lenses(cal, cal, specs)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
for C, interp in ((3,"cubic"),(3,"nearest"),(1,"nearest"),(1,"linear")):
    pair = torch.randint(0,256,(8,2,3840,3840,C),dtype=torch.uint8,device="cuda")
    out = remap360.alloc_views(8, len(views), 1750, 1750, C, torch.uint8, "cuda")
    ms = t(lambda: remap360.remap_fisheye(pair, cals, views, (1750,1750), interp=interp, out=out))
    print(json.dumps({"C":C,"interp":interp,"ms":ms,"Gpix_per_s":8*len(views)*1750*1750/ms/1e6}))
