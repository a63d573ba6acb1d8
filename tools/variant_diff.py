"""Kernel experiments: where two builds of the library disagree.  Runs the bench workload's first frames through the
in-tree library and through R360_LIBRARY_B (a variant) in two processes' worth of state -- here simply by loading the
second library with ctypes under another name is not possible (one process, one libremap360), so the tool is run
twice and compares dumps:

    python tools/variant_diff.py dump out_a.pt [--interp linear]            (R360_LIBRARY selects the build)
    python tools/variant_diff.py diff out_a.pt out_b.pt"""
import pathlib
import sys

import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "360cam-pgm-3dgs-tools_b200"))
sys.path.insert(0, str(ROOT))


def dump(path, interp):
    import remap360
    from bench import preset_views
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    src = torch.randint(0, 256, (2, 3840, 7680, 3), dtype=torch.uint8, device="cuda", generator=g)
    views = [remap360.PerspectiveView(y, p, hf, vf) for _, y, p, hf, vf in preset_views("full360coverage", 1600)]
    out = remap360.remap_erp(src, views, (1600, 1600), interp=interp)
    torch.save(out.cpu(), path)


def diff(a, b):
    x, y = torch.load(a).to(torch.int16), torch.load(b).to(torch.int16)
    d = (x - y).abs()
    print("differing elements:", int((d > 0).sum()), "of", d.numel(), "max", int(d.max()))
    for f in range(d.shape[0]):
        for v in range(d.shape[1]):
            n = int((d[f, v] > 0).sum())
            if n:
                idx = (d[f, v] > 0).nonzero()
                print("frame", f, "view", v, "n", n, "rows", int(idx[:, 0].min()), int(idx[:, 0].max()), "cols",
                      int(idx[:, 1].min()), int(idx[:, 1].max()), "first", idx[:5].tolist())


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2], sys.argv[4] if len(sys.argv) > 4 else "linear")
    else:
        diff(sys.argv[2], sys.argv[3])
