"""Raw pinned-memory copy bandwidth of the box (one direction and both at once) next to what the
streaming remapper achieves: tells how close the e2e number is to the PCIe link."""
import json
import torch

def bw(fn, nbytes, iters=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9

def main():
    n = 1 << 30
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {"h2d_GBs": bw(lambda: d_in.copy_(h_in, non_blocking=True), n),
           "d2h_GBs": bw(lambda: h_out.copy_(d_out, non_blocking=True), n)}
    def both():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
    res["both_each_GBs"] = bw(both, n)
    print(json.dumps(res))

if __name__ == "__main__":
    main()
