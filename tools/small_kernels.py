"""CUDA-event timings and achieved HBM GB/s of the per-ray kernels (compositing, resampling) on the bench workload's
shapes.  Algorithmic bytes per ray are SURVEY.md §8(d)'s: coarse composite 1300 B, sample_pdf 776 B, fine composite 3852 B."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_tiny_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
ctx = ops.Context(0); dev = ctx.dev
g = torch.Generator(device=dev).manual_seed(0)
near = torch.full((n,), 2.0, device=dev); far = torch.full((n,), 6.0, device=dev)
t_c = ctx.sample_coarse(near, far)
rgb_c = torch.rand(n, 64, 3, device=dev, generator=g); sig_c = torch.rand(n, 64, device=dev, generator=g) * 3
rgb_f = torch.rand(n, 128, 3, device=dev, generator=g); sig_f = torch.rand(n, 128, device=dev, generator=g) * 3
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    ms = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return sum(ms) / len(ms), out


res = {}
ms, (w_c, c_c) = timed(lambda: ctx.composite_coarse(near, far, rgb_c, sig_c))
res["composite_coarse"] = (ms, 1300)
ms, out = timed(lambda: ctx.sample_pdf(t_c, w_c))
t_f = out[0] if isinstance(out, tuple) else out
res["sample_pdf"] = (ms, 776)
ms, _ = timed(lambda: ctx.composite_fine(t_c, rgb_c, sig_c, t_f, rgb_f, sig_f, want_perm=False))
res["composite_fine (render)"] = (ms, 3852)
ms, _ = timed(lambda: ctx.composite_fine(t_c, rgb_c, sig_c, t_f, rgb_f, sig_f, want_perm=True))
res["composite_fine (train, permutations kept)"] = (ms, 3852 + 960)
for k, (ms, bpr) in res.items():
    print(f"{k:45s} {ms*1e3:9.1f} us   {n * bpr / (ms * 1e-3) / 1e9:8.1f} GB/s algorithmic")
print(json.dumps({k: {"us": ms * 1e3, "GBps": n * bpr / (ms * 1e-3) / 1e9} for k, (ms, bpr) in res.items()}))
