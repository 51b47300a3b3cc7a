"""Kernel-only timing of sample_pdf (preallocated outputs, CUDA events over 20 launches, idx output off like the render path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_tiny_b200 import ops, _lib
from nerf_tiny_b200.ops import _p, _st
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
ctx = ops.Context(0); dev = ctx.dev
g = torch.Generator(device=dev).manual_seed(0)
near = torch.full((n,), 2.0, device=dev); far = torch.full((n,), 6.0, device=dev)
t_c = ctx.sample_coarse(near, far)
for name, scale in (("dense sigma (rand*3)", 3.0), ("thin sigma (rand*0.03, the init regime)", 0.03)):
    rgb_c = torch.rand(n, 64, 3, device=dev, generator=g); sig_c = torch.rand(n, 64, device=dev, generator=g) * scale
    w_c, _ = ctx.composite_coarse(near, far, rgb_c, sig_c)
    t_f = torch.empty(n, ctx.nf, device=dev)
    for _ in range(3):
        _lib.check(ctx.lib.nt_sample_pdf(ctx.h, n, _p(t_c), _p(w_c), None, _p(t_f), None, _st()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        _lib.check(ctx.lib.nt_sample_pdf(ctx.h, n, _p(t_c), _p(w_c), None, _p(t_f), None, _st()))
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    print("%s | %s: %.1f us  %.0f GB/s algorithmic (776 B/ray)  checksum %.6f" % (os.environ.get("NT_LIB_PATH", "default").split("/")[-1], name, us, n * 776 / us / 1e3, float(t_f.double().sum())))
