"""Diagnostic: where do two schedules of the fused MLP kernel disagree (sample indices)?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nerf_tiny_b200 import ops, synth
from oracle import nerf_oracle as O
ctx = ops.Context(0); dev = ctx.dev
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2501
va, vb = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (5, 7)
rows17 = synth.pose_rows(8, 400, 400, synth.focal_of(400))
row, col, pix, pb, pic = synth.random_batch(rows17, n, 400, 400, torch.Generator().manual_seed(0))
flat = ops.flatten_state_dict(O.init_state_dict(624), dev)
packed = ctx.pack(flat, 2)
rays, _, de = ctx.raygen(row.to(dev), col.to(dev), pb.float().to(dev), synth.k_inv_of(400, 400, synth.focal_of(400)).to(dev))
t = (torch.rand(n, 128, device=dev) * 4 + 2)
out = {}
for v in (va, vb):
    ctx.set_option(2, v)
    for rep in range(3):
        rgb, sig, _ = ctx.mlp_forward(2, t, rays, de, flat, packed)
        torch.cuda.synchronize()
        out[(v, rep)] = (rgb.clone(), sig.clone())
for rep in range(3):
    d = (out[(va, 0)][0] - out[(vb, rep)][0]).abs().amax(-1).reshape(-1)
    bad = torch.nonzero(d > 2e-3).reshape(-1).cpu().numpy()
    tiles = np.unique(bad // 128)
    print(f"rep {rep}: max {float(d.max()):.4g}  bad samples {bad.size}  tiles {tiles[:20]} (of {tiles.size})  pairs {np.unique(tiles//2)[:12]}  tl {np.unique(tiles%2)}")
    if bad.size:
        print("   rows in tile", np.unique(bad % 128)[:40])
d2 = (out[(vb, 0)][0] - out[(vb, 1)][0]).abs().max()
print("run-to-run", float(d2))
