"""Run-to-run bit comparison of the cta_group::2 MLP kernel over many launches (constant-memory bias slots rotate per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_tiny_b200 import ops, synth
from oracle import nerf_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
ctx = ops.Context(0); dev = ctx.dev
rows17 = synth.pose_rows(8, 400, 400, synth.focal_of(400))
row, col, pix, pb, pic = synth.random_batch(rows17, n, 400, 400, torch.Generator().manual_seed(0))
rays, _, de = ctx.raygen(row.to(dev), col.to(dev), pb.float().to(dev), synth.k_inv_of(400, 400, synth.focal_of(400)).to(dev))
t = (torch.rand(n, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 4 + 2)
for prec in (3, 2):
    ref = None
    bad = 0
    for seed in (624, 7):
        flat = ops.flatten_state_dict(O.init_state_dict(seed), dev)
        packed = ctx.pack(flat, prec)
        ref = None
        for r in range(reps):
            if r % 5 == 4:
                packed = ctx.pack(flat, prec)     # re-pack in place now and then, like NeRFModel.forward does every call
            rgb, sig, _ = ctx.mlp_forward(prec, t, rays, de, flat, packed)
            torch.cuda.synchronize()
            if ref is None:
                ref = (rgb.clone(), sig.clone())
            else:
                d = max(float((rgb - ref[0]).abs().max()), float((sig - ref[1]).abs().max()))
                if d != 0.0:
                    bad += 1
                    print(f"prec {prec} seed {seed} rep {r}: max diff {d:.3e}, samples differing {int(((rgb - ref[0]).abs().amax(-1) > 0).sum())}")
    print(f"prec {prec}: {bad} of {2 * (reps - 1)} launches differ from the first")
