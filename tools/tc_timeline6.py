"""Diagnostic: per-layer clock64 timeline of the v6 (staggered, cluster-multicast) mlp_tc kernel, block 0."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nerf_tiny_b200 import ops, synth
from oracle import nerf_oracle as O
ctx = ops.Context(0); dev = ctx.dev
n = 148 * 2 * 6
rows17 = synth.pose_rows(8, 400, 400, synth.focal_of(400))
row, col, pix, pb, pic = synth.random_batch(rows17, n, 400, 400, torch.Generator().manual_seed(0))
flat = ops.flatten_state_dict(O.init_state_dict(624), dev)
packed = ctx.pack(flat, 2)
rays, _, de = ctx.raygen(row.to(dev), col.to(dev), pb.float().to(dev), synth.k_inv_of(400, 400, synth.focal_of(400)).to(dev))
t = (torch.rand(n, 128, device=dev) * 4 + 2)
for _ in range(2):
    rgb, sig, dbg = ctx.mlp_forward_debug(t, rays, de, flat, packed, 100 + int(os.environ.get("NT_TL_FLAGS", "0")))
torch.cuda.synchronize()
prof = dbg.cpu().numpy().view(np.int64).reshape(-1)[:4 * 10 * 16].reshape(4, 10, 16)
t0 = prof[1, 0, 0]
names = ["mmaA_start", "mmaA_end", "mmaB_start", "mmaB_end", "epiA_wake", "epiA_done", "epiB_wake", "epiB_done",
         "iss0", "iss1", "iss2", "iss3", "preActA", "rdy", "actA_ok", "actB_ok"]
NS = 8 if os.environ.get("NT_MLP_TC_VERSION") == "7" else 8
for pl in (1, 2):
    print("pair", pl)
    for L in range(10):
        r = prof[pl, L] - t0; r[13:14] = prof[pl, L, 13:14]
        print(f" L{L}: " + "  ".join(f"{names[i]}={r[i]}" for i in range(NS)))
