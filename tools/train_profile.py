"""Runs a few training steps (cfg3 shape: 1024 rays) — target of the ncu launch list for the train path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_tiny_b200 import nerf, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
nerf.seed_everything(624)
m = nerf.NeRFModel(64, 128, batch_ray=n, precision="bf16").to(dev)
m.check_range = False
opt = nerf.FusedAdam(m, lr=3e-4)
rows17 = synth.pose_rows(8, 400, 400, synth.focal_of(400))
kinv = synth.k_inv_of(400, 400, synth.focal_of(400))
b = synth.random_batch(rows17, n, 400, 400, torch.Generator().manual_seed(1))
m.train()
for i in range(steps):
    loss, _, _ = nerf.train_step(m, opt, b[0], b[1], b[2], b[3], kinv)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(steps):
    loss, _, _ = nerf.train_step(m, opt, b[0], b[1], b[2], b[3], kinv)
ev[1].record()
torch.cuda.synchronize()
print("ms/step", ev[0].elapsed_time(ev[1]) / steps, "loss", float(loss))
