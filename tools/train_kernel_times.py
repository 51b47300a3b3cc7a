"""Per-kernel breakdown of one training step via the torch profiler (kernel durations summed by name), plus the
CUDA-graph replay time per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from nerf_tiny_b200 import nerf, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
nerf.seed_everything(624)
m = nerf.NeRFModel(64, 128, batch_ray=n, precision="bf16").to(dev)
m.check_range = False
opt = nerf.FusedAdam(m, lr=3e-4)
rows17 = synth.pose_rows(8, 400, 400, synth.focal_of(400))
kinv = synth.k_inv_of(400, 400, synth.focal_of(400))
b = synth.random_batch(rows17, n, 400, 400, torch.Generator().manual_seed(1))
m.train()
gs = nerf.GraphedTrainStep(m, opt, n, kinv)
for i in range(5):
    gs(b[0], b[1], b[2], b[3])
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(20):
    gs(b[0], b[1], b[2], b[3])
ev[1].record()
torch.cuda.synchronize()
print("graph ms/step %.4f" % (ev[0].elapsed_time(ev[1]) / 20))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(5):
        nerf.train_step(m, opt, b[0], b[1], b[2], b[3], kinv)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows) / 5
print("kernel time per step %.1f us" % tot)
for e in rows[:24]:
    print("%8.1f us  x%-3d %s" % (e.device_time_total / 5, e.count // 5, e.key[:90]))
