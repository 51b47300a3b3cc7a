"""Same-box timing of the fused MLP kernel (fine pass of the bench workload: 160000 rays x 128 samples, fp16 operands):
CUDA events around single launches, L2 flushed in between.  Run once per NT_LIB_PATH variant inside ONE gpurun call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nerf_tiny_b200 import ops, synth
from oracle import nerf_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = ops.Context(0); dev = ctx.dev
rows17 = synth.pose_rows(8, 400, 400, synth.focal_of(400))
row, col, pix, pb, pic = synth.random_batch(rows17, n, 400, 400, torch.Generator().manual_seed(0))
flat = ops.flatten_state_dict(O.init_state_dict(624), dev)
packed = ctx.pack(flat, prec)
rays, _, de = ctx.raygen(row.to(dev), col.to(dev), pb.float().to(dev), synth.k_inv_of(400, 400, synth.focal_of(400)).to(dev))
t = (torch.rand(n, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 4 + 2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ctx.mlp_forward(prec, t, rays, de, flat, packed)
ms = []
for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); rgb, sig, _ = ctx.mlp_forward(prec, t, rays, de, flat, packed); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
ms.sort()
print("%s: median %.3f ms  min %.3f  max %.3f  checksum %.6f" % (os.environ.get("NT_LIB_PATH", "default").split("/")[-1],
      ms[len(ms) // 2], ms[0], ms[-1], float(rgb.double().sum() + sig.double().sum())))
