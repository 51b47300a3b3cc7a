"""Chaos spread of the training PSNR: the same 250-step run with the initial weights perturbed by 1e-6 relative."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nerf_tiny_b200 import nerf, synth
from oracle import nerf_oracle as O
dev = torch.device("cuda", 0)
h = w = 100
f = synth.focal_of(w); rows17 = synth.pose_rows(8, h, w, f); k_inv = synth.k_inv_of(h, w, f)
for prec in ("fp32", "bf16"):
    outs = []
    for trial in range(6):
        sd = O.init_state_dict(624)
        if trial > 0:
            g2 = torch.Generator().manual_seed(trial)
            sd = {k: v * (1 + 1e-6 * torch.randn(v.shape, generator=g2)) for k, v in sd.items()}
        m = nerf.NeRFModel(64, 128, batch_ray=256, precision=prec); m.load_state_dict(sd); m = m.to(dev); m.train(); m.check_range = False
        opt = nerf.FusedAdam(m, lr=1e-3)
        gen = torch.Generator().manual_seed(123)
        ps = []
        for it in range(250):
            row, col, pix, pb, pic = synth.random_batch(rows17, 256, h, w, gen)
            loss, cc, cf = nerf.train_step(m, opt, row, col, pix, pb, k_inv)
            ps.append(-10.0 * torch.log10(torch.mean(torch.square(cf - pix.to(dev)))))
        outs.append(float(torch.stack(ps)[-50:].mean()))
    print("SPREAD", prec, ["%.3f" % x for x in outs], "mean %.3f std %.3f" % (np.mean(outs), np.std(outs)))
