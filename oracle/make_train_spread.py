"""How reproducible is 'PSNR after N steps' for the UNMODIFIED reference itself?  (build container only, ~4 min/run)

    python -m oracle.make_train_spread [n_runs]

Repeats the 250-step training of oracle/make_trained_golden.py (same batches, seed 123) with the initial weights
perturbed by 1e-6 relative.  The fp32 gradient of the reference is chaotic along the t_fine path, so the
final PSNR is a distribution, not a number; tests/golden/train_spread.npz records samples of it and the GPU training
test compares means.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O          # noqa: E402
from oracle import ref_harness as RH         # noqa: E402
from nerf_tiny_b200 import synth             # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def run(ref, trial, steps=250, n_rays=256, lr=1e-3):
    h = w = 100
    f = synth.focal_of(w)
    rows17 = synth.pose_rows(8, h, w, f)
    k_inv = synth.k_inv_of(h, w, f)
    gen = torch.Generator().manual_seed(123)
    sd = O.init_state_dict(624)
    if trial > 0:
        g2 = torch.Generator().manual_seed(trial)
        sd = {k: v * (1 + 1e-6 * torch.randn(v.shape, generator=g2)) for k, v in sd.items()}
    model = RH.make_model(ref, n_rays, sd)
    opt = torch.optim.Adam([{"params": model.network.parameters(), "initial_lr": lr}], lr=lr, betas=(0.9, 0.999), eps=1e-7)
    ps = []
    for it in range(steps):
        row, col, pix, pb, pic = synth.random_batch(rows17, n_rays, h, w, gen)
        opt.zero_grad()
        model.train()
        cc, cf = model(row, col, pb, k_inv)
        loss = model.ray_loss(cc, cf, pix)
        loss.backward()
        opt.step()
        ps.append(O.psnr(cf.detach(), pix))
    return float(np.mean(ps[-50:]))


def main():
    n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    torch.set_num_threads(8)
    ref = RH.import_reference()
    z = np.load(os.path.join(GOLD, "trained_weights_fp16.npz"))
    path = os.path.join(GOLD, "train_spread.npz")
    if os.path.exists(path):                              # resume: keep the samples already measured
        samples = [float(x) for x in np.load(path)["ref_psnr_last50"]]
    else:
        samples = [float(z["__psnr"][-50:].mean())]      # trial 0 = the unperturbed run already in the fixture
    for trial in range(len(samples), n_runs + 1):
        samples.append(run(ref, trial))
        print("[spread] reference PSNR(last 50) samples:", ["%.3f" % s for s in samples], flush=True)
        np.savez(path, ref_psnr_last50=np.array(samples), perturbation=1e-6, steps=250, n_rays=256, lr=1e-3, seed=123)


if __name__ == "__main__":
    main()
