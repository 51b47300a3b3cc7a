"""Train the UNMODIFIED reference on the analytic scene (CPU, build container only) and store a
genuinely-trained weight state + the reference's outputs for a held-out batch.

    python -m oracle.make_trained_golden            (~5 min on 8 vCPU)   -> trained64 (250 steps)
    python -m oracle.make_trained_golden --steps 2500 --name trained2k5 --threads 4      (~1 h)  -> trained2k5

Writes tests/golden/trained_weights_fp16.npz (state_dict rounded to fp16 so the fixture stays ~1 MB; the
rounded values ARE the fixture's weights — reference and CUDA path both load exactly these) and
tests/golden/trained64.npz.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nerf_oracle as O          # noqa: E402
from oracle import ref_harness as RH         # noqa: E402
from oracle.make_golden import forward_case  # noqa: E402
from nerf_tiny_b200 import synth             # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main(steps=250, n_rays=256, lr=1e-3, name="trained64", threads=8):
    torch.set_num_threads(threads)
    wfile = "trained_weights_fp16.npz" if name == "trained64" else name + "_weights_fp16.npz"
    ref = RH.import_reference()
    h = w = 100
    f = synth.focal_of(w)
    rows17 = synth.pose_rows(8, h, w, f)
    k_inv = synth.k_inv_of(h, w, f)
    gen = torch.Generator().manual_seed(123)
    model = RH.make_model(ref, n_rays, O.init_state_dict(624))
    opt = torch.optim.Adam([{"params": model.network.parameters(), "initial_lr": lr}], lr=lr, betas=(0.9, 0.999), eps=1e-7)
    losses, psnrs = [], []
    for it in range(steps):
        row, col, pix, pb, pic = synth.random_batch(rows17, n_rays, h, w, gen)
        opt.zero_grad()
        model.train()
        cc, cf = model(row, col, pb, k_inv)                 # nerf.py:470
        loss = model.ray_loss(cc, cf, pix)                  # nerf.py:472
        loss.backward()
        opt.step()
        losses.append(float(loss))
        psnrs.append(O.psnr(cf.detach(), pix))
        if it % 500 == 499 and name != "trained64":      # long runs: keep a recoverable snapshot outside the repo history
            os.makedirs(os.path.join(ROOT, ".scratch"), exist_ok=True)
            torch.save({k: v.detach().clone() for k, v in model.state_dict().items()},
                       os.path.join(ROOT, ".scratch", f"{name}_it{it + 1}.pt"))
        if it % 25 == 0:
            print(f"[train-ref] it {it} loss {losses[-1]:.3f} psnr {psnrs[-1]:.2f}", flush=True)
    sd16 = {k: v.detach().half() for k, v in model.state_dict().items()}
    np.savez_compressed(os.path.join(GOLD, wfile), **{k: v.numpy() for k, v in sd16.items()},
                        __losses=np.array(losses), __psnr=np.array(psnrs))
    sd = {k: v.float() for k, v in sd16.items()}
    # held-out pose (not among the 8 training poses), 64 pixels
    held = synth.pose_rows(1, h, w, f)
    held[0, :15] = np.concatenate((synth.sphere_pose(0.4, 0.5), np.array([[h], [w], [f]])), axis=1).flatten()
    g2 = torch.Generator().manual_seed(5)
    row, col, pix, pb, pic = synth.random_batch(held, 64, h, w, g2)
    r = forward_case(ref, name, sd, row, col, pb, k_inv, rgb_tol=5e-6)
    print("sigma range", r["sigma_f"].min(), r["sigma_f"].max(), "psnr(last 25)", np.mean(psnrs[-25:]))


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=250)
    ap.add_argument("--name", default="trained64")
    ap.add_argument("--threads", type=int, default=8)
    a = ap.parse_args()
    main(steps=a.steps, name=a.name, threads=a.threads)
