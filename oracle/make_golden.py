"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle to it.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden
For every case the script (1) runs /root/reference/nerf.py, (2) runs oracle/nerf_oracle.py on
the same inputs and weights, (3) asserts agreement (bit-exact where SURVEY.md Appendix A says
so), (4) stores the REFERENCE's outputs as the fixture.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import hashlib
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


def sd_hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(sd[k].detach().numpy().tobytes())
    return h.hexdigest()


def bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


def assert_bits(name, a, b):
    a, b = bits(a), bits(b)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    bad = int((a != b).sum())
    assert bad == 0, f"{name}: {bad}/{a.size} elements differ bitwise"


def ref_stages(ref, model, row, col, pb, k_inv):
    """Drive the reference's own methods piecewise (nerf.py:286-323) and keep every intermediate."""
    with torch.no_grad():
        pbf = pb.to(torch.float)
        c2w, _, _, _, near, far = ref.poses_extract(pbf)
        nc, nf = model.num_coarse, model.num_fine
        t_c = torch.tensor(np.linspace(tuple(near), tuple(far), nc)).transpose(0, 1)
        color_c, sigma_c = model.net_out(t_c, row, col, c2w, k_inv, nc)
        delta_c = ((far - near) / nc).unsqueeze(1).repeat(1, nc)
        w_c = model.get_density(delta_c, sigma_c.squeeze())
        t_f = model.resample(t_c, w_c)
        color_f, sigma_f = model.net_out(t_f, row, col, c2w, k_inv, nf)
        cc, cf = model(row, col, pb, k_inv)
    return dict(c2w=c2w.numpy(), near=near.numpy(), far=far.numpy(), t_coarse=t_c.numpy(),
                color_c=color_c.numpy(), sigma_c=sigma_c.squeeze(-1).numpy(), w_c=w_c.numpy(),
                t_fine=t_f.numpy(), color_f=color_f.numpy(), sigma_f=sigma_f.squeeze(-1).numpy(),
                c_coarse=cc.numpy(), c_fine=cf.numpy())


def ref_dirs(ref, model, row, col, c2w, k_inv):
    """dir_cam / dir_wrd exactly as nerf.py:186-197, 211 builds them (P=1 slice)."""
    n = row.shape[0]
    xy = torch.cat((row.unsqueeze(0), col.unsqueeze(0), torch.ones(1, n)), dim=0)
    ps = torch.matmul(xy.unsqueeze(2).permute(1, 2, 0), k_inv)
    d_cam = torch.nn.functional.normalize(ps, p=2.0, dim=2).repeat(1, 2, 1)
    bm = torch.as_tensor(c2w).unsqueeze(1).repeat(1, 2, 1, 1)
    d_wrd = torch.matmul(bm[:, :, :3, :3], d_cam.unsqueeze(3)).squeeze()
    return d_cam[:, 0].numpy(), d_wrd[:, 0].numpy()


def forward_case(ref, name, sd, row, col, pb, k_inv, rgb_tol=2e-6):
    n = row.shape[0]
    model = RH.make_model(ref, n, sd)
    r = ref_stages(ref, model, row, col, pb, k_inv)
    r["d_cam"], r["d_wrd"] = ref_dirs(ref, model, row, col, r["c2w"], k_inv)
    with torch.no_grad():
        cc, cf, aux = O.forward(sd, row.numpy(), col.numpy(), pb, k_inv, return_aux=True)
    assert_bits(name + ".d_cam", aux["d_cam"], r["d_cam"])
    assert_bits(name + ".d_wrd", aux["d_wrd"], r["d_wrd"])
    assert_bits(name + ".t_coarse", aux["t_coarse"].numpy(), r["t_coarse"])
    assert_bits(name + ".w_c", aux["w_c"].numpy(), r["w_c"])
    assert_bits(name + ".t_fine", aux["t_fine"].numpy(), r["t_fine"])
    for k, a, b in (("color_c", aux["color_c"], r["color_c"]), ("sigma_c", aux["sigma_c"], r["sigma_c"]),
                    ("color_f", aux["color_f"], r["color_f"]), ("c_coarse", cc, r["c_coarse"]),
                    ("c_fine", cf, r["c_fine"])):
        err = float(np.abs(a.numpy() - b).max())
        assert err <= rgb_tol * max(1.0, float(np.abs(b).max())), (name, k, err)
    r["idx"] = aux["idx"].numpy().astype(np.int32)
    r["u"] = aux["u"].numpy()
    r["cdf"] = aux["cdf"].numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), row=row.numpy(), col=col.numpy(), poses_bound=pb.numpy(),
                        k_inv=k_inv.numpy(), **r)
    print(f"[golden] {name}: N={n} ok  C_fine[0]={r['c_fine'][0]}")
    return r


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    ref = RH.import_reference()

    # ---- known-answer test recorded in SURVEY.md §8(c): first model after import (seed 624)
    m0 = ref.NeRFModel(num_coarse=64, num_fine=128, batch_ray=8)
    sd0 = {k: v.detach().clone() for k, v in m0.state_dict().items()}
    assert sd_hash(sd0).startswith("dc46354e1af7cd4b"), sd_hash(sd0)
    h = w = 100
    focal = 138.88
    k_inv = O.make_k_inv(h, w, focal)
    assert torch.equal(k_inv, synth.k_inv_of(h, w, focal))
    c2w = np.concatenate((np.eye(3), np.array([[0.0], [0.0], [4.0]])), axis=1)
    pb = torch.from_numpy(np.tile(np.concatenate((np.concatenate((c2w, [[h], [w], [focal]]), axis=1).flatten(),
                                                 [2.0, 6.0])), (8, 1)))
    row = torch.arange(8) * 7 + 20
    col = torch.arange(8) * 5 + 30
    cc, cf = m0(row, col, pb, k_inv)
    np.testing.assert_allclose(cc[0].detach().numpy(), [0.108975239, 0.098129459, 0.106431261], atol=2e-7)
    np.testing.assert_allclose(cf[0].detach().numpy(), [0.109713502, 0.098793820, 0.107147627], atol=2e-7)
    loss = m0.ray_loss(cc, cf, torch.full((8, 3), 0.5))
    assert abs(float(loss) - 7.535269737) < 2e-5, float(loss)
    print("[golden] survey KAT reproduced; loss", float(loss))

    # ---- deterministic weights shared by every fixture (seeded generator, not the global RNG)
    sd = O.init_state_dict(624)
    sd_t = O.trained_like(sd)
    np.savez(os.path.join(GOLD, "weights_sha.npz"), init=np.array(sd_hash(sd)), trained=np.array(sd_hash(sd_t)))

    # ---- KAT inputs with the shared weights
    forward_case(ref, "kat8", sd, row, col, pb, k_inv)

    # ---- cfg1-shaped: 100x100 lego view, 48 pixels spread over the image
    rows17 = synth.pose_rows(8, 100, 100, synth.focal_of(100))
    g = torch.Generator().manual_seed(11)
    r_, c_, pix, pbb, pic = synth.random_batch(rows17, 48, 100, 100, g)
    forward_case(ref, "lego48", sd, r_, c_, pbb, synth.k_inv_of(100, 100, synth.focal_of(100)))
    forward_case(ref, "lego48_trained", sd_t, r_, c_, pbb, synth.k_inv_of(100, 100, synth.focal_of(100)), rgb_tol=5e-6)

    # ---- cfg4-shaped: 504x378 LLFF-like, per-image near/far (no NDC), trained-like weights
    hh, ww, ff = 378, 504, 407.6
    rows17f = synth.pose_rows(20, hh, ww, ff, llff_bounds=True, seed=3)
    g = torch.Generator().manual_seed(12)
    r_, c_, pix, pbb, pic = synth.random_batch(rows17f, 64, hh, ww, g)
    forward_case(ref, "fern64", sd, r_, c_, pbb, synth.k_inv_of(hh, ww, ff))
    forward_case(ref, "fern64_trained", sd_t, r_, c_, pbb, synth.k_inv_of(hh, ww, ff), rgb_tol=5e-6)

    # ---- encoder / network piecewise (nerf.py:135-167, 101-124)
    g = torch.Generator().manual_seed(5)
    pts = (torch.rand(4, 16, 3, generator=g) * 12 - 6)
    dirs = torch.nn.functional.normalize(torch.randn(4, 16, 3, generator=g), dim=-1)
    enc = ref.Encoder(batch_size=4)
    with torch.no_grad():
        gp, gd = enc.forward(16, pts, dirs)
        net = RH.make_model(ref, 4, sd_t).network
        color, sigma = net.forward(16, gp, gd)
        oc, osig = O.network_forward(sd_t, O.encode(pts, 10), O.encode(dirs, 4))
    assert_bits("enc.point", O.encode(pts, 10).numpy(), gp.flatten(start_dim=2).numpy())
    assert_bits("enc.dir", O.encode(dirs, 4).numpy(), gd.flatten(start_dim=2).numpy())
    assert_bits("freq.point", O.freq_table(10), O.freq_from_hex(O.FREQ_POINT_HEX))
    assert_bits("freq.dir", O.freq_table(4), O.freq_from_hex(O.FREQ_DIR_HEX))
    assert float((oc - color).abs().max()) < 1e-6 and float((osig - sigma).abs().max()) < 1e-4
    np.savez_compressed(os.path.join(GOLD, "encoder_network.npz"), pts=pts.numpy(), dirs=dirs.numpy(),
                        point_enc=gp.flatten(start_dim=2).numpy(), dir_enc=gd.flatten(start_dim=2).numpy(),
                        color=color.numpy(), sigma=sigma.numpy())
    print("[golden] encoder/network ok")

    # ---- gradients: fp64 reference vs fp64 oracle, N=8
    torch.set_default_dtype(torch.float64)
    try:
        m64 = ref.NeRFModel(num_coarse=64, num_fine=128, batch_ray=8)
        sd64 = {k: v.double() for k, v in sd_t.items()}
        m64.load_state_dict(sd64)
        cc, cf = m64(row, col, pb, k_inv.double())
        tgt = torch.full((8, 3), 0.5)
        loss64 = m64.ray_loss(cc, cf, tgt)
        loss64.backward()
        gref = {k: p.grad.detach().clone() for k, p in m64.named_parameters()}
        sdo = {k: v.clone().requires_grad_(True) for k, v in sd64.items()}
        occ, ocf = O.forward(sdo, row.numpy(), col.numpy(), pb, k_inv)
        lo = O.ray_loss(occ, ocf, tgt)
        lo.backward()
        worst = 0.0
        for k in gref:
            rel = float((sdo[k].grad - gref[k]).norm() / (gref[k].norm() + 1e-30))
            worst = max(worst, rel)
        assert abs(float(lo) - float(loss64)) < 1e-6 * abs(float(loss64)), (float(lo), float(loss64))
        print(f"[golden] fp64 grads: oracle vs reference worst rel err {worst:.3e}; loss {float(loss64):.9f}")
        assert worst < 1e-3, worst
        pick = torch.Generator().manual_seed(9)
        store = {"loss": np.array(float(loss64))}
        for k in gref:
            flat = gref[k].flatten()
            sel = torch.randperm(flat.numel(), generator=pick)[:64]
            store["norm/" + k] = np.array(float(flat.norm()))
            store["sel/" + k] = sel.numpy()
            store["val/" + k] = flat[sel].numpy()
        np.savez_compressed(os.path.join(GOLD, "grads_fp64_kat8_trained.npz"), **store)
    finally:
        torch.set_default_dtype(torch.float32)

    # ---- Adam (nerf.py:425): oracle.adam_step vs torch.optim.Adam, 5 steps
    g = torch.Generator().manual_seed(2)
    p = torch.randn(1000, generator=g)
    p_ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([p_ref], lr=3e-4, betas=(0.9, 0.999), eps=1e-7)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    grads = []
    for s in range(1, 6):
        gr = torch.randn(1000, generator=g) * (10.0 ** (s - 3))
        grads.append(gr.numpy())
        p_ref.grad = gr.clone()
        opt.step()
        O.adam_step(p, gr, m, v, s, 3e-4)
        assert float((p - p_ref.detach()).abs().max()) < 1e-6
    np.savez_compressed(os.path.join(GOLD, "adam5.npz"), grads=np.stack(grads), p_final=p_ref.detach().numpy(),
                        p0=torch.randn(1000, generator=torch.Generator().manual_seed(2)).numpy())
    print("[golden] adam ok")

    # ---- short training run on the analytic scene (cfg3-shaped, small): reference loop body nerf.py:467-475
    train_case(ref, sd)


def train_case(ref, sd, n_rays=64, steps=24, lr=1e-3):
    h = w = 100
    f = synth.focal_of(w)
    rows17 = synth.pose_rows(8, h, w, f)
    k_inv = synth.k_inv_of(h, w, f)
    gen = torch.Generator().manual_seed(77)
    batches = [synth.random_batch(rows17, n_rays, h, w, gen) for _ in range(steps)]
    model = RH.make_model(ref, n_rays, sd)
    opt = torch.optim.Adam([{"params": model.network.parameters(), "initial_lr": lr}], lr=lr, betas=(0.9, 0.999), eps=1e-7)
    losses = []
    for (row, col, pix, pb, pic) in batches:
        opt.zero_grad()
        model.train()
        cc, cf = model(row, col, pb, k_inv)
        loss = model.ray_loss(cc, cf, pix)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    # oracle on the same batches
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ms = {k: torch.zeros_like(v) for k, v in sd.items()}
    vs = {k: torch.zeros_like(v) for k, v in sd.items()}
    olosses = []
    for s, (row, col, pix, pb, pic) in enumerate(batches, 1):
        for p in sdo.values():
            p.grad = None
        cc, cf = O.forward(sdo, row.numpy(), col.numpy(), pb, k_inv)
        loss = O.ray_loss(cc, cf, pix)
        loss.backward()
        with torch.no_grad():
            for k in sdo:
                O.adam_step(sdo[k], sdo[k].grad, ms[k], vs[k], s, lr)
        olosses.append(float(loss))
    print("[golden] train ref   :", np.round(losses, 3))
    print("[golden] train oracle:", np.round(olosses, 3))
    assert abs(olosses[0] - losses[0]) < 1e-3 * losses[0]
    np.savez_compressed(os.path.join(GOLD, "train_lego64.npz"), ref_losses=np.array(losses), oracle_losses=np.array(olosses),
                        n_rays=n_rays, steps=steps, lr=lr, seed=77)


if __name__ == "__main__":
    main()
