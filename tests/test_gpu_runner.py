"""Runner shell on the GPU: image folder -> GPU-resident batches -> train -> checkpoint ->
resume -> display; whole-module pickle of the model as the reference saves it (nerf.py:491)."""
import glob
import io
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make_scene(root, n=2, h=16, w=16):
    from PIL import Image
    rng = np.random.RandomState(0)
    for mode in ("train", "val", "test"):
        os.makedirs(root + mode, exist_ok=True)
        frames = []
        for i in range(n):
            rgba = rng.randint(0, 256, (h, w, 4), dtype=np.uint8)
            rgba[..., 3] = 255
            Image.fromarray(rgba, "RGBA").save(root + f"{mode}/r_{i}.png")
            m = np.eye(4)
            m[:3, 3] = [0.3 * i, 0.1, 4.0]
            frames.append({"file_path": f"./{mode}/r_{i}", "transform_matrix": m.tolist()})
        json.dump({"camera_angle_x": 0.69, "frames": frames}, open(root + f"transforms_{mode}.json", "w"))


def test_runner_train_checkpoint_resume_display(tmp_path):
    from nerf_tiny_b200 import nerf
    root, ck, res = str(tmp_path) + "/scene/", str(tmp_path) + "/ck/", str(tmp_path) + "/res/"
    os.makedirs(root)
    _make_scene(root)
    nerf.seed_everything(7)
    kw = dict(gpu=0, img_dir=root, results_path=res, ckpt_path=ck, low_res=1, batch_ray=64, learning=1e-3, n_coarse=64,
              n_fine=128, data_type="sync", step=4, decay_end=1000, sched="EXP")
    r1 = nerf.NeRFRunner(total_iter=8, continue_=False, **kw)
    assert r1.height == 16 and r1.num_pic == 2 and len(r1.train_dataloader) == 8
    r1.trainer("train")
    rows = [l.split(",") for l in open(glob.glob(res + "*_loss.csv")[0]).read().split()]
    assert [int(r[0]) for r in rows] == list(range(8)) and all(np.isfinite(float(r[1])) for r in rows)
    files = sorted(glob.glob(ck + "*.pkl"))
    assert [int(f.split("_")[-1][:-4]) for f in files] == [3, 7]
    w_end = r1.model.network.flat_params().clone()
    m_end, step_end = r1.optimizer.m.clone(), r1.optimizer.step_count

    r2 = nerf.NeRFRunner(total_iter=10, continue_=True, **kw)       # resumes at iteration 8 with weights AND moments
    assert r2.last_iter == 7 and r2.optimizer.step_count == step_end
    assert torch.equal(r2.model.network.flat_params(), w_end)
    assert torch.equal(r2.optimizer.m.to(w_end.device), m_end)
    assert abs(r2.optimizer.param_groups[0]["lr"] - 1e-3 * 0.1 ** (8 / 1000)) < 1e-12
    r2.trainer("train")
    assert r2.last_iter == 9 and np.isfinite(r2.last_loss)
    assert len(open(glob.glob(res + r2.start_time + "_loss.csv")[0]).read().split()) in (2, 10)   # same second -> same file
    assert not torch.equal(r2.model.network.flat_params(), w_end)

    img = r2.display(save=True)
    assert img.shape == (2, 16, 16, 3) and bool(torch.isfinite(img).all())
    assert len(glob.glob(res + "*/*.jpg")) == 2 and len(glob.glob(res + "*/video.gif")) == 1
    assert float((img - 1.0).abs().max()) > 0 and bool((img != 1.0).reshape(2, -1).any(dim=1).all())   # every view rendered


def test_model_whole_module_pickle():
    from nerf_tiny_b200 import nerf, synth
    dev = torch.device("cuda:0")
    nerf.seed_everything(3)
    model = nerf.NeRFModel(batch_ray=32).to(dev)
    rows17 = synth.pose_rows(3, 100, 100, synth.focal_of(100))
    row, col, _, pb, _ = synth.random_batch(rows17, 32, 100, 100, torch.Generator().manual_seed(5))
    kinv = synth.k_inv_of(100, 100, synth.focal_of(100))
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    clone = torch.load(buf, map_location=dev, weights_only=False)
    assert torch.equal(clone.network.flat_params(), model.network.flat_params())
    assert list(clone.state_dict().keys()) == list(model.state_dict().keys())
    with torch.no_grad():
        a = model(row, col, pb, kinv)[1]
        b = clone(row, col, pb, kinv)[1]
    assert torch.equal(a, b)


def test_graphed_train_step_equals_launch_by_launch():
    """GraphedTrainStep (CUDA-graph replay of forward / loss / backward, Adam outside) against train_step on changing
    batches.  The learning rate is 0 for the comparison so both paths see the same weights at every step: training itself
    is chaotic and the loss / split-K weight gradients are summed with fp32 atomics, so trajectories of
    two runs of EITHER path drift apart; per-step losses and gradients agree to rounding."""
    from nerf_tiny_b200 import nerf, synth
    dev = torch.device("cuda:0")
    rows17 = synth.pose_rows(6, 100, 100, synth.focal_of(100))
    kinv = synth.k_inv_of(100, 100, synth.focal_of(100))
    gen = torch.Generator().manual_seed(17)
    batches = [synth.random_batch(rows17, 256, 100, 100, gen) for _ in range(4)]
    results = []
    for graphed in (False, True):
        nerf.seed_everything(5)
        model = nerf.NeRFModel(batch_ray=256).to(dev)
        model.check_range = False
        model.train()
        opt = nerf.FusedAdam(model, lr=0.0)
        step = nerf.GraphedTrainStep(model, opt, 256, kinv) if graphed else None
        out = []
        for b in batches:
            if graphed:
                loss, _, cf = step(b[0], b[1], b[2], b[3])
            else:
                loss, _, cf = nerf.train_step(model, opt, b[0], b[1], b[2], b[3], kinv)
            out.append((float(loss), cf.clone(), model.network.flat_grads().clone()))
        model.check_status()
        results.append(out)
    for (la, cfa, ga), (lb, cfb, gb) in zip(*results):
        assert abs(la - lb) <= 1e-5 * abs(la)
        assert torch.equal(cfa, cfb)                                   # the forward has no atomics: bit-identical
        assert float((ga - gb).abs().max()) <= 1e-4 * float(ga.abs().max())
    assert len({round(r[0], 3) for r in results[1]}) == len(batches)   # every replay consumed its own batch
    # with a real learning rate the optimizer step after the replay moves the weights
    opt = nerf.FusedAdam(model, lr=1e-3)
    step = nerf.GraphedTrainStep(model, opt, 256, kinv)
    before = model.network.flat_params().clone()
    step(*batches[0][:4])
    assert not torch.equal(before, model.network.flat_params())


def test_get_density_and_color_cum_methods():
    """NeRFModel.get_density / color_cum (nerf.py:263-281) on caller-supplied tensors: against the reference's own
    coarse-pass outputs (goldens written by the unmodified reference) and against the oracle for 192 merged samples."""
    import numpy as np
    from nerf_tiny_b200 import nerf
    from oracle import nerf_oracle as O
    dev = torch.device("cuda:0")
    model = nerf.NeRFModel(batch_ray=8).to(dev)
    for name in ("lego48_trained", "fern64_trained"):
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
        near, far = torch.from_numpy(g["near"]), torch.from_numpy(g["far"])
        sigma, color = torch.from_numpy(g["sigma_c"]), torch.from_numpy(g["color_c"])
        delta = ((far - near) / 64).unsqueeze(1).repeat(1, 64)                   # nerf.py:293
        w = model.get_density(delta, sigma)
        assert float((w.cpu() - torch.from_numpy(g["w_c"])).abs().max()) <= 2e-7   # fp64 prefix, 1-ulp expf differences
        c = model.color_cum(w, color)
        assert float((c.cpu() - torch.from_numpy(g["c_coarse"])).abs().max()) <= 1e-6
    gen = torch.Generator().manual_seed(2)
    delta = torch.rand(37, 192, generator=gen) * 0.05
    sigma = torch.rand(37, 192, generator=gen) * 30
    color = torch.rand(37, 192, 3, generator=gen)
    w = model.get_density(delta, sigma.unsqueeze(-1))                             # sigma [N,P,1] as Network returns it
    w_ref = O.get_density(delta, sigma)
    assert float((w.cpu() - w_ref).abs().max()) <= 2e-7
    assert float((model.color_cum(w, color).cpu() - O.color_cum(w_ref, color)).abs().max()) <= 1e-6
    with pytest.raises(nerf._lib.NerfTinyError):
        model.get_density(delta[:, :100], sigma[:, :100])                         # not a multiple of 32


def test_piecewise_methods_are_differentiable():
    """The reference's piecewise methods are ordinary autograd code (nerf.py:179-281); here get_density / color_cum /
    resample / net_out record autograd nodes whose gradients match fp64 autograd of the oracle on the same inputs."""
    import numpy as np
    from nerf_tiny_b200 import nerf, synth
    from oracle import nerf_oracle as O
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    model = nerf.NeRFModel(batch_ray=8, precision="fp32").to(dev)
    # ---- get_density + color_cum chained, gradients w.r.t. delta, sigma, colour
    delta = (torch.rand(19, 64, generator=gen) * 0.06 + 1e-3)
    sigma = torch.rand(19, 64, generator=gen) * 8
    color = torch.rand(19, 64, 3, generator=gen)
    seed = torch.rand(19, 3, generator=gen)
    d, s_, c = (x.clone().to(dev).requires_grad_(True) for x in (delta, sigma, color))
    out = model.color_cum(model.get_density(d, s_), c)
    (out * seed.to(dev)).sum().backward()
    d64, s64, c64 = (x.double().requires_grad_(True) for x in (delta, sigma, color))
    (O.color_cum(O.get_density(d64, s64), c64) * seed.double()).sum().backward()
    for got, ref in ((d.grad, d64.grad), (s_.grad, s64.grad), (c.grad, c64.grad)):
        assert float((got.cpu().double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    # ---- resample, gradient w.r.t. the coarse weights (no gradient to t_coarse, like the reference)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fern64.npz"))
    t_c, w_c = torch.from_numpy(g["t_coarse"]), torch.from_numpy(g["w_c"])
    seed = torch.rand(t_c.shape[0], 128, generator=gen)
    w_dev = w_c.clone().to(dev).requires_grad_(True)
    (model.resample(t_c, w_dev) * seed.to(dev)).sum().backward()
    w64 = w_c.double().requires_grad_(True)
    t_ref, idx, u, cdf = O.resample(t_c.double(), w64, 128, return_aux=True)
    (t_ref * seed.double()).sum().backward()
    # the fp32 kernel and the fp64 oracle agree on idx for this fixture (checked bit-exactly elsewhere); the factor
    # delta0 / (w + 1e-7)^2 makes single entries huge, hence the relative measure
    assert float((w_dev.grad.cpu().double() - w64.grad).abs().max()) <= 2e-3 * float(w64.grad.abs().max())
    # ---- net_out: parameter gradients and the gradient w.r.t. t (fp32 model -> fp32 training kernels)
    sd = O.init_state_dict(624)
    model.load_state_dict(sd)
    model = model.to(dev)
    h = w = 100
    f = synth.focal_of(w)
    rows17 = synth.pose_rows(8, h, w, f)
    k_inv = synth.k_inv_of(h, w, f)
    row, col, pix, pb, pic = synth.random_batch(rows17, 12, h, w, gen)
    c2w = O.poses_extract(pb)[0]
    t = (torch.rand(12, 32, generator=gen) * 4 + 2)
    t_dev = t.clone().to(dev).requires_grad_(True)
    sr, ss = torch.rand(12, 32, 3, generator=gen), torch.rand(12, 32, 1, generator=gen)
    col_g, sig_g = model.net_out(t_dev, row, col, pb, k_inv, 32)
    assert col_g.shape == (12, 32, 3) and sig_g.shape == (12, 32, 1)
    ((col_g * sr.to(dev)).sum() + (sig_g * ss.to(dev)).sum()).backward()
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    d_cam, d_wrd = O.ray_dirs(row.numpy(), col.numpy(), k_inv.numpy(), c2w.numpy())
    t64 = t.double().requires_grad_(True)
    c_ref, s_ref = O.net_out(sd64, t64, d_cam.astype(np.float64), d_wrd.astype(np.float64), c2w.numpy().astype(np.float64),
                             faithful32=True)
    ((c_ref * sr.double()).sum() + (s_ref.unsqueeze(-1) * ss.double()).sum()).backward()
    assert float((col_g.detach().cpu().double() - c_ref.detach()).abs().max()) <= 1e-4
    rel = lambda got, ref: float((got.cpu().double() - ref).norm() / ref.norm().clamp_min(1e-20))
    # d enc / dt carries frequencies up to 3217 with cancellation: same bound as test_mlp_fp32_backward
    assert rel(t_dev.grad, t64.grad) <= 2e-2
    for name, prm in model.named_parameters():
        assert rel(prm.grad, sd64[name].grad) <= 2e-3, name
    # without grad the same calls take the rendering kernels and record nothing
    with torch.no_grad():
        assert not model.net_out(t, row, col, pb, k_inv, 32)[0].requires_grad


def test_encoder_and_network_standalone_modules():
    """Encoder.forward / Network.forward called the way the reference exposes them (nerf.py:101-167), against the outputs
    the unmodified reference wrote for the same inputs and weights (tests/golden/encoder_network.npz)."""
    import numpy as np
    from nerf_tiny_b200 import nerf
    from oracle import nerf_oracle as O
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "encoder_network.npz"))
    pts, dirs = torch.from_numpy(g["pts"]), torch.from_numpy(g["dirs"])
    model = nerf.NeRFModel(batch_ray=pts.shape[0])
    model.load_state_dict(O.trained_like(O.init_state_dict(624)))
    model = model.to(dev)
    nerf.device = dev
    state = torch.get_rng_state()
    gp, gd = model.encoder(pts.shape[1], pts, dirs)
    after = torch.get_rng_state()
    torch.set_rng_state(state)
    torch.rand(1, 1, 14, 1, 1)
    assert torch.equal(after, torch.get_rng_state())                    # same 14 draws as nerf.py:141
    assert gp.shape == (4, 16, 3, 20) and gd.shape == (4, 16, 3, 8)
    # CUDA sincosf vs ATen's CPU sin / cos: arguments reach |w x| ~ 2e4, both are correctly reduced, results differ by ulps
    assert float((gp.flatten(start_dim=2).cpu() - torch.from_numpy(g["point_enc"])).abs().max()) <= 5e-7
    assert float((gd.flatten(start_dim=2).cpu() - torch.from_numpy(g["dir_enc"])).abs().max()) <= 5e-7
    ref_gp = torch.from_numpy(g["point_enc"]).reshape(4, 16, 3, 20)
    ref_gd = torch.from_numpy(g["dir_enc"]).reshape(4, 16, 3, 8)
    color, sigma = model.network(pts.shape[1], ref_gp, ref_gd)
    assert color.shape == (4, 16, 3) and sigma.shape == (4, 16, 1)
    assert float((color.cpu() - torch.from_numpy(g["color"])).abs().max()) <= 5e-6
    assert float((sigma.cpu() - torch.from_numpy(g["sigma"])).abs().max()) <= 5e-6 * max(1.0, float(np.abs(g["sigma"]).max()))
