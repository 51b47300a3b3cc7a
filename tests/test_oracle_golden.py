"""The oracle (oracle/nerf_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py, run in the build container).  CPU only."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O

CASES = ["kat8", "lego48", "lego48_trained", "fern64", "fern64_trained", "trained64"]


def _bits(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


def _sd(case):
    if case == "trained64":      # weights the reference itself trained (oracle/make_trained_golden.py), stored as fp16
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", "trained_weights_fp16.npz"))
        return {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if not k.startswith("__")}
    sd = O.init_state_dict(624)
    return O.trained_like(sd) if case.endswith("trained") else sd


def _hash(sd):
    h = hashlib.sha256()
    for k in sd:
        h.update(sd[k].numpy().tobytes())
    return h.hexdigest()


def test_weights_reproducible(golden_dir):
    g = np.load(os.path.join(golden_dir, "weights_sha.npz"))
    sd = O.init_state_dict(624)
    assert _hash(sd) == str(g["init"])
    assert _hash(O.trained_like(sd)) == str(g["trained"])
    assert sum(v.numel() for v in sd.values()) == O.N_PARAMS == 593924


@pytest.mark.parametrize("case", CASES)
def test_forward_matches_reference(case, golden_dir):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    sd = _sd(case)
    with torch.no_grad():
        cc, cf, aux = O.forward(sd, g["row"], g["col"], torch.from_numpy(g["poses_bound"]),
                                torch.from_numpy(g["k_inv"]), return_aux=True)
    # bit-exact targets (north_star: ray directions, sample indices, searchsorted outputs)
    for name in ("d_cam", "d_wrd"):
        assert np.array_equal(_bits(aux[name]), _bits(g[name])), name
    for name in ("t_coarse", "t_fine", "u", "cdf", "w_c"):
        assert np.array_equal(_bits(aux[name].numpy()), _bits(g[name])), name
    assert np.array_equal(aux["idx"].numpy().astype(np.int32), g["idx"])
    # tolerance targets
    tol = 5e-6
    assert np.abs(cc.numpy() - g["c_coarse"]).max() <= tol
    assert np.abs(cf.numpy() - g["c_fine"]).max() <= tol
    assert np.abs(aux["color_f"].numpy() - g["color_f"]).max() <= tol
    assert np.abs(aux["sigma_c"].numpy() - g["sigma_c"]).max() <= tol * max(1.0, np.abs(g["sigma_c"]).max())


def test_encoder_and_network(golden_dir):
    g = np.load(os.path.join(golden_dir, "encoder_network.npz"))
    pts, dirs = torch.from_numpy(g["pts"]), torch.from_numpy(g["dirs"])
    pe, de = O.encode(pts, 10), O.encode(dirs, 4)
    assert np.array_equal(_bits(pe.numpy()), _bits(g["point_enc"]))
    assert np.array_equal(_bits(de.numpy()), _bits(g["dir_enc"]))
    with torch.no_grad():
        color, sigma = O.network_forward(O.trained_like(O.init_state_dict(624)), pe, de)
    assert np.abs(color.numpy() - g["color"]).max() < 1e-6
    assert np.abs(sigma.numpy() - g["sigma"]).max() < 1e-4
    assert np.array_equal(_bits(O.freq_table(10)), _bits(O.freq_from_hex(O.FREQ_POINT_HEX)))
    assert np.array_equal(_bits(O.freq_table(4)), _bits(O.freq_from_hex(O.FREQ_DIR_HEX)))


def test_fp64_gradients(golden_dir):
    """fp64 oracle gradients vs the fp64 reference, sampled entries + norms."""
    g = np.load(os.path.join(golden_dir, "grads_fp64_kat8_trained.npz"))
    k = np.load(os.path.join(golden_dir, "kat8.npz"))
    sd = {n: v.double().requires_grad_(True) for n, v in O.trained_like(O.init_state_dict(624)).items()}
    cc, cf = O.forward(sd, k["row"], k["col"], torch.from_numpy(k["poses_bound"]), torch.from_numpy(k["k_inv"]))
    loss = O.ray_loss(cc, cf, torch.full((8, 3), 0.5, dtype=torch.float64))
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-9
    for n in sd:
        flat = sd[n].grad.flatten()
        assert abs(float(flat.norm()) - float(g["norm/" + n])) <= 1e-8 * float(g["norm/" + n]) + 1e-12
        np.testing.assert_allclose(flat[torch.from_numpy(g["sel/" + n])].numpy(), g["val/" + n], rtol=1e-7, atol=1e-10)


def test_adam(golden_dir):
    g = np.load(os.path.join(golden_dir, "adam5.npz"))
    p = torch.from_numpy(g["p0"]).clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for s, gr in enumerate(g["grads"], 1):
        O.adam_step(p, torch.from_numpy(gr), m, v, s, 3e-4)
    assert np.abs(p.numpy() - g["p_final"]).max() < 1e-6


def test_lr_lambda():
    # nerf.py:426: collapses to gamma*learning after decay_end
    assert O.lr_lambda(0, 0.1, 1e4, 3e-4) == 1.0
    assert abs(O.lr_lambda(5000, 0.1, 1e4, 3e-4) - 0.1 ** 0.5) < 1e-12
    assert O.lr_lambda(10000, 0.1, 1e4, 3e-4) == 0.1 * 3e-4


def test_linspace_edge_cases():
    # step==0 in ANY row switches every row to the other formula (numpy's any_step_zero branch)
    a = O.linspace_rows(np.array([2.0, 3.0]), np.array([6.0, 3.0]), 64)
    ref = np.linspace((np.float32(2.0), np.float32(3.0)), (np.float32(6.0), np.float32(3.0)), 64).T
    assert np.array_equal(_bits(a), _bits(ref))
    rng = np.random.RandomState(0)
    near = rng.uniform(0.5, 3, 257).astype(np.float32)
    far = (near + rng.uniform(0.1, 12, 257)).astype(np.float32)
    ref = np.linspace(tuple(near), tuple(far), 64).T
    assert ref.dtype == np.float32
    assert np.array_equal(_bits(O.t_coarse_of(near, far)), _bits(ref))


def test_resample_range_error():
    # the reference exit(0)s here (nerf.py:251-253); the oracle raises
    t = torch.from_numpy(O.t_coarse_of(np.array([2.0]), np.array([6.0])))
    with pytest.raises(O.ResampleRangeError):
        O.resample(t, torch.zeros(1, 64))
