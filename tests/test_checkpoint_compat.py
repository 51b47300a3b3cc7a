"""Checkpoint files are the reference's own format (nerf.py:491 torch.save(self.model); nerf.py:410 torch.load(...)):
a whole-module pickle whose classes are recorded as `nerf.*`.  CPU-only: pickling never touches the GPU."""
import os
import sys

import pytest
import torch

REF = "/root/reference"


def _same(a, b):
    sa, sb = a.state_dict(), b.state_dict()
    return list(sa.keys()) == list(sb.keys()) and all(torch.equal(sa[k].cpu(), sb[k].cpu()) for k in sa)


def test_roundtrip_and_recorded_class_names(tmp_path):
    from nerf_tiny_b200 import nerf
    m = nerf.NeRFModel(64, 128, batch_ray=8, precision="bf16")
    path = str(tmp_path / "t_7.pkl")
    nerf.save_checkpoint(m, path)
    m2 = nerf.load_checkpoint(path)
    assert isinstance(m2, nerf.NeRFModel) and m2.precision == "bf16" and _same(m, m2)
    assert m2.network.flat_params().numel() == 593924
    import zipfile
    blob = zipfile.ZipFile(path).read([n for n in zipfile.ZipFile(path).namelist() if n.endswith("data.pkl")][0])
    for name in (b"NeRFModel", b"Network", b"Encoder", b"Activation"):
        assert b"cnerf\n" + name + b"\n" in blob                 # GLOBAL nerf.<class>: what the reference's torch.load resolves
    assert b"nerf_tiny_b200" not in blob


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference lives in the build container only")
def test_reference_reads_ours_and_we_read_the_reference(tmp_path):
    from nerf_tiny_b200 import nerf
    from oracle import ref_harness as RH
    ref = RH.import_reference()
    sys.modules.setdefault("nerf", ref)                           # the name the reference's own process has it under
    try:
        ours = nerf.NeRFModel(64, 128, batch_ray=8)
        p1 = str(tmp_path / "a_3.pkl")
        nerf.save_checkpoint(ours, p1)
        theirs = torch.load(p1, weights_only=False)              # nerf.py:410, unmodified
        assert type(theirs) is ref.NeRFModel and type(theirs.network) is ref.Network and _same(ours, theirs)
        ref.device = torch.device("cpu")
        row, col = torch.arange(8) * 7 + 20, torch.arange(8) * 5 + 30
        pb = torch.zeros(8, 17, dtype=torch.float64)
        pb[:, [0, 6, 12]] = 1.0
        pb[:, 14], pb[:, 15], pb[:, 16] = 4.0, 2.0, 6.0
        pb[:, [4, 9]] = 100.0
        pb[:, 14] = 138.88
        pb[:, 11] = 4.0
        k_inv = torch.tensor([[1.0, 0.0, -50.0], [0.0, -1.0, 50.0], [0.0, 0.0, -138.88]]).t()
        theirs.batch_ray = theirs.encoder.batch_size = theirs.network.batch_size = 8
        with torch.no_grad():
            cc, cf = theirs(row, col, pb, k_inv)                  # the reference runs a model restored from OUR file
        assert torch.isfinite(cf).all()

        rm = ref.NeRFModel(64, 128, 8)
        p2 = str(tmp_path / "b_5.pkl")
        torch.save(rm, p2)                                        # nerf.py:491, unmodified
        back = nerf.load_checkpoint(p2)
        assert isinstance(back, nerf.NeRFModel) and isinstance(back.network, nerf.Network) and _same(rm, back)
        assert back.precision in nerf.PRECISION and back.check_range is True
    finally:
        if sys.modules.get("nerf") is ref:
            del sys.modules["nerf"]
