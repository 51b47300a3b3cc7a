"""Ray sharding host logic with world_size=2 on CPU (gloo): sharded == unsharded, bit for bit, including the
batch-global delta0 / any_step_zero quantities; SUM all-reduce == full-batch gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nerf_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(step_zero=False):
    from nerf_tiny_b200 import synth
    h, w, f = 378, 504, 407.6
    rows17 = synth.pose_rows(20, h, w, f, llff_bounds=True, seed=3)          # per-image near/far (cfg4 shape)
    row, col, pix, pb, pic = synth.random_batch(rows17, 10, h, w, torch.Generator().manual_seed(21))
    if step_zero:
        pb = pb.clone()
        pb[9, 16] = pb[9, 15]                                                  # a degenerate ray in rank 1's shard (rays 8-9)
    return row, col, pix, pb, synth.k_inv_of(h, w, f)


def _worker(rank, world, port, step_zero, out):
    from nerf_tiny_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    row, col, pix, pb, k_inv = _batch(step_zero)
    sl = D.shard_slice(row.shape[0], rank, world)
    pbf = pb.to(torch.float32)
    zero, delta0 = D.global_quantities(pbf[sl, 15].numpy(), pbf[sl, 16].numpy())
    sd = {k: v.double().requires_grad_(True) for k, v in O.init_state_dict(624).items()}
    sd32 = O.init_state_dict(624)
    if step_zero:      # the degenerate ray makes resample raise (reference: exit(0)); only the global quantities matter
        cf = torch.zeros(sl.stop - sl.start, 3)
    else:
        with torch.no_grad():
            cc, cf = O.forward(sd32, row[sl].numpy(), col[sl].numpy(), pb[sl], k_inv, any_step_zero=zero,
                               delta0=torch.tensor(float(delta0)))
    full_c = D.gather_rows(cf, row.shape[0])
    # gradient: local sum-loss backward, then ONE all-reduce of the flat buffer
    if not step_zero:
        c1, c2 = O.forward(sd, row[sl].numpy(), col[sl].numpy(), pb[sl], k_inv, any_step_zero=zero,
                           delta0=torch.tensor(float(delta0), dtype=torch.float64))
        O.ray_loss(c1, c2, pix[sl].double()).backward()
        flat = torch.cat([sd[k + n].grad.reshape(-1) for k in O.LAYER_KEYS for n in (".weight", ".bias")])
        D.allreduce_sum_(flat)
    else:
        flat = torch.zeros(1)
    if rank == 0:
        torch.save({"c_fine": full_c, "grad": flat, "zero": zero, "delta0": float(delta0)}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("step_zero", [False, True])
def test_sharded_equals_unsharded(tmp_path, step_zero):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), step_zero, out), nprocs=2, join=True)
    got = torch.load(out)
    row, col, pix, pb, k_inv = _batch(step_zero)
    sd32 = O.init_state_dict(624)
    if step_zero:
        # the reference itself dies on this batch (degenerate ray -> exit(0)); the global branch must still be agreed
        assert got["zero"] is True
        t_ref = O.t_coarse_of(pb[:, 15].float().numpy(), pb[:, 16].float().numpy())
        assert got["delta0"] == float(t_ref[0, 1] - t_ref[0, 0])
        return
    with torch.no_grad():
        cc, cf, aux = O.forward(sd32, row.numpy(), col.numpy(), pb, k_inv, return_aux=True)
    assert got["zero"] is False
    assert got["delta0"] == float(aux["t_coarse"][0, 1] - aux["t_coarse"][0, 0])
    assert torch.equal(got["c_fine"], cf)                       # sharded render == unsharded, bit for bit
    sd = {k: v.double().requires_grad_(True) for k, v in sd32.items()}
    c1, c2 = O.forward(sd, row.numpy(), col.numpy(), pb, k_inv)
    O.ray_loss(c1, c2, pix.double()).backward()
    flat = torch.cat([sd[k + n].grad.reshape(-1) for k in O.LAYER_KEYS for n in (".weight", ".bias")])
    rel = float((got["grad"] - flat).norm() / flat.norm())
    assert rel < 1e-9, rel                                       # SUM all-reduce == full-batch gradient


def test_shard_slices_cover():
    from nerf_tiny_b200 import dist as D
    for n in (0, 1, 7, 8, 4096, 160000):
        for world in (1, 2, 4, 8):
            idx = np.concatenate([np.arange(n)[D.shard_slice(n, r, world)] for r in range(world)])
            assert np.array_equal(idx, np.arange(n))
