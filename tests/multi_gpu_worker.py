"""Worker of tests/test_gpu_multi.py (run under torchrun, one process per GPU, NCCL).  Prints one PASS / FAIL line on rank 0.
  1. ray-sharded render of a fern-shape batch (per-ray near/far) with shard=(rank, world): the batch-global quantities are
     reduced on the device with ONE 16-byte NCCL MAX all-reduce; gathered result == unsharded launch, bit for bit.
  2. ray-sharded training steps (GraphedTrainStep + fused NVLink peer all-reduce + Adam) == unsharded training, and all
     ranks' parameters stay bit-identical.
  3. fused peer all-reduce + Adam == NCCL all-reduce + Adam.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from nerf_tiny_b200 import dist as D
from nerf_tiny_b200 import nerf, synth
from oracle import nerf_oracle as O

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok, msgs = True, []


def check(cond, msg):
    global ok
    ok = ok and bool(cond)
    msgs.append(("ok  " if cond else "FAIL") + " " + msg)


# ---- 1. sharded render == unsharded (cfg4 shape: 378 x 504, per-image near/far) --------------------------------------
h, w, f = 378, 504, 407.6
rows17 = synth.pose_rows(20, h, w, f, llff_bounds=True, seed=3)
k_inv = synth.k_inv_of(h, w, f)
N = 4096
row, col, pix, pb, pic = synth.random_batch(rows17, N, h, w, torch.Generator().manual_seed(21))
sd = O.init_state_dict(624)


def model(prec):
    m = nerf.NeRFModel(64, 128, batch_ray=N, precision=prec)
    m.load_state_dict(sd)
    m = m.to(dev)
    m.check_range = False
    return m


for prec in ("fp32", "fp16"):
    m = model(prec)
    sl = D.shard_slice(N, rank, world)
    with torch.no_grad():
        cc_full, cf_full = m(row, col, pb, k_inv)                                  # every rank renders the whole batch
        cc, cf = m(row[sl], col[sl], pb[sl], k_inv, shard=(rank, world))         # ... and its shard with the globals
        cf_naive = m(row[sl], col[sl], pb[sl], k_inv)[1]
    got = D.gather_rows(cf, N)
    check(torch.equal(got, cf_full), f"sharded render == unsharded ({prec}, {world} ranks)")
    if rank > 0:
        check(not torch.equal(cf_naive, cf_full[sl]), f"shard without the globals differs ({prec})")
    m.check_status()

# ---- 2. sharded training == unsharded training --------------------------------------------------------------------------
batches = [synth.random_batch(rows17, N, h, w, torch.Generator().manual_seed(100 + s)) for s in range(3)]


def train(sharded):
    m = model("bf16")
    m.train()
    opt = nerf.FusedAdam(m, lr=3e-4)
    if sharded:
        fused = opt.enable_peer_allreduce()
        sl = D.shard_slice(N, rank, world)
        gs = nerf.GraphedTrainStep(m, opt, sl.stop - sl.start, k_inv, shard=(rank, world))
        ar = (lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM)) if not fused else None
        for b in batches:
            gs(b[0][sl], b[1][sl], b[2][sl], b[3][sl], grad_allreduce=ar)
    else:
        fused = None
        for b in batches:
            nerf.train_step(m, opt, b[0], b[1], b[2], b[3], k_inv)
    torch.cuda.synchronize()
    return m.network.flat_params().clone(), fused


p_full, _ = train(False)
p_shard, fused = train(True)
ref = p_shard.clone()
dist.broadcast(ref, src=0)
check(torch.equal(ref, p_shard), "ranks bit-identical after sharded training (fused peer exchange: %s)" % fused)
rel = float((p_shard - p_full).norm() / p_full.norm())
upd = float((p_full - torch.cat([sd[k + n].reshape(-1) for k in O.LAYER_KEYS for n in (".weight", ".bias")]).to(dev)).norm() / p_full.norm())
check(rel < 0.2 * upd, f"sharded training == unsharded: rel diff {rel:.2e} vs update size {upd:.2e}")

# ---- 3. fused peer all-reduce + Adam == NCCL all-reduce + Adam -----------------------------------------------------------
rows17b = synth.pose_rows(8, 400, 400, synth.focal_of(400))
kinv_b = synth.k_inv_of(400, 400, synth.focal_of(400))
bb = [synth.random_batch(rows17b, 1024, 400, 400, torch.Generator().manual_seed(100 * s + rank)) for s in range(2)]


def run(use_peer):
    m = nerf.NeRFModel(64, 128, batch_ray=1024, precision="bf16")
    m.load_state_dict(sd)
    m = m.to(dev)
    m.train()
    m.check_range = False
    opt = nerf.FusedAdam(m, lr=3e-4)
    got = opt.enable_peer_allreduce() if use_peer else False
    ar = lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM)
    for b in bb:
        nerf.train_step(m, opt, b[0], b[1], b[2], b[3], kinv_b, grad_allreduce=ar)
    torch.cuda.synchronize()
    return m.network.flat_params().clone(), got


p_nccl, _ = run(False)
p_peer, got = run(True)
rel = float((p_peer - p_nccl).norm() / p_nccl.norm())
check(got and rel < 2e-3, f"fused peer exchange vs NCCL after 2 steps: rel {rel:.2e} (peer enabled: {got})")

flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
for r in range(world):
    if r == rank and (rank == 0 or not ok):
        print(f"[rank {rank}]\n  " + "\n  ".join(msgs), flush=True)
    dist.barrier()
if rank == 0:
    print("PASS" if flag.item() == 1.0 else "FAIL", flush=True)
dist.destroy_process_group()
