"""Diagnostic: clock64 timeline of the training (stash) instantiation of the lock-step forward kernel, block 0, pairs 1-2 of
the step's last forward launch (the fine pass).  Slots: MMA thread 0/1 = operand of tile A/B ready, 2 = first chunk issued,
3 = last chunk issued (after the stash-store wait); epilogue warp groups g = 0..3 (tile A half 0, tile B half 0, A half 1,
B half 1): 4+3g = accumulator ready, 5+3g = epilogue math done, 6+3g = arrived."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NT_DW_OVERLAP_CTAS", "0")
import numpy as np, torch
from nerf_tiny_b200 import nerf, synth, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
nerf.seed_everything(624)
m = nerf.NeRFModel(64, 128, batch_ray=n, precision="bf16").to(dev)
m.check_range = False
opt = nerf.FusedAdam(m, lr=3e-4)
rows17 = synth.pose_rows(8, 400, 400, synth.focal_of(400))
kinv = synth.k_inv_of(400, 400, synth.focal_of(400))
b = synth.random_batch(rows17, n, 400, 400, torch.Generator().manual_seed(1))
m.train()
for _ in range(3):
    nerf.train_step(m, opt, b[0], b[1], b[2], b[3], kinv)
torch.cuda.synchronize()
prof = torch.zeros(4 * 10 * 16, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.nt_debug_set_fwd_prof.argtypes = [ctypes.c_void_p]
lib.nt_debug_set_fwd_prof.restype = None
lib.nt_debug_set_fwd_prof(ctypes.c_void_p(prof.data_ptr()))
nerf.train_step(m, opt, b[0], b[1], b[2], b[3], kinv)
torch.cuda.synchronize()
lib.nt_debug_set_fwd_prof(ctypes.c_void_p(0))
p = prof.cpu().numpy().reshape(4, 10, 16)
names = ["actA", "actB", "iss_first", "iss_last", "A0_wake", "A0_done", "A0_arr", "B0_wake", "B0_done", "B0_arr",
         "A1_wake", "A1_done", "A1_arr", "B1_wake", "B1_done", "B1_arr"]
t0 = p[1, 0, 0]
for pl in (1, 2):
    print("pair", pl)
    for L in range(10):
        print(f" L{L}: " + "  ".join(f"{names[i]}={p[pl, L, i] - t0}" for i in (0, 1, 2, 3, 4, 5, 6, 7, 8)))
