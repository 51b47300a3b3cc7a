"""Diagnostic: clock64 timeline of the fused backward-data chain kernel (bwd_tc.cu), block 0, pairs 1-2 of the step's last
chain launch (the coarse pass).  NT_DW_OVERLAP_CTAS=0 keeps the weight-gradient launch off the SMs while it runs."""
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
prof = torch.zeros(4 * 9 * 16, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.nt_debug_set_chain_prof.argtypes = [ctypes.c_void_p]
lib.nt_debug_set_chain_prof.restype = None
lib.nt_debug_set_chain_prof(ctypes.c_void_p(prof.data_ptr()))
nerf.train_step(m, opt, b[0], b[1], b[2], b[3], kinv)
torch.cuda.synchronize()
lib.nt_debug_set_chain_prof(ctypes.c_void_p(0))
p = prof.cpu().numpy().reshape(4, 9, 16)
names = ["actA", "actB", "w0", "w1", "w2", "w3", "issA", "issB", "ld1", "ld3", "epiA_wake", "epiB_wake", "epiA_done", "epiB_done",
         "csA", "csB"]
t0 = p[1, 0, 0]
for pl in (1, 2):
    print("pair", pl)
    for st in range(9):
        print(f" s{st}: " + "  ".join(f"{names[i]}={p[pl, st, i] - t0}" for i in range(16) if p[pl, st, i]))
