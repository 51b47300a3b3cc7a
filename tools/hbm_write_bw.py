"""Pure-write / pure-read / copy HBM bandwidth with stock torch kernels (context for the stash-writing training kernels)."""
import torch
dev = torch.device("cuda", 0)
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device=dev)   # 4 GiB
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
w = t(lambda: a.fill_(1.0)); print("fill  4 GiB: %.3f ms  %.2f TB/s written" % (w, 4 * 1.0737 / w))
r = t(lambda: a.sum()); print("sum   4 GiB: %.3f ms  %.2f TB/s read" % (r, 4 * 1.0737 / r))
c = t(lambda: b.copy_(a)); print("copy  4 GiB: %.3f ms  %.2f TB/s read+written" % (c, 8 * 1.0737 / c))
