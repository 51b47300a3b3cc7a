"""Ray sharding across the GPUs of one box: one process per GPU, torch.distributed for plumbing.

Rays are independent, so each rank renders a contiguous slice with replicated weights and no data-path
collective.  Two quantities of the reference are *batch-global* and must be agreed on so that
sharded == unsharded bit for bit:
  * any_step_zero — numpy's linspace switches formula for EVERY ray if any ray has step == 0 (nerf.py:288)
  * delta0        — resample uses t_coarse[0,1] - t_coarse[0,0] of the batch's FIRST ray for all rays (nerf.py:234)
Training adds one SUM all-reduce of the flat gradient (the loss is a sum over rays, nerf.py:328-331).
Host logic only (numpy + torch.distributed); works with the gloo backend on CPU, which is how it is tested.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

f32 = np.float32


SHARD_ALIGN = 4      # rays: 4 x 64 coarse samples = 2 x 128 fine samples = one tile PAIR of the fused MLP kernel


def shard_slice(n: int, rank: int, world: int) -> slice:
    """Contiguous ceil(n/world)-sized slices, rounded up to a multiple of 4 rays; the last ranks may get fewer (or zero)
    rays.  The alignment keeps every sample in the same half of its 256-sample tile pair as in the unsharded launch: the
    cta_group::2 MLP kernel accumulates the skip / view K-chunk first for one tile of a pair and last for the other, so a
    shard boundary in the middle of a pair would change fp32 summation order (last-bit differences) - with aligned
    boundaries sharded == unsharded bit for bit."""
    per = (n + world - 1) // world
    per = (per + SHARD_ALIGN - 1) // SHARD_ALIGN * SHARD_ALIGN
    lo = min(n, rank * per)
    return slice(lo, min(n, lo + per))


def step_is_zero(near: np.ndarray, far: np.ndarray, n_coarse: int = 64) -> bool:
    near = np.asarray(near, dtype=f32)
    far = np.asarray(far, dtype=f32)
    step = ((far - near).astype(f32) / f32(n_coarse - 1)).astype(f32)
    return bool(np.any(step == 0))


def delta0_of(near0: float, far0: float, any_step_zero: bool, n_coarse: int = 64) -> np.float32:
    """t_coarse[0,1] - t_coarse[0,0] exactly as np.linspace produces them in fp32."""
    a, b = f32(near0), f32(far0)
    div = f32(n_coarse - 1)
    delta = f32(b - a)
    if any_step_zero:
        t0 = f32(f32(f32(f32(0) / div) * delta) + a)
        t1 = f32(f32(f32(f32(1) / div) * delta) + a)
    else:
        step = f32(delta / div)
        t0 = f32(f32(f32(0) * step) + a)
        t1 = f32(f32(f32(1) * step) + a)
    return f32(t1 - t0)


def global_quantities(near_local: np.ndarray, far_local: np.ndarray, group=None, n_coarse: int = 64):
    """(any_step_zero, delta0) of the GLOBAL batch, from each rank's local near/far (rank 0 owns global ray 0)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    flag = torch.tensor([1 if step_is_zero(near_local, far_local, n_coarse) else 0], dtype=torch.int32)
    first = torch.zeros(2, dtype=torch.float32)
    if rank == 0 and len(near_local) > 0:
        first[0], first[1] = float(near_local[0]), float(far_local[0])
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        dist.broadcast(first, src=0, group=group)
    zero = bool(flag.item())
    return zero, delta0_of(first[0].item(), first[1].item(), zero, n_coarse)


class ShardGlobals:
    """The same two quantities for the CUDA path, reduced ON THE DEVICE (no host round trip): each rank's
    `nt_shard_globals_local` writes {flag, delta0 candidate if flag == 0, candidate if flag == 1, 0} (candidates = -inf
    unless it owns global ray 0), ONE element-wise MAX all-reduce of those 16 bytes makes them global, and
    `nt_shard_globals_resolve` leaves g[0] = delta0, g[1] = any_step_zero where the render kernels read them.

    `reduce_max(t)`: in-place MAX reduction of the 4-float device tensor over the shards; defaults to
    torch.distributed.all_reduce(MAX) on `group` (NCCL on GPUs).  Tests that emulate several shards on one GPU pass their
    own."""

    def __init__(self, rank: int, world: int, group=None, reduce_max=None):
        self.rank, self.world, self.group, self.reduce_max = int(rank), int(world), group, reduce_max
        if reduce_max is None and self.world > 1:
            if not dist.is_initialized():
                raise RuntimeError("shard=(rank, world) with world > 1 needs an initialised torch.distributed group")
            if dist.get_world_size(group) != self.world or dist.get_rank(group) != self.rank:
                raise RuntimeError("shard=(%d, %d) does not match the process group (rank %d of %d)" %
                                   (self.rank, self.world, dist.get_rank(group), dist.get_world_size(group)))

    def local(self, model, near: torch.Tensor, far: torch.Tensor) -> torch.Tensor:
        import ctypes as C
        from . import _lib
        dev = model._ensure_ctx()
        g = torch.empty(4, dtype=torch.float32, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        n = int(near.shape[0])
        _lib.check(model._lib.nt_shard_globals_local(model._ctx, n, C.c_void_p(near.data_ptr() if n else 0),
                                                     C.c_void_p(far.data_ptr() if n else 0), 1 if self.rank == 0 else 0,
                                                     C.c_void_p(g.data_ptr()), st))
        return g

    def resolve(self, model, g: torch.Tensor) -> torch.Tensor:
        import ctypes as C
        from . import _lib
        st = C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
        _lib.check(model._lib.nt_shard_globals_resolve(model._ctx, C.c_void_p(g.data_ptr()), st))
        return g

    def compute(self, model, near: torch.Tensor, far: torch.Tensor) -> torch.Tensor:
        g = self.local(model, near, far)
        if self.reduce_max is not None:
            self.reduce_max(g)
        elif self.world > 1:
            dist.all_reduce(g, op=dist.ReduceOp.MAX, group=self.group)
        return self.resolve(model, g)


def allreduce_sum_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """One collective over the flat 593 924-float gradient (2 375 696 B); NCCL over NVLink on GPUs."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All ranks' [n_local, C] outputs -> [n_total, C] in ray order (render output gather, 12 B/ray)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    first = shard_slice(n_total, 0, world)
    per = first.stop - first.start                     # every shard but the last ones has this many rows
    pad = torch.zeros(per, *local.shape[1:], dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat(out, dim=0)[:n_total]


class PeerGradExchange:
    """Gradient exchange over NVLink peer memory instead of an NCCL all-reduce.

    Every rank's flat gradient lives in a symmetric (peer-mapped) buffer.  After backward, `nt_adam_step_allreduce`
    reads all ranks' buffers directly (NVLink 5 / NVSwitch loads), sums them in rank order and applies Adam in ONE
    kernel; two device-side barriers bracket it.  Falls back to None (caller uses NCCL) when symmetric memory cannot be
    set up (single process, no P2P)."""

    def __init__(self, n_params: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.buf = symm_mem.empty(n_params, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.world = self.hdl.world_size
        self.rank = self.hdl.rank
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]       # peer-mapped base pointers, rank order

    @staticmethod
    def create(n_params: int, device: torch.device, group=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2 or device.type != "cuda":
            return None
        try:
            return PeerGradExchange(n_params, device, group)
        except Exception as e:       # pragma: no cover  (no P2P / unsupported build)
            import warnings
            warnings.warn(f"symmetric memory unavailable ({e}); using the NCCL all-reduce")
            return None

    def barrier(self):
        """Device-side cross-rank barrier on the current stream."""
        self.hdl.barrier(channel=0)
