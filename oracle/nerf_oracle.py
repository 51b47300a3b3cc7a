"""CPU oracle for the NeRF-tiny per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This file is a *restatement* (not a copy) of the algorithm in the reference's
``nerf.py``; every function cites the reference ``file:line`` it follows.  It
is written functionally (explicit weights dict, explicit arithmetic for the
bit-sensitive pieces) and runs on CPU with numpy + torch only, so it travels
to the GPU box where ``/root/reference`` does not exist.

Who may import this module: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product path
(``nerf_tiny_b200``) never imports it; it fails loudly if the CUDA library is
missing instead of falling back to this code.

Parity pin: the reference ships no tests / golden vectors, so the
oracle is pinned against *outputs of the reference itself run in the build
container*: ``oracle/make_golden.py`` imports ``/root/reference/nerf.py``,
drives it piecewise and end to end, asserts this restatement reproduces it
(bit-exact for ray dirs / t_coarse / searchsorted idx / t_fine / encoder; rgb
to <=1e-6) and writes ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
re-checks the oracle against those fixtures on every run.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32

# ---------------------------------------------------------------------------
# Weight table (state_dict keys of the reference, nerf.py:85-99)
# ---------------------------------------------------------------------------
LAYER_KEYS = [
    "network.point_layer.0.0", "network.point_layer.1.0", "network.point_layer.2.0",
    "network.point_layer.3.0", "network.point_layer.4.0", "network.point_layer.5.0",
    "network.point_layer.6.0", "network.point_layer.7.0", "network.sigma_layer.0",
    "network.point_info", "network.dir_info.0", "network.color_layer.0",
]
LAYER_SHAPES = [  # (out, in)
    (256, 60), (256, 256), (256, 256), (256, 256), (256, 316), (256, 256), (256, 256),
    (256, 256), (1, 256), (256, 256), (128, 280), (3, 128),
]
N_PARAMS = sum(o * i + o for o, i in LAYER_SHAPES)  # 593 924


def init_state_dict(seed: int = 624, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """nn.Linear default init, created in the reference's order (nerf.py:85-99).

    The *values* only equal the reference's when the global RNG is in the same
    state; parity tests share a state_dict instead of re-seeding.
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, (o, i) in zip(LAYER_KEYS, LAYER_SHAPES):
        bound = 1.0 / math.sqrt(i)
        sd[key + ".weight"] = ((torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        sd[key + ".bias"] = ((torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
    return sd


# ---------------------------------------------------------------------------
# a1/a2: pose decode and K_inv (nerf.py:52-67, nerf.py:433)
# ---------------------------------------------------------------------------
def poses_extract(pb: torch.Tensor):
    """nerf.py:52-67.  pb [N,17] -> c2w [N,4,4], H, W, focal, near[N], far[N]."""
    n = pb.shape[0]
    pose = pb[:, :15].reshape(n, 3, 5)
    bottom = torch.zeros(n, 1, 4, dtype=pb.dtype)
    bottom[:, 0, 3] = 1.0
    c2w = torch.cat((pose[:, :, :4], bottom), dim=1)
    return c2w, pose[0, 0, 4], pose[0, 1, 4], pose[0, 2, 4], pb[:, 15], pb[:, 16]


def make_k_inv(height: float, width: float, focal: float) -> torch.Tensor:
    """nerf.py:433 (already transposed): [x,y,1] @ K_inv = [x-W/2, -y+H/2, -f]."""
    k = torch.tensor([[1.0, 0.0, -0.5 * width], [0.0, -1.0, 0.5 * height], [0.0, 0.0, -focal]],
                     dtype=torch.float32)
    return k.t().contiguous()


# ---------------------------------------------------------------------------
# a7 / A.1: numpy linspace in fp32, batch-global any_step_zero branch
# ---------------------------------------------------------------------------
def linspace_rows(start: np.ndarray, stop: np.ndarray, num: int, force_last: bool = True,
                  any_step_zero=None, f32=np.float32) -> np.ndarray:
    """np.linspace(tuple(start), tuple(stop), num) as the reference calls it
    (nerf.py:288 and nerf.py:243), restated element by element:
    step = fl((stop-start)/(num-1)); y_i = fl(fl(i*step)+start); y_last = stop.
    If ANY row has step==0 every row uses y_i = fl(fl(fl(i/(num-1))*(stop-start))+start).
    Returns [N, num] fp32 (``f32=np.float64`` gives the fp64-oracle variant).
    """
    start = np.asarray(start, dtype=f32)
    stop = np.asarray(stop, dtype=f32)
    div = f32(num - 1)
    delta = (stop - start).astype(f32)
    step = (delta / div).astype(f32)
    idx = np.arange(num, dtype=f32)[None, :]
    if any_step_zero is None:
        any_step_zero = bool(np.any(step == 0))
    if any_step_zero:
        frac = (idx / div).astype(f32)
        y = ((frac * delta[:, None]).astype(f32) + start[:, None]).astype(f32)
    else:
        y = ((idx * step[:, None]).astype(f32) + start[:, None]).astype(f32)
    if force_last and num > 1:
        y[:, -1] = stop
    return y


def t_coarse_of(near: np.ndarray, far: np.ndarray, n_coarse: int = 64) -> np.ndarray:
    """nerf.py:288."""
    return linspace_rows(near, far, n_coarse)


# ---------------------------------------------------------------------------
# a3/a4 / A.2: ray directions and sample positions, explicit non-fused fp32
# ---------------------------------------------------------------------------
def ray_dirs(row: np.ndarray, col: np.ndarray, k_inv: np.ndarray, c2w: np.ndarray, f32=np.float32):
    """nerf.py:186-197, 211.  x := row, y := column (nerf.py:343-344).
    k_inv is the transposed matrix of nerf.py:433; the general 3x3 product is
    evaluated in the order ATen's (N,1,3)x(3,3) matmul uses: ((x*k0j + y*k1j) + 1*k2j).
    Returns dir_cam [N,3], dir_wrd [N,3] (fp32).
    """
    x = np.asarray(row).astype(f32)
    y = np.asarray(col).astype(f32)
    k = np.asarray(k_inv, dtype=f32)
    v = np.empty((x.shape[0], 3), dtype=f32)
    for j in range(3):
        v[:, j] = ((x * k[0, j]).astype(f32) + (y * k[1, j]).astype(f32)).astype(f32) + k[2, j]
    n2 = ((v[:, 0] * v[:, 0]).astype(f32) + (v[:, 1] * v[:, 1]).astype(f32)).astype(f32)
    n2 = (n2 + (v[:, 2] * v[:, 2]).astype(f32)).astype(f32)
    nrm = np.sqrt(n2).astype(f32)  # numpy sqrt is correctly rounded
    nrm = np.maximum(nrm, f32(1e-12))
    d_cam = (v / nrm[:, None]).astype(f32)
    rot = np.asarray(c2w, dtype=f32)[:, :3, :3]
    d_wrd = np.empty_like(d_cam)
    for r in range(3):
        acc = ((rot[:, r, 0] * d_cam[:, 0]).astype(f32) + (rot[:, r, 1] * d_cam[:, 1]).astype(f32)).astype(f32)
        d_wrd[:, r] = (acc + (rot[:, r, 2] * d_cam[:, 2]).astype(f32)).astype(f32)
    return d_cam, d_wrd


def sample_points(d_cam: np.ndarray, t: np.ndarray, c2w: np.ndarray) -> np.ndarray:
    """nerf.py:200-216: p_cam = fl(d_cam*t); p_wrd = ((R0*p0+R1*p1)+R2*p2)+T, no FMA."""
    c2w = np.asarray(c2w, dtype=f32)
    pc = (d_cam[:, None, :] * np.asarray(t, dtype=f32)[:, :, None]).astype(f32)
    out = np.empty_like(pc)
    for r in range(3):
        acc = ((c2w[:, None, r, 0] * pc[:, :, 0]).astype(f32) + (c2w[:, None, r, 1] * pc[:, :, 1]).astype(f32)).astype(f32)
        acc = (acc + (c2w[:, None, r, 2] * pc[:, :, 2]).astype(f32)).astype(f32)
        out[:, :, r] = (acc + c2w[:, None, r, 3]).astype(f32)
    return out


# ---------------------------------------------------------------------------
# a5 / A.4: encoder
# ---------------------------------------------------------------------------
def freq_table(L: int) -> np.ndarray:
    """nerf.py:141-145: exp2(linspace(0,L,L))*pi evaluated by torch in fp32."""
    return (torch.exp2(torch.linspace(0, L, L)) * math.pi).numpy().astype(f32)


FREQ_POINT_HEX = [0x40490fdb, 0x40d928ae, 0x416a8b6c, 0x41fd527b, 0x4288cd33,
                  0x4313c0fa, 0x439f953c, 0x442c5bef, 0x44ba2881, 0x45490fdb]
FREQ_DIR_HEX = [0x40490fdb, 0x40fd527a, 0x419f953c, 0x42490fdb]


def freq_from_hex(table) -> np.ndarray:
    return np.array(table, dtype=np.uint32).view(f32)


def encode(x: torch.Tensor, L: int, faithful32: bool = False) -> torch.Tensor:
    """nerf.py:135-167 + flatten at nerf.py:103-104: feature index c*2L + 2l + s,
    argument fl(w_l * x_c).  x [...,3] -> [...,6L].
    faithful32 (fp64 inputs only): the sin/cos ARGUMENT takes the value the fp32 path computes
    (fl32(w32 * fl32(x))) while autograd still differentiates the fp64 expression — the fp64
    gradient oracle for kernels that consume fp32 positions."""
    if x.dtype == torch.float32:
        w = torch.from_numpy(freq_from_hex(FREQ_POINT_HEX if L == 10 else FREQ_DIR_HEX))
    else:  # fp64 oracle: the reference then evaluates nerf.py:141-145 in double
        w = torch.exp2(torch.linspace(0, L, L, dtype=x.dtype)) * math.pi
    arg = x.unsqueeze(-1) * w  # [...,3,L]
    if faithful32 and x.dtype == torch.float64:
        w32 = torch.from_numpy(freq_from_hex(FREQ_POINT_HEX if L == 10 else FREQ_DIR_HEX))
        arg32 = (x.detach().float().unsqueeze(-1) * w32).double()
        arg = arg + (arg32 - arg).detach()
    out = torch.stack((torch.sin(arg), torch.cos(arg)), dim=-1)  # [...,3,L,2]
    return out.flatten(start_dim=-3)


# ---------------------------------------------------------------------------
# a6: the shared MLP (nerf.py:101-124)
# ---------------------------------------------------------------------------
def _bf16_st(x: torch.Tensor) -> torch.Tensor:
    """Round to bf16 with a straight-through gradient (value of bf16(x), derivative of x)."""
    return x + (x.detach().float().bfloat16().to(x.dtype) - x.detach())


def network_forward(sd: Dict[str, torch.Tensor], penc: torch.Tensor, denc: torch.Tensor, return_acts: bool = False,
                    emulate_bf16: bool = False):
    """penc [...,60], denc [...,24] -> color [...,3], sigma [...,1].
    emulate_bf16: model of the NT_PREC_BF16 kernels — the operands of the 10 tensor-core layers (activations and
    weights) are rounded to bf16, accumulation / biases / the sigma and colour heads stay in the working precision.
    Its autograd is the exact gradient of that rounded forward, i.e. what the bf16 training path must reproduce."""
    q = _bf16_st if emulate_bf16 else (lambda x: x)
    acts = []
    penc_q = q(penc)
    h = penc_q
    pre = None
    for i in range(8):
        k = f"network.point_layer.{i}.0"
        inp = torch.cat((h, penc_q), dim=-1) if i == 4 else h   # nerf.py:109 (hidden first)
        pre = torch.relu(F.linear(inp, q(sd[k + ".weight"]), sd[k + ".bias"]))
        h = q(pre)
        acts.append(pre)
    # the sigma head reads the un-rounded activations (fp32 accumulator registers in the kernel)
    sigma = torch.abs(F.linear(pre, sd["network.sigma_layer.0.weight"], sd["network.sigma_layer.0.bias"]))  # nerf.py:74
    info = F.linear(h, q(sd["network.point_info.weight"]), sd["network.point_info.bias"])   # nerf.py:117 (no act)
    u = torch.relu(F.linear(torch.cat((q(denc), q(info)), dim=-1), q(sd["network.dir_info.0.weight"]),
                            sd["network.dir_info.0.bias"]))                              # nerf.py:118 (dir first)
    color = torch.sigmoid(F.linear(u, sd["network.color_layer.0.weight"], sd["network.color_layer.0.bias"]))
    if return_acts:
        return color, sigma, acts + [info, u]
    return color, sigma


# ---------------------------------------------------------------------------
# a8/a9: compositing (nerf.py:263-281)
# ---------------------------------------------------------------------------
def get_density(delta: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
    a = delta * sigma
    s = torch.cumsum(a, dim=1)           # inclusive (nerf.py:267)
    return torch.exp(-s) * (1 - torch.exp(-a))


def color_cum(w: torch.Tensor, color: torch.Tensor) -> torch.Tensor:
    return torch.sum(color * w.unsqueeze(2), dim=1)


# ---------------------------------------------------------------------------
# a10 / A.3: inverse-CDF resampling (nerf.py:225-261)
# ---------------------------------------------------------------------------
class ResampleRangeError(RuntimeError):
    """The reference calls exit(0) here (nerf.py:251-253)."""


def resample(t_coarse: torch.Tensor, w: torch.Tensor, n_fine: int = 128, delta0=None,
             any_step_zero=None, return_aux: bool = False):
    """Differentiable w.r.t. ``w`` through cdf[idx] and slope[idx]; u and idx are constants."""
    cdf = torch.cumsum(w, dim=1).contiguous()
    hi = cdf.max(dim=1).values.detach().numpy()
    lo = cdf.min(dim=1).values.detach().numpy()
    if delta0 is None:
        delta0 = t_coarse[0, 1] - t_coarse[0, 0]          # ray 0 only (nerf.py:234)
    slope = delta0 / (w[:, 1:] + 1e-7)
    npdt = np.float64 if w.dtype == torch.float64 else f32
    u_np = linspace_rows(lo, hi, n_fine + 2, any_step_zero=any_step_zero, f32=npdt)[:, 1:-1]
    u = torch.from_numpy(np.ascontiguousarray(u_np)).to(w.dtype)
    idx = torch.searchsorted(cdf.detach(), u) - 1          # left: #{cdf<u}-1
    if bool((idx < 0).any()) or bool((idx > n_fine - 1).any()):
        raise ResampleRangeError("resample index out of range (reference exits at nerf.py:253)")
    slope = torch.cat((slope, torch.zeros(w.shape[0], 1, dtype=w.dtype)), dim=1)
    t_fine = torch.gather(t_coarse, 1, idx) + (u - torch.gather(cdf, 1, idx)) * torch.gather(slope, 1, idx)
    if return_aux:
        return t_fine, idx, u, cdf
    return t_fine


# ---------------------------------------------------------------------------
# a11: merge + per-channel sort + fine compositing (nerf.py:302-321)
# ---------------------------------------------------------------------------
def merge_sort_composite(t_c, color_c, sigma_c, t_f, color_f, sigma_f, last: float = 1e-4):
    """Each of the 5 channels (t,r,g,b,sigma) is sorted independently (nerf.py:308)."""
    t = torch.cat((t_c, t_f), dim=1)
    color = torch.cat((color_c, color_f), dim=1)
    sigma = torch.cat((sigma_c, sigma_f), dim=1)
    t_s = torch.sort(t, dim=1).values
    color_s = torch.sort(color, dim=1).values
    sigma_s = torch.sort(sigma, dim=1).values
    delta = torch.cat((t_s[:, 1:] - t_s[:, :-1], torch.full((t.shape[0], 1), last, dtype=t.dtype)), dim=1)
    w = get_density(delta, sigma_s)
    return color_cum(w, color_s), w, t_s, color_s, sigma_s


# ---------------------------------------------------------------------------
# a13: full forward (nerf.py:286-348)
# ---------------------------------------------------------------------------
def net_out(sd, t: torch.Tensor, d_cam: np.ndarray, d_wrd: np.ndarray, c2w: np.ndarray, t_requires_path: bool = False,
            faithful32: bool = False, emulate_bf16: bool = False):
    """nerf.py:179-222 for a [N,P] array of t.  If ``t`` carries grad the sample
    positions are rebuilt in torch (p = R*(d_cam*t)+T) so autograd reaches t."""
    n, p = t.shape
    dt = t.dtype
    # same op order as sample_points() (each torch op rounds once, no FMA) so values are
    # bit-identical to the numpy restatement while autograd still reaches t
    dc = torch.from_numpy(d_cam).to(dt)
    c = torch.from_numpy(np.asarray(c2w)).to(dt)
    pc = dc[:, None, :] * t[:, :, None]
    cols = []
    for r in range(3):
        acc = c[:, None, r, 0] * pc[:, :, 0] + c[:, None, r, 1] * pc[:, :, 1]
        acc = acc + c[:, None, r, 2] * pc[:, :, 2]
        cols.append(acc + c[:, None, r, 3])
    pts = torch.stack(cols, dim=-1)
    if faithful32 and dt == torch.float64:   # positions take their fp32 values, gradients stay fp64
        p32 = sample_points(d_cam.astype(f32), t.detach().float().numpy(), np.asarray(c2w, dtype=f32))
        pts = pts + (torch.from_numpy(p32).double() - pts).detach()
    dirs = torch.from_numpy(d_wrd).to(dt)[:, None, :].expand(n, p, 3)
    color, sigma = network_forward(sd, encode(pts, 10, faithful32), encode(dirs, 4, faithful32), emulate_bf16=emulate_bf16)
    return color, sigma.squeeze(-1)


def render_rays(sd, row, col, c2w, k_inv, near, far, n_coarse: int = 64, n_fine: int = 128,
                last: float = 1e-4, return_aux: bool = False, any_step_zero=None, delta0=None,
                detach_t_fine: bool = False, emulate_bf16: bool = False, faithful32: bool = False):
    """nerf.py:286-323.  row/col int arrays [N]; c2w [N,4,4]; near/far [N] (fp32 numpy)."""
    near = np.asarray(near, dtype=f32)
    far = np.asarray(far, dtype=f32)
    c2w = np.asarray(c2w, dtype=f32)
    dt = next(iter(sd.values())).dtype
    npdt = np.float64 if dt == torch.float64 else f32
    d_cam, d_wrd = ray_dirs(row, col, np.asarray(k_inv, dtype=f32), c2w, f32=npdt)
    # any_step_zero / delta0: batch-global quantities a ray shard is told explicitly
    t_c = torch.from_numpy(linspace_rows(near, far, n_coarse, any_step_zero=any_step_zero)).to(dt)
    color_c, sigma_c = net_out(sd, t_c, d_cam, d_wrd, c2w, faithful32=faithful32, emulate_bf16=emulate_bf16)
    delta_c = torch.from_numpy(((far - near) / f32(n_coarse)).astype(f32)).to(dt)[:, None].expand(-1, n_coarse)
    w_c = get_density(delta_c, sigma_c)                                   # nerf.py:293-295
    t_f, idx, u, cdf = resample(t_c, w_c, n_fine, delta0=delta0, return_aux=True)        # nerf.py:298
    if detach_t_fine:   # diagnostic only: the reference does NOT detach (nerf.py:255-259)
        t_f = t_f.detach()
    color_f, sigma_f = net_out(sd, t_f, d_cam, d_wrd, c2w, t_requires_path=True, faithful32=faithful32,
                               emulate_bf16=emulate_bf16)
    c_fine, w_f, t_s, _, _ = merge_sort_composite(t_c, color_c, sigma_c, t_f, color_f, sigma_f, last)
    c_coarse = color_cum(w_c, color_c)
    if return_aux:
        aux = dict(d_cam=d_cam, d_wrd=d_wrd, t_coarse=t_c, color_c=color_c, sigma_c=sigma_c, w_c=w_c,
                   t_fine=t_f, idx=idx, u=u, cdf=cdf, color_f=color_f, sigma_f=sigma_f, w_f=w_f, t_sorted=t_s)
        return c_coarse, c_fine, aux
    return c_coarse, c_fine


def forward(sd, row, col, poses_bound, k_inv, n_coarse: int = 64, n_fine: int = 128, return_aux: bool = False,
            any_step_zero=None, delta0=None, **kw):
    """nerf.py:333-348.  poses_bound [N,17] (float64 from the loader), cast to fp32 first."""
    pb = torch.as_tensor(poses_bound).to(torch.float32)
    c2w, _, _, _, near, far = poses_extract(pb)
    return render_rays(sd, np.asarray(row), np.asarray(col), c2w.numpy(), torch.as_tensor(k_inv).numpy(),
                       near.numpy(), far.numpy(), n_coarse, n_fine, return_aux=return_aux,
                       any_step_zero=any_step_zero, delta0=delta0, **kw)


def ray_loss(c_coarse, c_fine, c_true):
    """nerf.py:325-331 (sum, not mean)."""
    return torch.sum(torch.square(c_coarse - c_true)) + torch.sum(torch.square(c_fine - c_true))


def psnr(c_fine, c_true) -> float:
    """Harness-defined (the reference never computes PSNR): -10 log10 mean((C_fine-C_true)^2)."""
    return float(-10.0 * torch.log10(torch.mean(torch.square(c_fine - c_true))))


# ---------------------------------------------------------------------------
# a15: Adam + LR schedule (nerf.py:425-427, 474-475)
# ---------------------------------------------------------------------------
def lr_lambda(it: int, lr_gamma: float, decay_end: float, learning: float) -> float:
    """nerf.py:426: gamma**(it/decay_end) before decay_end, then gamma*learning (sic)."""
    return lr_gamma ** (it / decay_end) if it < decay_end else lr_gamma * learning


def adam_step(p, g, m, v, step: int, lr: float, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-7):
    """torch.optim.Adam single-tensor math (no weight decay / amsgrad), in place. step is 1-based."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# ---------------------------------------------------------------------------
# Synthetic inputs shared by tests / bench
# ---------------------------------------------------------------------------
def trained_like(sd: Dict[str, torch.Tensor], sigma_gain: float = 400.0, seed: int = 1) -> Dict[str, torch.Tensor]:
    """A 'trained-like' stress state: sigma spans 0..tens, colours spread over (0,1)."""
    g = torch.Generator().manual_seed(seed)
    out = {k: v.clone() for k, v in sd.items()}
    out["network.sigma_layer.0.weight"] *= sigma_gain
    out["network.sigma_layer.0.bias"] *= sigma_gain
    out["network.color_layer.0.weight"] *= 30.0
    for i in range(8):
        k = f"network.point_layer.{i}.0.weight"
        out[k] = out[k] * 1.8 + 0.02 * torch.randn(out[k].shape, generator=g)
    return out
