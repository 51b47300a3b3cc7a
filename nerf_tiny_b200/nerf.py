"""Drop-in for the reference's `nerf.py` module surface, backed by libnerftiny.so.

Same names, positional order and defaults as the reference:
  seed_everything (nerf.py:43), poses_extract (:52), Activation (:69), Network (:76), Encoder (:126),
  NeRFModel (:169) with net_out / resample / get_density / color_cum / render_rays / ray_loss / forward,
  NeRFRunner (:353).  Module globals `device` and `writer` behave as in nerf.py:39-40.

PyTorch is used here for device memory, streams, parameter bookkeeping and autograd *plumbing* only;
every arithmetic step of the hot path runs in the hand-written CUDA kernels behind the C-ABI
(include/nerftiny.h).  There is no CPU fallback: constructing a NeRFModel without the shared library or
without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import pickle as _pickle
import random
import time

import numpy as np
import torch
import torch.nn as nn

from . import _lib

writer = None
device = None

# MLP arithmetic (include/nerftiny.h NT_PREC_*).  "fp16" / "bf16": single-pass 16-bit tcgen05 kernels (fp16 operands round 8x
# finer than bf16 at the same speed; the network's activations are bounded, so fp16 is the default for rendering);
# "tc32": 3-pass split-fp16 tcgen05 kernel meeting the fp32 tolerance on the tensor cores; "fp32": CUDA-core FFMA.
# Training runs on the bf16 tensor-core path for "bf16" / "fp16" and on the fp32 path for "fp32" / "tc32".
# "mixed": coarse pass in tc32 arithmetic (its weights steer the resampling, which amplifies their errors), fine pass fp16.
PRECISION = {"fp32": _lib.PREC_FP32, "tc32": _lib.PREC_TC32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16,
             "mixed": _lib.PREC_MIXED}
TRAIN_PRECISION = {"fp32": _lib.PREC_FP32, "tc32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_BF16,
                   "mixed": _lib.PREC_BF16}
DEFAULT_PRECISION = "fp16"


def seed_everything(seed):
    """nerf.py:43-48."""
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def poses_extract(pb_matrix):
    """nerf.py:52-67: [N,17] -> (c_to_w [N,4,4], H, W, focal, near [N], far [N]); pure data movement."""
    n = pb_matrix.shape[0]
    pose = pb_matrix[:, :-2].reshape(-1, 3, 5)
    bottom = torch.zeros(n, 1, 4, dtype=pb_matrix.dtype, device=pb_matrix.device)
    bottom[:, 0, 3] = 1.0
    c_to_w = torch.cat((pose[:, :, :-1], bottom), dim=1)
    return c_to_w, pose[0, 0, -1], pose[0, 1, -1], pose[0, 2, -1], pb_matrix[:, -2], pb_matrix[:, -1]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(dev=None):
    """The caller's current stream ON `dev` (the model's device), whatever device is current."""
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _dev_f32(t, dev):
    return t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()


_MODULE_CTX = {}


def _module_ctx(dev):
    """A library context for stand-alone sub-module calls (Encoder / Network used outside a NeRFModel)."""
    if dev.type != "cuda":
        raise _lib.NerfTinyError("this module lives on %s: move it to a CUDA device (no CPU path exists)" % dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _MODULE_CTX:
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.nt_create(C.byref(h), idx, 64, 128))
        _MODULE_CTX[idx] = (lib, h)
    return _MODULE_CTX[idx]


class Activation(nn.Module):
    """nerf.py:69-74 (sigma = |x|).  Inside Network / NeRFModel the abs is fused into the CUDA MLP kernels; called on its
    own it is the reference's one-liner (nerf.py:73-74)."""

    def forward(self, x):
        return torch.abs(x)


class Network(nn.Module):
    """nerf.py:76-99: same module tree, hence the same state_dict keys and nn.Linear default init.

    All 24 tensors are views into ONE flat fp32 buffer in state_dict order (the layout the C-ABI uses), so the
    fused Adam and the gradient all-reduce each touch a single 2.4 MB buffer.
    """

    def __init__(self, point_dim=60, dir_dim=24, depth=8, width=256, batch_size=8, layers_skip=[4]):
        super().__init__()
        if (point_dim, dir_dim, depth, width, list(layers_skip)) != (60, 24, 8, 256, [4]):
            raise _lib.NerfTinyError("libnerftiny is built for the 8x256 skip-4 network of nerf.py:77")
        self.depth, self.width, self.batch_size, self.layers_skip = depth, width, batch_size, layers_skip
        self.point_layer = nn.ModuleList([nn.Sequential(nn.Linear(point_dim, width), nn.ReLU(True))])
        for i in range(1, depth):
            fan_in = width + point_dim if i in layers_skip else width
            self.point_layer.append(nn.Sequential(nn.Linear(fan_in, width), nn.ReLU(True)))
        self.sigma_layer = nn.Sequential(nn.Linear(width, 1), Activation())
        self.point_info = nn.Linear(width, width)
        self.dir_info = nn.Sequential(nn.Linear(width + dir_dim, width // 2), nn.ReLU(True))
        self.color_layer = nn.Sequential(nn.Linear(width // 2, 3), nn.Sigmoid())
        self._flat = None
        self._flat_grad = None
        self._grad_external = None
        self._flatten()

    # -- flat parameter storage ------------------------------------------------------------------
    def _flatten(self, dev=None):
        params = list(self.parameters())
        dev = dev if dev is not None else params[0].device
        flat = torch.empty(_lib.N_PARAMS, dtype=torch.float32, device=dev)
        ext = getattr(self, "_grad_external", None)
        if ext is not None and ext.device == flat.device:
            grad = ext.zero_()          # keep an installed gradient buffer (symmetric / peer-mapped memory) across re-flattens
        else:
            self._grad_external = None
            grad = torch.zeros(_lib.N_PARAMS, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            p.grad = grad[off:off + n].view(p.shape)
            off += n
        assert off == _lib.N_PARAMS
        self._flat, self._flat_grad = flat, grad

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._flatten()
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._flatten()
        return out

    def flat_params(self):
        p0 = next(self.parameters())
        if p0.data_ptr() != self._flat.data_ptr() or p0.device != self._flat.device:
            self._flatten()
        return self._flat

    def use_grad_buffer(self, buf):
        """Make `buf` (flat fp32 [593924], e.g. a symmetric / peer-mapped allocation) the gradient storage."""
        assert buf.numel() == _lib.N_PARAMS and buf.dtype == torch.float32
        buf.copy_(self._flat_grad)
        self._flat_grad = buf
        self._grad_external = buf
        off = 0
        for p in self.parameters():
            n = p.numel()
            p.grad = buf[off:off + n].view(p.shape)
            off += n

    def flat_grads(self):
        """Gather .grad into the flat buffer (no copy when the grads are already its views)."""
        off = 0
        for p in self.parameters():
            n = p.numel()
            g = p.grad
            if g is None:
                self._flat_grad[off:off + n].zero_()
                p.grad = self._flat_grad[off:off + n].view(p.shape)
            elif g.data_ptr() != self._flat_grad.data_ptr() + 4 * off:
                self._flat_grad[off:off + n].copy_(g.reshape(-1))
                p.grad = self._flat_grad[off:off + n].view(p.shape)
            off += n
        return self._flat_grad

    def forward(self, num_points, point, dir):
        """nerf.py:101-124 on materialised encodings point [N,P,3,20], dir [N,P,3,8] -> (color [N,P,3], sigma [N,P,1]).
        fp32 accuracy path (nt_network_forward), no autograd: training goes through NeRFModel.forward, whose fused kernels
        never write the encodings to memory."""
        flat = self.flat_params()
        lib, h = _module_ctx(flat.device)
        n, p = point.shape[0], point.shape[1]
        enc_p = _dev_f32(point, flat.device).reshape(n * p, 60)      # flatten(start_dim=2): channel-major (nerf.py:103)
        enc_d = _dev_f32(dir, flat.device).reshape(n * p, 24)
        color = torch.empty(n, p, 3, dtype=torch.float32, device=flat.device)
        sigma = torch.empty(n, p, 1, dtype=torch.float32, device=flat.device)
        need = max(256, lib.nt_mlp_workspace_bytes(h, _lib.PREC_FP32, n * p, 1, 0))
        ws = torch.empty(need, dtype=torch.uint8, device=flat.device)
        _lib.check(lib.nt_network_forward(h, n * p, _ptr(enc_p), _ptr(enc_d), _ptr(flat.detach()), _ptr(color), _ptr(sigma),
                                          _ptr(ws), ws.numel(), _stream(flat.device)))
        return color, sigma


class Encoder(nn.Module):
    """nerf.py:126-167.  The sin/cos features are generated inside the MLP kernels (never materialised);
    the class is kept so the module tree matches.  Like the reference it has no parameters."""

    def __init__(self, L_point=10, L_dir=4, batch_size=8):
        super().__init__()
        if (L_point, L_dir) != (10, 4):
            raise _lib.NerfTinyError("libnerftiny is built for L_point=10, L_dir=4 (nerf.py:127)")
        self.L_point, self.L_dir, self.batch_size = L_point, L_dir, batch_size

    def forward(self, num_points, point, dir):
        """nerf.py:135-167: point, dir [N,P,3] -> (gamma_point [N,P,3,20], gamma_dir [N,P,3,8]) on the CUDA device.  Like the
        reference, 14 numbers are drawn from the global CPU generator and thrown away (nerf.py:141) so that code sharing
        that generator sees the same stream.  The fused render path does not call this (it never materialises them)."""
        torch.rand(1, 1, self.L_point + self.L_dir, 1, 1)
        dev = device if (device is not None and device.type == "cuda") else torch.device("cuda", torch.cuda.current_device())
        lib, h = _module_ctx(dev)
        n, p = point.shape[0], point.shape[1]
        pt, dr = _dev_f32(point, dev), _dev_f32(dir, dev)
        g_p = torch.empty(n, p, 3, 2 * self.L_point, dtype=torch.float32, device=dev)
        g_d = torch.empty(n, p, 3, 2 * self.L_dir, dtype=torch.float32, device=dev)
        _lib.check(lib.nt_encode(h, n * p, _ptr(pt), _ptr(dr), _ptr(g_p), _ptr(g_d), _stream(dev)))
        return g_p, g_d


class _RenderFn(torch.autograd.Function):
    """render_rays forward/backward as one autograd node (nerf.py:286-323 + nerf.py:473)."""

    @staticmethod
    def forward(ctx, model, flat, row, col, pose17, kinv, near, far, any_step_zero, delta0):
        cc, cf, ws = model._render_raw(flat, row, col, pose17, kinv, near, far, train=True, any_step_zero=any_step_zero,
                                       delta0=delta0)
        ctx.model, ctx.ws, ctx.near, ctx.far, ctx.flat, ctx.delta0 = model, ws, near, far, flat, delta0
        return cc, cf

    @staticmethod
    def backward(ctx, g_cc, g_cf):
        m = ctx.model
        grads = torch.zeros_like(ctx.flat)
        n = ctx.near.shape[0]
        g_cc = g_cc.contiguous().float()
        g_cf = g_cf.contiguous().float()
        prec = m._prec_train
        _lib.check(m._lib.nt_render_backward(m._ctx, prec, n, _ptr(ctx.near), _ptr(ctx.far), _ptr(ctx.flat),
                                             _ptr(m._packed.get(prec)), _ptr(ctx.delta0), _ptr(g_cc), _ptr(g_cf), _ptr(grads),
                                             _ptr(ctx.ws), ctx.ws.numel(), _stream(ctx.flat.device)))
        ctx.ws = None
        return None, grads, None, None, None, None, None, None, None, None


class _NetOutFn(torch.autograd.Function):
    """net_out (nerf.py:179-222) as an autograd node: nt_mlp_forward with the training stash, nt_mlp_backward for the
    parameter gradients and the gradient w.r.t. the sample positions t (nerf.py:200)."""

    @staticmethod
    def forward(ctx, model, flat, t, rays, denc):
        prec = model._prec_train
        n, p = t.shape
        L, dev = model._lib, t.device
        packed = model._pack(flat, prec)
        need = L.nt_mlp_workspace_bytes(model._ctx, prec, n, p, 1)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        rgb = torch.empty(n, p, 3, device=dev)
        sigma = torch.empty(n, p, device=dev)
        _lib.check(L.nt_mlp_forward(model._ctx, prec, n, p, _ptr(t), _ptr(rays), _ptr(denc), _ptr(flat), _ptr(packed),
                                    _ptr(rgb), _ptr(sigma), _ptr(ws), ws.numel(), 1, _stream(dev)))
        ctx.model, ctx.ws, ctx.flat, ctx.t, ctx.rays, ctx.denc, ctx.rgb, ctx.prec = model, ws, flat, t, rays, denc, rgb, prec
        return rgb, sigma.unsqueeze(-1)

    @staticmethod
    def backward(ctx, g_rgb, g_sigma):
        m, t = ctx.model, ctx.t
        n, p = t.shape
        dev = t.device
        g_rgb = g_rgb.contiguous().float()
        g_sigma = g_sigma.reshape(n, p).contiguous().float()
        grads = torch.zeros_like(ctx.flat)
        g_t = torch.empty(n, p, device=dev)
        # the weights must be the ones the forward saw: re-pack (cheap) in case another call packed different ones since
        packed = m._pack(ctx.flat, ctx.prec)
        _lib.check(m._lib.nt_mlp_backward(m._ctx, ctx.prec, n, p, _ptr(t), _ptr(ctx.rays), _ptr(ctx.denc), _ptr(ctx.flat),
                                          _ptr(packed), _ptr(ctx.rgb), _ptr(g_rgb), _ptr(g_sigma), _ptr(grads), _ptr(g_t),
                                          _ptr(ctx.ws), ctx.ws.numel(), _stream(dev)))
        ctx.ws = None
        return None, grads, g_t, None, None


class _DensityFn(torch.autograd.Function):
    """get_density (nerf.py:263-272): forward = nt_get_density; backward in closed form with a = sigma * delta,
    w_i = T_i (1 - e^{-a_i}), T_i = exp(-sum_{j<=i} a_j):  dL/da_j = g_j T_j e^{-a_j} - sum_{i>=j} g_i w_i."""

    @staticmethod
    def forward(ctx, model, delta, sigma):
        n, p = delta.shape
        w = torch.empty(n, p, dtype=torch.float32, device=delta.device)
        _lib.check(model._lib.nt_get_density(model._ctx, n, p, _ptr(delta), _ptr(sigma), _ptr(w), _stream(delta.device)))
        ctx.save_for_backward(delta, sigma, w)
        return w

    @staticmethod
    def backward(ctx, g):
        delta, sigma, w = ctx.saved_tensors
        a = sigma * delta
        T = torch.exp(-torch.cumsum(a.double(), -1)).float()
        gw = g * w
        tail = torch.flip(torch.cumsum(torch.flip(gw.double(), [-1]), -1), [-1]).float()   # sum_{i>=j} g_i w_i
        g_a = g * T * torch.exp(-a) - tail
        return None, g_a * sigma, g_a * delta


class _ColorCumFn(torch.autograd.Function):
    """color_cum (nerf.py:274-281): C = sum_i w_i rgb_i."""

    @staticmethod
    def forward(ctx, model, w, c):
        n, p = w.shape
        out = torch.empty(n, 3, dtype=torch.float32, device=w.device)
        _lib.check(model._lib.nt_color_cum(model._ctx, n, p, _ptr(w), _ptr(c), _ptr(out), _stream(w.device)))
        ctx.save_for_backward(w, c)
        return out

    @staticmethod
    def backward(ctx, g):
        w, c = ctx.saved_tensors
        return None, (c * g[:, None, :]).sum(-1), w[..., None] * g[:, None, :]


class _ResampleFn(torch.autograd.Function):
    """resample (nerf.py:225-261): forward = nt_sample_pdf, backward = nt_sample_pdf_backward (the gradient
    reaches the coarse weights through cdf[idx] and the slope; u, idx, delta0 and t_coarse carry none)."""

    @staticmethod
    def forward(ctx, model, t_coarse, w):
        n = t_coarse.shape[0]
        out = torch.empty(n, model.num_fine, device=t_coarse.device)
        _lib.check(model._lib.nt_sample_pdf(model._ctx, n, _ptr(t_coarse), _ptr(w), None, _ptr(out), None, _stream(t_coarse.device)))
        ctx.model = model
        ctx.save_for_backward(t_coarse, w)
        return out

    @staticmethod
    def backward(ctx, g):
        t_coarse, w = ctx.saved_tensors
        m = ctx.model
        g = g.contiguous().float()
        g_w = torch.empty_like(w)
        _lib.check(m._lib.nt_sample_pdf_backward(m._ctx, t_coarse.shape[0], _ptr(t_coarse), _ptr(w), None, _ptr(g), _ptr(g_w),
                                                 _stream(w.device)))
        return None, None, g_w


class NeRFModel(nn.Module):
    """nerf.py:169-348.  `precision` is the one extra knob (see PRECISION above; default "fp16", or $NERF_TINY_PRECISION).

    Ray sharding: `forward`, `render_rays`, `train_step` and `GraphedTrainStep` take `shard=`
      None                      this call holds the whole batch (the reference's situation);
      (rank, world)             this call holds rank's contiguous slice of a batch sharded over `world` processes of the
                                default torch.distributed group: the two batch-global quantities (numpy's any_step_zero
                                branch, nerf.py:288, and delta0 of global ray 0, nerf.py:234) are reduced on the device
                                (one 16-byte MAX all-reduce, no host sync) so that sharded == unsharded bit for bit;
      a dist.ShardGlobals       the same with an explicit group / reduction;
      {"any_step_zero": 0|1, "delta0": tensor[1]}   explicit values.
    """

    def __init__(self, num_coarse=64, num_fine=128, batch_ray=8, precision=None):
        super().__init__()
        self.encoder = Encoder(batch_size=batch_ray)
        self.network = Network(batch_size=batch_ray)
        self.num_coarse, self.num_fine, self.batch_ray = num_coarse, num_fine, batch_ray
        self.precision = precision or os.environ.get("NERF_TINY_PRECISION", DEFAULT_PRECISION)
        if self.precision not in PRECISION:
            raise _lib.NerfTinyError("precision must be one of %s" % sorted(PRECISION))
        self.check_range = True          # raise (the reference: exit(0)) right after forward
        self._init_runtime()

    def _init_runtime(self):
        self._lib = _lib.load()          # raises if the extension is missing: no fallback
        self._ctx = None
        self._ctx_dev = None
        self._packed = {}                # precision code -> packed weight image
        self._ws_cache = {}
        self._shards = {}
        self._last = 1e-4

    # -- context -----------------------------------------------------------------------------------
    def _ensure_ctx(self):
        dev = self.network.flat_params().device
        if dev.type != "cuda":
            raise _lib.NerfTinyError("NeRFModel lives on %s: move it to a CUDA device (no CPU path exists)" % dev)
        if self._ctx is None or self._ctx_dev != dev:
            if self._ctx is not None:
                self._lib.nt_destroy(self._ctx)
            h = C.c_void_p()
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            dev = torch.device("cuda", idx)
            _lib.check(self._lib.nt_create(C.byref(h), idx, self.num_coarse, self.num_fine))
            self._ctx, self._ctx_dev = h, dev
            self._packed, self._ws_cache, self._last = {}, {}, 1e-4
        return self._ctx_dev

    def __del__(self):
        try:
            if getattr(self, "_ctx", None) is not None:
                self._lib.nt_destroy(self._ctx)
        except Exception:
            pass

    _RUNTIME_ATTRS = ("_lib", "_ctx", "_ctx_dev", "_packed", "_ws_cache", "_shards", "_last")

    def __getstate__(self):
        """torch.save(model) as the reference does (nerf.py:491): drop the library handle, context and scratch."""
        st = dict(self.__dict__)
        for k in self._RUNTIME_ATTRS:
            st.pop(k, None)
        return st

    def __setstate__(self, st):
        """Also accepts the state of a whole-module pickle written by the reference itself (nerf.py:491): it has no
        `precision` / `check_range`, and its Network has no flat buffer yet."""
        self.__dict__.update(st)
        self.__dict__.setdefault("precision", os.environ.get("NERF_TINY_PRECISION", DEFAULT_PRECISION))
        self.__dict__.setdefault("check_range", True)
        self._init_runtime()
        self.network._flatten()

    @property
    def _prec(self):
        return PRECISION[self.precision]

    @property
    def _prec_train(self):
        return TRAIN_PRECISION[self.precision]

    @property
    def launch_count(self):
        return int(self._lib.nt_launch_count(self._ctx)) if self._ctx is not None else 0

    def _workspace(self, n, train, prec):
        need = self._lib.nt_render_workspace_bytes(self._ctx, prec, n, 1 if train else 0)
        if train:   # owned by the autograd node
            return torch.empty(need, dtype=torch.uint8, device=self._ctx_dev)
        ws = self._ws_cache.get("render")
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self._ctx_dev)
            self._ws_cache["render"] = ws
        return ws

    def _pack(self, flat, prec):
        """fp32 flat parameters -> the packed operand image of `prec` (bf16 / fp16 / split fp16); None for fp32."""
        if prec == _lib.PREC_FP32:
            return None
        buf = self._packed.get(prec)
        if buf is None:
            buf = torch.empty(max(256, self._lib.nt_packed_weight_bytes(self._ctx, prec)), dtype=torch.uint8,
                              device=self._ctx_dev)
            self._packed[prec] = buf
        _lib.check(self._lib.nt_pack_weights(self._ctx, prec, _ptr(flat), _ptr(buf), _stream(self._ctx_dev)))
        return buf

    def _set_last(self, last):
        last = float(last)
        if last != self._last:
            bits = int(np.float32(last).view(np.int32))
            _lib.check(self._lib.nt_set_option(self._ctx, _lib.OPT_LAST_DELTA, bits))
            self._last = last

    def _globals(self, shard, near, far):
        """-> (any_step_zero argument, delta0 device tensor or None) for nt_render_forward / nt_render_backward."""
        if shard is None:
            return -1, None
        if isinstance(shard, dict):
            d0 = shard["delta0"]
            d0 = d0 if torch.is_tensor(d0) else torch.tensor([float(d0)], dtype=torch.float32)
            return int(bool(shard["any_step_zero"])), d0.to(self._ctx_dev, torch.float32).reshape(-1).contiguous()
        if isinstance(shard, tuple):
            from . import dist as D
            key = tuple(shard)
            if key not in self._shards:
                self._shards[key] = D.ShardGlobals(*key)
            shard = self._shards[key]
        return _lib.ANY_STEP_ZERO_DEVICE, shard.compute(self, near, far)

    def _render_raw(self, flat, row, col, pose17, kinv, near, far, train, any_step_zero=-1, delta0=None):
        n = row.shape[0]
        dev = self._ctx_dev
        prec = self._prec_train if train else self._prec
        cc = torch.empty(n, 3, dtype=torch.float32, device=dev)
        cf = torch.empty(n, 3, dtype=torch.float32, device=dev)
        ws = self._workspace(n, train, prec)
        packed = self._pack(flat, prec)
        stride = 17 if pose17.shape[-1] == 17 else pose17[0].numel()
        _lib.check(self._lib.nt_render_forward(self._ctx, prec, n, _ptr(row), _ptr(col), _ptr(pose17), stride,
                                               _ptr(kinv), _ptr(near), _ptr(far), _ptr(flat), _ptr(packed),
                                               any_step_zero, _ptr(delta0), _ptr(cc), _ptr(cf), _ptr(ws), ws.numel(),
                                               1 if train else 0, _stream(dev)))
        return cc, cf, ws

    # -- reference methods -----------------------------------------------------------------------------
    def render_rays(self, batch_hor, batch_ver, trans_mat, K_inv, near, far, last=0.0001, shard=None):
        """nerf.py:286-323.  trans_mat [N,4,4] (or fp32 pose rows [N,17]); near/far [N]."""
        dev = self._ensure_ctx()
        self._set_last(last)
        row = batch_hor.to(dev, non_blocking=True).to(torch.int64).contiguous()
        col = batch_ver.to(dev, non_blocking=True).to(torch.int64).contiguous()
        pose = _dev_f32(trans_mat, dev)
        kinv = _dev_f32(K_inv, dev)
        near = _dev_f32(near, dev)
        far = _dev_f32(far, dev)
        flat = self.network.flat_params()
        asz, d0 = self._globals(shard, near, far)
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.network.parameters())
        if needs_grad:
            cc, cf = _RenderFn.apply(self, _FlatView.apply(flat, *self.network.parameters()), row, col, pose, kinv, near, far,
                                     asz, d0)
        else:
            cc, cf, _ = self._render_raw(flat, row, col, pose, kinv, near, far, train=False, any_step_zero=asz, delta0=d0)
        if self.check_range:
            _lib.check(self._lib.nt_check_status(self._ctx, _stream(dev)))
        return cc, cf

    def forward(self, row, column, poses_bound, K_inv, shard=None):
        """nerf.py:333-348: row/column int64 [N] and poses_bound float64 [N,17] arrive on the CPU from the loader."""
        dev = self._ensure_ctx()
        pb = poses_bound.to(dev, non_blocking=True).to(torch.float32).contiguous()   # nerf.py:338
        near = pb[:, 15].contiguous()
        far = pb[:, 16].contiguous()
        return self.render_rays(row, column, pb, K_inv, near, far, shard=shard)

    def ray_loss(self, C_coarse, C_fine, C_true):
        """nerf.py:325-331.  Differentiable (autograd seeds 2(C-Ct) for the render node)."""
        return _RayLossFn.apply(self, C_coarse, C_fine, C_true.to(C_coarse.device, torch.float32))

    def check_status(self):
        _lib.check(self._lib.nt_check_status(self._ctx, _stream(self._ctx_dev)))

    def set_detach_t_fine(self, on: bool):
        """Diagnostic (NT_OPT_DETACH_T_FINE): treat t_fine as a constant in backward.  The reference does not."""
        self._ensure_ctx()
        _lib.check(self._lib.nt_set_option(self._ctx, 1, 1 if on else 0))

    # -- piecewise methods (same math, single kernels) ---------------------------------------------
    def net_out(self, t_array, batch_x, batch_y, trans_mat, K_inv, num_points):
        """nerf.py:179-222 -> (color [N,P,3], sigma [N,P,1]).  Differentiable like the reference's: with grad enabled and a
        trainable network (or a t_array that requires grad) the call runs the training kernels (precision bf16, or fp32 for the
        fp32 / tc32 models) and records an autograd node; otherwise the rendering kernels of `precision`."""
        dev = self._ensure_ctx()
        n = t_array.shape[0]
        t = _dev_f32(t_array, dev)
        row = batch_x.to(dev).to(torch.int64).contiguous()
        col = batch_y.to(dev).to(torch.int64).contiguous()
        pose = _dev_f32(trans_mat, dev)
        kinv = _dev_f32(K_inv, dev)
        rays = torch.empty(n, 16, device=dev)
        denc = torch.empty(n, 24, device=dev)
        stride = 17 if pose.shape[-1] == 17 else pose[0].numel()
        _lib.check(self._lib.nt_raygen(self._ctx, n, _ptr(row), _ptr(col), _ptr(pose), stride, _ptr(kinv), _ptr(rays), None,
                                       _ptr(denc), _stream(dev)))
        flat = self.network.flat_params()
        t_grad = torch.is_tensor(t_array) and t_array.requires_grad
        if torch.is_grad_enabled() and (t_grad or any(q.requires_grad for q in self.network.parameters())):
            t_in = t_array.to(dev, torch.float32).contiguous() if t_grad else t
            return _NetOutFn.apply(self, _FlatView.apply(flat, *self.network.parameters()), t_in, rays, denc)
        rgb = torch.empty(n, num_points, 3, device=dev)
        sigma = torch.empty(n, num_points, 1, device=dev)
        packed = self._pack(flat, self._prec)
        need = self._lib.nt_mlp_workspace_bytes(self._ctx, self._prec, n, num_points, 0)
        ws = torch.empty(max(need, 256), dtype=torch.uint8, device=dev)
        _lib.check(self._lib.nt_mlp_forward(self._ctx, self._prec, n, num_points, _ptr(t), _ptr(rays), _ptr(denc), _ptr(flat),
                                            _ptr(packed), _ptr(rgb), _ptr(sigma), _ptr(ws), ws.numel(), 0, _stream(dev)))
        return rgb, sigma

    def get_density(self, delta, sigma):
        """nerf.py:263-272: w = exp(-cumsum(sigma*delta)) * (1 - exp(-sigma*delta)), [N,P] (P a multiple of 32, <= 256)."""
        dev = self._ensure_ctx()
        d, s_ = _dev_f32(delta, dev), _dev_f32(sigma, dev)
        if s_.dim() == 3:
            s_ = s_.squeeze(-1).contiguous()
        if torch.is_grad_enabled() and (d.requires_grad or s_.requires_grad):
            return _DensityFn.apply(self, d, s_)
        n, p = d.shape
        w = torch.empty(n, p, dtype=torch.float32, device=dev)
        _lib.check(self._lib.nt_get_density(self._ctx, n, p, _ptr(d), _ptr(s_), _ptr(w), _stream(dev)))
        return w

    def color_cum(self, density, color):
        """nerf.py:274-281: C = sum_i w_i * rgb_i, [N,3]."""
        dev = self._ensure_ctx()
        w, c = _dev_f32(density, dev), _dev_f32(color, dev)
        if torch.is_grad_enabled() and (w.requires_grad or c.requires_grad):
            return _ColorCumFn.apply(self, w, c)
        n, p = w.shape
        out = torch.empty(n, 3, dtype=torch.float32, device=dev)
        _lib.check(self._lib.nt_color_cum(self._ctx, n, p, _ptr(w), _ptr(c), _ptr(out), _stream(dev)))
        return out

    def resample(self, t_coarse, dense_coarse):
        """nerf.py:225-261 -> t_fine [N,Nf]; raises ResampleRangeError where the reference exits."""
        dev = self._ensure_ctx()
        t = _dev_f32(t_coarse, dev)
        w = _dev_f32(dense_coarse, dev)
        if torch.is_grad_enabled() and w.requires_grad:
            out = _ResampleFn.apply(self, t, w)
        else:
            out = torch.empty(t.shape[0], self.num_fine, device=dev)
            _lib.check(self._lib.nt_sample_pdf(self._ctx, t.shape[0], _ptr(t), _ptr(w), None, _ptr(out), None, _stream(dev)))
        self.check_status()
        return out


class _FlatView(torch.autograd.Function):
    """Identity on the flat buffer whose backward scatters the flat gradient to the 24 parameters."""

    @staticmethod
    def forward(ctx, flat, *params):
        ctx.shapes = [p.shape for p in params]
        return flat.detach()

    @staticmethod
    def backward(ctx, g):
        outs, off = [], 0
        for s in ctx.shapes:
            n = math.prod(s)
            outs.append(g[off:off + n].view(s))
            off += n
        return (None, *outs)


class _RayLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, cc, cf, ct):
        n = cc.shape[0]
        loss = torch.empty(1, dtype=torch.float32, device=cc.device)
        g_cc = torch.empty_like(cc)
        g_cf = torch.empty_like(cf)
        cc, cf, ct = cc.contiguous(), cf.contiguous(), ct.contiguous()
        _lib.check(model._lib.nt_ray_loss(model._ctx, n, _ptr(cc), _ptr(cf), _ptr(ct), _ptr(loss), _ptr(g_cc), _ptr(g_cf),
                                          _stream(cc.device)))
        ctx.save_for_backward(g_cc, g_cf)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        g_cc, g_cf = ctx.saved_tensors
        return None, g_cc * g, g_cf * g, None


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam as configured at nerf.py:425, as ONE kernel over the flat buffers (nt_adam_step)."""

    def __init__(self, model: NeRFModel, lr=1e-3, betas=(0.9, 0.999), eps=1e-7, initial_lr=None):
        group = {"params": list(model.network.parameters()), "lr": lr, "betas": betas, "eps": eps}
        group["initial_lr"] = lr if initial_lr is None else initial_lr
        super().__init__([group], dict(lr=lr, betas=betas, eps=eps))
        self.model = model
        self.step_count = 0
        self.m = None
        self.v = None
        self.peer = None          # dist.PeerGradExchange: fused all-reduce + Adam over NVLink peer memory

    def enable_peer_allreduce(self, group=None):
        """Multi-GPU: keep the flat gradient in symmetric memory and fuse the SUM all-reduce into the Adam kernel
        (nt_adam_step_allreduce).  Returns False (NCCL path stays) if symmetric memory is unavailable."""
        from . import dist as D
        self.model._ensure_ctx()
        net = self.model.network
        peer = D.PeerGradExchange.create(_lib.N_PARAMS, net.flat_params().device, group)
        if peer is None:
            return False
        net.use_grad_buffer(peer.buf)
        self.peer = peer
        return True

    def zero_grad(self, set_to_none=False):
        net = self.model.network
        net.flat_grads().zero_()

    @torch.no_grad()
    def step(self, closure=None):
        model = self.model
        model._ensure_ctx()
        net = model.network
        flat, grad = net.flat_params(), net.flat_grads()
        if self.m is None:
            self.m = torch.zeros_like(flat)
            self.v = torch.zeros_like(flat)
        elif self.m.device != flat.device:                     # moments restored from a checkpoint
            self.m, self.v = self.m.to(flat.device).contiguous(), self.v.to(flat.device).contiguous()
        self.step_count += 1
        g = self.param_groups[0]
        if self.peer is not None:
            # every rank's gradient is complete -> one kernel sums all ranks' buffers over NVLink and updates -> nobody
            # rewrites its gradient before all ranks have read it
            peer = self.peer
            arr = (C.c_void_p * peer.world)(*peer.ptrs)
            peer.barrier()
            _lib.check(model._lib.nt_adam_step_allreduce(model._ctx, flat.numel(), _ptr(flat), arr, peer.world, _ptr(self.m),
                                                         _ptr(self.v), float(g["lr"]), g["betas"][0], g["betas"][1], g["eps"],
                                                         self.step_count, 1.0, None, _stream(flat.device)))
            peer.barrier()
            return
        _lib.check(model._lib.nt_adam_step(model._ctx, flat.numel(), _ptr(flat), _ptr(grad), _ptr(self.m), _ptr(self.v),
                                           float(g["lr"]), g["betas"][0], g["betas"][1], g["eps"], self.step_count, 1.0,
                                           _stream(flat.device)))

    def state_dict(self):
        sd = super().state_dict()
        sd["fused"] = {"step": self.step_count, "m": self.m, "v": self.v}
        return sd

    def load_state_dict(self, sd):
        fused = sd.pop("fused", None)
        super().load_state_dict(sd)
        if fused is not None:
            self.step_count, self.m, self.v = fused["step"], fused["m"], fused["v"]


def train_step(model: NeRFModel, optimizer: FusedAdam, row, column, pix_val, poses_bound, K_inv,
               grad_allreduce=None, shard=None):
    """One iteration of the reference loop body nerf.py:464-475 without autograd bookkeeping:
    forward(train) -> ray_loss -> backward -> [all-reduce] -> fused Adam.  Returns (loss, C_coarse, C_fine)
    as device tensors (no host sync).  `shard`: see NeRFModel (this rank's slice of a ray-sharded batch)."""
    dev = model._ensure_ctx()
    L = model._lib
    n = row.shape[0]
    pb = poses_bound.to(dev, non_blocking=True).to(torch.float32).contiguous()
    rowd = row.to(dev, non_blocking=True).to(torch.int64).contiguous()
    cold = column.to(dev, non_blocking=True).to(torch.int64).contiguous()
    ct = pix_val.to(dev, non_blocking=True).to(torch.float32).contiguous()
    kinv = _dev_f32(K_inv, dev)
    near, far = pb[:, 15].contiguous(), pb[:, 16].contiguous()
    net = model.network
    flat, grads = net.flat_params(), net.flat_grads()
    asz, d0 = model._globals(shard, near, far)
    grads.zero_()                                                     # optimizer.zero_grad() (nerf.py:467)
    cc, cf, ws = model._render_raw(flat, rowd, cold, pb, kinv, near, far, train=True, any_step_zero=asz, delta0=d0)  # :470
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    g_cc, g_cf = torch.empty_like(cc), torch.empty_like(cf)
    prec = model._prec_train
    _lib.check(L.nt_ray_loss(model._ctx, n, _ptr(cc), _ptr(cf), _ptr(ct), _ptr(loss), _ptr(g_cc), _ptr(g_cf), _stream(dev)))
    _lib.check(L.nt_render_backward(model._ctx, prec, n, _ptr(near), _ptr(far), _ptr(flat), _ptr(model._packed.get(prec)),
                                    _ptr(d0), _ptr(g_cc), _ptr(g_cf), _ptr(grads), _ptr(ws), ws.numel(), _stream(dev)))
    if grad_allreduce is not None and getattr(optimizer, "peer", None) is None:
        grad_allreduce(grads)                                         # NCCL SUM all-reduce of the flat gradient
    optimizer.step()                                                  # nerf.py:474 (fused with the all-reduce if peer)
    return loss, cc, cf


class GraphedTrainStep:
    """`train_step` for a fixed batch size with the forward / loss / backward launches (~45 kernels, most of them a few
    microseconds) captured ONCE into a CUDA graph and replayed: the inputs are copied into static device buffers, the
    fused Adam - whose step count and learning rate change every iteration - and the multi-GPU exchanges (the 16-byte
    shard-globals reduction before the replay, the gradient exchange after it) stay ordinary launches.  Same arithmetic,
    same kernels, same order as `train_step`.  The graph is re-captured if the parameter / gradient buffers it was
    recorded on have been replaced since (Network._flatten, FusedAdam.enable_peer_allreduce)."""

    def __init__(self, model: NeRFModel, optimizer: FusedAdam, n_rays: int, K_inv, shard=None):
        dev = model._ensure_ctx()
        self.model, self.opt, self.n, self.shard = model, optimizer, int(n_rays), shard
        self.row = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self.col = torch.zeros(self.n, dtype=torch.int64, device=dev)
        self.pix = torch.zeros(self.n, 3, dtype=torch.float32, device=dev)
        self.pb = torch.zeros(self.n, 17, dtype=torch.float32, device=dev)
        self._nf = torch.zeros(2, self.n, dtype=torch.float32, device=dev)   # near / far rows, filled by ONE strided copy
        self.near, self.far = self._nf[0], self._nf[1]
        self.kinv = _dev_f32(K_inv, dev)
        # batch-global quantities of a sharded batch live in a static buffer the captured kernels read
        self.g = torch.zeros(4, dtype=torch.float32, device=dev) if shard is not None else None
        self.graph = None
        self.out = None
        self._key = None

    def _buffers_key(self):
        net = self.model.network
        return (net.flat_params().data_ptr(), net.flat_grads().data_ptr())

    def _body(self):
        model, L = self.model, self.model._lib
        net = model.network
        dev = self.row.device
        flat, grads = net.flat_params(), net.flat_grads()
        asz = _lib.ANY_STEP_ZERO_DEVICE if self.g is not None else -1
        grads.zero_()
        cc, cf, ws = model._render_raw(flat, self.row, self.col, self.pb, self.kinv, self.near, self.far, train=True,
                                       any_step_zero=asz, delta0=self.g)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        g_cc, g_cf = torch.empty_like(cc), torch.empty_like(cf)
        prec = model._prec_train
        _lib.check(L.nt_ray_loss(model._ctx, self.n, _ptr(cc), _ptr(cf), _ptr(self.pix), _ptr(loss), _ptr(g_cc), _ptr(g_cf),
                                 _stream(dev)))
        _lib.check(L.nt_render_backward(model._ctx, prec, self.n, _ptr(self.near), _ptr(self.far), _ptr(flat),
                                        _ptr(model._packed.get(prec)), _ptr(self.g), _ptr(g_cc), _ptr(g_cf), _ptr(grads),
                                        _ptr(ws), ws.numel(), _stream(dev)))
        return loss, cc, cf, (ws, g_cc, g_cf)

    def _capture(self):
        dev = self.row.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                 # eager warm-up: one-time attribute / table setup inside the library
            self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._body()
        self._key = self._buffers_key()

    def __call__(self, row, column, pix_val, poses_bound, grad_allreduce=None):
        if row.shape[0] != self.n:
            raise _lib.NerfTinyError("GraphedTrainStep was built for %d rays, got %d" % (self.n, row.shape[0]))
        self.row.copy_(row, non_blocking=True)
        self.col.copy_(column, non_blocking=True)
        self.pix.copy_(pix_val, non_blocking=True)
        self.pb.copy_(poses_bound, non_blocking=True)      # float64 loader rows are converted by the copy (nerf.py:338)
        self._nf.copy_(self.pb[:, 15:17].t())
        if self.g is not None:
            _, g = self.model._globals(self.shard, self.near, self.far)
            self.g.copy_(g)
        if self.graph is None or self._key != self._buffers_key():
            self._capture()
        self.graph.replay()
        if grad_allreduce is not None and getattr(self.opt, "peer", None) is None:
            grad_allreduce(self.model.network.flat_grads())
        self.opt.step()
        return self.out[0], self.out[1], self.out[2]


# ----------------------------------------------------------------------------------------------------
# Checkpoint files (nerf.py:491 / :410): whole-module pickles that both this module and the reference can read.
# ----------------------------------------------------------------------------------------------------
_REF_CLASSES = ("NeRFModel", "Network", "Encoder", "Activation")


class _RefPickler(_pickle._Pickler):
    """Records this module's classes under the reference's module name (`nerf`)."""

    def save_global(self, obj, name=None):
        if isinstance(obj, type) and obj.__module__ == __name__ and obj.__name__ in _REF_CLASSES:
            self.write(b"c" + b"nerf\n" + obj.__name__.encode() + b"\n")     # GLOBAL opcode: module `nerf`
            self.memoize(obj)
            return
        return super().save_global(obj, name)

    dispatch = dict(_pickle._Pickler.dispatch)
    dispatch[type] = save_global


class _RefUnpickler(_pickle._Unpickler):
    """Resolves `nerf.*` / `__main__.*` class references of reference-written files to the classes here."""

    def find_class(self, mod_name, name):
        if mod_name in ("nerf", "__main__", "ref_nerf") and name in _REF_CLASSES:   # ref_nerf: the name the test harness imports the reference under
            return globals()[name]
        return super().find_class(mod_name, name)


class _RefPickle:
    """pickle_module for torch.save / torch.load."""
    __name__ = "nerf_tiny_b200.nerf._RefPickle"
    HIGHEST_PROTOCOL, DEFAULT_PROTOCOL = _pickle.HIGHEST_PROTOCOL, _pickle.DEFAULT_PROTOCOL
    PickleError, PicklingError, UnpicklingError = _pickle.PickleError, _pickle.PicklingError, _pickle.UnpicklingError
    load, loads, dump, dumps = _pickle.load, _pickle.loads, _pickle.dump, _pickle.dumps
    Pickler, Unpickler = _RefPickler, _RefUnpickler


def save_checkpoint(model, path):
    torch.save(model, path, pickle_module=_RefPickle)


def load_checkpoint(path):
    return torch.load(path, map_location="cpu", weights_only=False, pickle_module=_RefPickle)


# ----------------------------------------------------------------------------------------------------
# NeRFRunner (nerf.py:353-530): constructor signature, trainer(mode) / display() call surface, resume-from-checkpoint
# and periodic save.  Batches come from loader.GpuRayBatches (pixels and pose rows
# resident in HBM) when an image folder is given, or from any iterable of loader-shaped tuples.
# ----------------------------------------------------------------------------------------------------
class NeRFRunner():
    DISPLAY_RAYS = 1 << 16      # rays per render launch of display()

    def __init__(self, gpu=0, img_dir=None, results_path="./results/", ckpt_path="./checkpoint/", low_res=1,
                 total_iter=100000, batch_ray=400, learning=1e-3, lr_gamma=0.1, lr_milestone=[10, 200], n_coarse=64,
                 n_fine=128, data_type="llff", step=100, decay_end=200000, sched="EXP", continue_=False,
                 train_batches=None, disp_batches=None, height=None, width=None, focal=None, num_pic=1,
                 precision=None):
        global device, writer
        import glob
        if not torch.cuda.is_available():
            raise _lib.NerfTinyError("NeRFRunner needs a CUDA device: the B200 path has no CPU fallback")
        device = torch.device("cuda:" + str(gpu))
        print("Using device", device)
        self.start_time = time.strftime("%m-%d-%H-%M-%S", time.localtime())
        self.model = NeRFModel(num_coarse=n_coarse, num_fine=n_fine, batch_ray=batch_ray, precision=precision).to(device)
        self.results_path, self.ckpt_path, self.low_res = results_path, ckpt_path, low_res
        self.total_iter, self.batch_ray, self.step, self.decay_end = total_iter, batch_ray, step, decay_end
        # resume: newest "<ckpt_path>*_<iter>.pkl" (nerf.py:404-415).  Accepted: whole-module pickles written by this runner
        # OR by the reference itself (torch.save(self.model), nerf.py:491 - its `nerf.*` classes are mapped onto ours), and
        # the {"model": state_dict, ...} dictionaries an earlier version of this runner wrote.
        last_iter, opt_state = -1, None
        ck_list = glob.glob(ckpt_path + "*.pkl")
        if continue_ is True and ck_list:
            best = max(ck_list, key=lambda f: int(f.split("_")[-1][:-4]))
            last_iter = int(best.split("_")[-1][:-4])
            print("Last iter:", last_iter)
            blob = load_checkpoint(best)
            if isinstance(blob, dict) and "model" in blob:
                self.model.load_state_dict(blob["model"])
                opt_state = blob.get("optimizer")
            else:
                self.model.load_state_dict(blob.state_dict())
            if opt_state is None and os.path.exists(best[:-4] + ".opt"):
                opt_state = torch.load(best[:-4] + ".opt", map_location="cpu", weights_only=False)
            self.model = self.model.to(device)
        else:
            print("New running created.")
        self.last_iter = last_iter
        if train_batches is None and img_dir is not None:
            from . import loader
            ds = loader.NeRFDataset(root_dir=img_dir, low_res=low_res, transform=None, type=data_type, mode="train")
            self.train_dataset = ds
            train_batches = loader.GpuRayBatches.from_dataset(ds, batch_ray, shuffle=True, device=device)
            height, width, focal, num_pic = ds.height, ds.width, ds.focal, ds.pic_num
            try:
                self.val_dataset = loader.NeRFDataset(root_dir=img_dir, low_res=low_res, type=data_type, mode="val")
                self.disp_dataset = loader.NeRFDataset(root_dir=img_dir, low_res=low_res, type=data_type, mode="test")
                self.val_dataloader = loader.GpuRayBatches.from_dataset(self.val_dataset, batch_ray, shuffle=True, device=device)
                # display: every pixel in flat order, tiled into launches large enough to fill the GPU (the reference's
                # batch_ray-sized, drop_last display loader leaves the tail of the last view white, nerf.py:442)
                disp_batches = loader.GpuRayBatches.from_dataset(self.disp_dataset, max(batch_ray, self.DISPLAY_RAYS),
                                                                 shuffle=False, drop_last=False, device=device)
                num_pic = self.disp_dataset.pic_num
            except (FileNotFoundError, OSError):
                self.val_dataloader, disp_batches = train_batches, None
        self.train_dataloader = train_batches
        if not hasattr(self, "val_dataloader"):
            self.val_dataloader = train_batches
        self.disp_dataloader = disp_batches
        self.optimizer = FusedAdam(self.model, lr=learning, betas=(0.9, 0.999), eps=1e-7, initial_lr=learning)
        if opt_state is not None:
            self.optimizer.load_state_dict(opt_state)
        lam = (lambda it: lr_gamma ** (it / decay_end) if it < decay_end else lr_gamma * learning)   # nerf.py:426
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lr_lambda=lam, last_epoch=self.last_iter) \
            if sched == "EXP" else torch.optim.lr_scheduler.MultiStepLR(self.optimizer, lr_milestone, lr_gamma,
                                                                         last_epoch=self.last_iter)
        self.height, self.width, self.focal, self.num_pic = height, width, focal, num_pic
        if height is not None:
            self.K_inv = torch.tensor([[1.0, 0.0, -0.5 * width], [0.0, -1.0, 0.5 * height], [0.0, 0.0, -focal]]
                                      ).to(torch.float).transpose(0, 1)                               # nerf.py:433
        self.losses = []

    def save_checkpoint(self, it):
        """nerf.py:491: torch.save(self.model) under '<ckpt_path><start_time>_<iter>.pkl' - the SAME on-disk format, a
        whole-module pickle whose classes are recorded as `nerf.NeRFModel / Network / Encoder / Activation`, so the
        reference's own resume (`torch.load(last_ckpt).to(device)`, nerf.py:410) reads it.  The Adam moments the reference
        forgot go into a sidecar '<...>_<iter>.opt' that its '*.pkl' glob ignores."""
        os.makedirs(os.path.dirname(self.ckpt_path) or ".", exist_ok=True)
        path = self.ckpt_path + self.start_time + "_" + str(it) + ".pkl"
        save_checkpoint(self.model, path)
        torch.save(self.optimizer.state_dict(), path[:-4] + ".opt")
        return path

    def _flush_losses(self, it):
        """The reference logs a TensorBoard scalar per step (a host sync each, nerf.py:478); here the per-step losses stay
        on the device and are appended to <results_path><start_time>_loss.csv once every `step` iterations (and to
        nerf.writer, if the caller installed a SummaryWriter-like object there)."""
        if not self.losses:
            return
        vals = torch.cat([l.reshape(1) for l in self.losses]).cpu().tolist()
        first = it - len(vals) + 1
        os.makedirs(os.path.dirname(self.results_path) or ".", exist_ok=True)
        with open(self.results_path + self.start_time + "_loss.csv", "a") as f:
            for k, v in enumerate(vals):
                f.write("%d,%.6f\n" % (first + k, v))
                if writer is not None:
                    writer.add_scalar("Loss/train", v, first + k)
        self.last_loss = vals[-1]
        self.losses = []

    def trainer(self, mode):
        """nerf.py:445-499 loop body (forward, ray_loss, backward, Adam, scheduler) over the batch source; the loss is
        kept on the device (no per-step host sync), a checkpoint is written every `step` iterations."""
        dataloader = getattr(self, mode + "_dataloader")
        it = self.last_iter + 1
        while it < self.total_iter:
            n_seen = 0
            for (row, column, pix_val, poses_bound, pic) in dataloader:
                if row.shape[0] == self.batch_ray and self.model._prec_train == _lib.PREC_BF16:   # fixed-size batches: graph replay
                    if getattr(self, "_graphed", None) is None:
                        self._graphed = GraphedTrainStep(self.model, self.optimizer, self.batch_ray, self.K_inv)
                    loss, _, _ = self._graphed(row, column, pix_val, poses_bound)
                    loss = loss.clone()                 # the graph's output buffer is overwritten by the next replay
                else:
                    loss, _, _ = train_step(self.model, self.optimizer, row, column, pix_val, poses_bound, self.K_inv)
                self.scheduler.step()
                self.losses.append(loss)
                if ((it + 1) % self.step) == 0:
                    print("\n[ITER]", it, " [LOSS] %.4f" % float(loss))
                    self._flush_losses(it)
                    self.save_checkpoint(it)
                it += 1
                n_seen += 1
                if it >= self.total_iter:
                    break
            if mode == "val" or n_seen == 0:
                break
        self.last_iter = it - 1
        self._flush_losses(self.last_iter)
        self.model.check_status()

    def display(self, save=False):
        """nerf.py:503-530: no-grad render of every test batch, scattered into (num_pic, H, W, 3); with save=True the
        frames are written as <results_path><start_time>/<i>.jpg plus an animated video.gif (the reference writes video.mp4
        through imageio, nerf.py:528; PIL is what this image has), uint8 conversion on the GPU."""
        result = torch.full((self.num_pic, self.height, self.width, 3), 1.0, device=device)
        with torch.no_grad():
            self.model.eval()
            for (row, column, pix_val, poses_bound, pic) in self.disp_dataloader:
                _, c_fine = self.model(row, column, poses_bound, self.K_inv)
                result[pic.to(device), row.to(device), column.to(device)] = c_fine
        if save:
            from PIL import Image
            out_dir = self.results_path + self.start_time + "/"
            os.makedirs(out_dir, exist_ok=True)
            frames = (result.clamp(0, 1) * 255.0).to(torch.uint8).cpu().numpy()
            images = [Image.fromarray(frames[i]) for i in range(self.num_pic)]
            for i, im in enumerate(images):
                im.save(out_dir + str(i) + ".jpg")
            if len(images) > 1:
                images[0].save(out_dir + "video.gif", save_all=True, append_images=images[1:], duration=33, loop=0)
        return result
