"""Thin tensor-in / tensor-out wrappers over the C-ABI, one per kernel family.

Used by the parity tests (which call every kernel through the same ABI a non-Python host would bind) and by
callers that want the reference's piecewise methods.  torch supplies device memory and the stream only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _st(dev=None):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class Context:
    def __init__(self, device=0, n_coarse=64, n_fine=128):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.nt_create(C.byref(h), device, n_coarse, n_fine))
        self.h = h
        self.dev = torch.device("cuda", device)
        self.nc, self.nf = n_coarse, n_fine

    def close(self):
        if self.h is not None:
            self.lib.nt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _f(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.dev)

    def set_option(self, key, value):
        _lib.check(self.lib.nt_set_option(self.h, key, value))

    @property
    def launches(self):
        return int(self.lib.nt_launch_count(self.h))

    # ---- geometry -------------------------------------------------------------------------------
    def raygen(self, row, col, c2w, kinv):
        n = row.shape[0]
        stride = 17 if c2w.shape[-1] == 17 else c2w[0].numel()
        rays, dw, de = self._f(n, 16), self._f(n, 3), self._f(n, 24)
        _lib.check(self.lib.nt_raygen(self.h, n, _p(row), _p(col), _p(c2w), stride, _p(kinv), _p(rays), _p(dw), _p(de), _st()))
        return rays, dw, de

    def sample_coarse(self, near, far, any_step_zero=-1):
        n = near.shape[0]
        t = self._f(n, self.nc)
        _lib.check(self.lib.nt_sample_coarse(self.h, n, _p(near), _p(far), any_step_zero, _p(t), _st()))
        return t

    def shard_globals_local(self, near, far, first_shard):
        """[flag, delta0 candidate if flag == 0, candidate if flag == 1, 0] of this shard."""
        g = self._f(4)
        n = near.shape[0]
        _lib.check(self.lib.nt_shard_globals_local(self.h, n, _p(near) if n else C.c_void_p(0), _p(far) if n else C.c_void_p(0),
                                                   1 if first_shard else 0, _p(g), _st()))
        return g

    def shard_globals_resolve(self, g):
        _lib.check(self.lib.nt_shard_globals_resolve(self.h, _p(g), _st()))
        return g

    # ---- MLP ----------------------------------------------------------------------------------------
    def pack(self, flat, precision):
        nbytes = max(256, self.lib.nt_packed_weight_bytes(self.h, precision))
        packed = torch.empty(nbytes, dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.nt_pack_weights(self.h, precision, _p(flat), _p(packed), _st()))
        return packed

    def mlp_forward(self, precision, t, rays, dir_enc, flat, packed=None, train=False):
        n, p = t.shape
        rgb, sigma = self._f(n, p, 3), self._f(n, p)
        need = max(256, self.lib.nt_mlp_workspace_bytes(self.h, precision, n, p, 1 if train else 0))
        ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        _lib.check(self.lib.nt_mlp_forward(self.h, precision, n, p, _p(t), _p(rays), _p(dir_enc), _p(flat), _p(packed),
                                           _p(rgb), _p(sigma), _p(ws), ws.numel(), 1 if train else 0, _st()))
        return rgb, sigma, ws

    def mlp_forward_debug(self, t, rays, dir_enc, flat, packed, layer):
        n, p = t.shape
        rgb, sigma = self._f(n, p, 3), self._f(n, p)
        dbg = torch.zeros(n * p, 256, dtype=torch.float32, device=self.dev)
        _lib.check(self.lib.nt_mlp_forward_debug(self.h, n, p, _p(t), _p(rays), _p(dir_enc), _p(flat), _p(packed), _p(rgb),
                                                 _p(sigma), _p(dbg), layer, _st()))
        return rgb, sigma, dbg

    def mlp_backward(self, precision, t, rays, dir_enc, flat, packed, rgb, g_rgb, g_sigma, ws, need_gt=True):
        n, p = t.shape
        grads = torch.zeros_like(flat)
        g_t = self._f(n, p) if need_gt else None
        _lib.check(self.lib.nt_mlp_backward(self.h, precision, n, p, _p(t), _p(rays), _p(dir_enc), _p(flat), _p(packed),
                                            _p(rgb), _p(g_rgb), _p(g_sigma), _p(grads), _p(g_t), _p(ws), ws.numel(), _st()))
        return grads, g_t

    # ---- compositing / resampling ---------------------------------------------------------------------
    def composite_coarse(self, near, far, rgb, sigma):
        n = near.shape[0]
        w, c = self._f(n, self.nc), self._f(n, 3)
        _lib.check(self.lib.nt_composite_coarse(self.h, n, _p(near), _p(far), _p(rgb), _p(sigma), _p(w), _p(c), _st()))
        return w, c

    def composite_coarse_backward(self, near, far, rgb, sigma, g_c, g_w_ext=None):
        n = near.shape[0]
        g_rgb, g_sigma = self._f(n, self.nc, 3), self._f(n, self.nc)
        _lib.check(self.lib.nt_composite_coarse_backward(self.h, n, _p(near), _p(far), _p(rgb), _p(sigma), _p(g_c),
                                                         _p(g_w_ext), _p(g_rgb), _p(g_sigma), _st()))
        return g_rgb, g_sigma

    def composite_fine(self, t_c, rgb_c, sig_c, t_f, rgb_f, sig_f, last=1e-4, want_perm=True):
        n = t_c.shape[0]
        tot = self.nc + self.nf
        c, w = self._f(n, 3), self._f(n, tot)
        perm = torch.empty(n, 5, tot, dtype=torch.uint8, device=self.dev) if want_perm else None
        _lib.check(self.lib.nt_composite_fine(self.h, n, _p(t_c), _p(rgb_c), _p(sig_c), _p(t_f), _p(rgb_f), _p(sig_f), last,
                                              _p(c), _p(w), _p(perm), _st()))
        return c, w, perm

    def composite_fine_backward(self, t_c, rgb_c, sig_c, t_f, rgb_f, sig_f, perm, g_c, last=1e-4):
        n = t_c.shape[0]
        g_rgb_c, g_sig_c = self._f(n, self.nc, 3), self._f(n, self.nc)
        g_rgb_f, g_sig_f, g_t_f = self._f(n, self.nf, 3), self._f(n, self.nf), self._f(n, self.nf)
        _lib.check(self.lib.nt_composite_fine_backward(self.h, n, _p(t_c), _p(rgb_c), _p(sig_c), _p(t_f), _p(rgb_f), _p(sig_f),
                                                       last, _p(perm), _p(g_c), _p(g_rgb_c), _p(g_sig_c), _p(g_rgb_f),
                                                       _p(g_sig_f), _p(g_t_f), _st()))
        return g_rgb_c, g_sig_c, g_rgb_f, g_sig_f, g_t_f

    def sample_pdf(self, t_coarse, w, delta0=None):
        n = t_coarse.shape[0]
        t_f = self._f(n, self.nf)
        idx = torch.empty(n, self.nf, dtype=torch.int32, device=self.dev)
        _lib.check(self.lib.nt_sample_pdf(self.h, n, _p(t_coarse), _p(w), _p(delta0), _p(t_f), _p(idx), _st()))
        return t_f, idx

    def sample_pdf_backward(self, t_coarse, w, g_tf, delta0=None):
        n = t_coarse.shape[0]
        g_w = self._f(n, self.nc)
        _lib.check(self.lib.nt_sample_pdf_backward(self.h, n, _p(t_coarse), _p(w), _p(delta0), _p(g_tf), _p(g_w), _st()))
        return g_w

    def check_status(self):
        _lib.check(self.lib.nt_check_status(self.h, _st()))

    def ray_loss(self, cc, cf, ct):
        n = cc.shape[0]
        loss, g_cc, g_cf = self._f(1), self._f(n, 3), self._f(n, 3)
        _lib.check(self.lib.nt_ray_loss(self.h, n, _p(cc), _p(cf), _p(ct), _p(loss), _p(g_cc), _p(g_cf), _st()))
        return loss, g_cc, g_cf

    def adam_step(self, p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-7, grad_scale=1.0):
        _lib.check(self.lib.nt_adam_step(self.h, p.numel(), _p(p), _p(g), _p(m), _p(v), lr, b1, b2, eps, step, grad_scale, _st()))


def flatten_state_dict(sd, device=None):
    """state_dict (reference keys, nerf.py:85-99 order) -> flat fp32 [593924] in the C-ABI layout."""
    from .nerf_keys import LAYER_KEYS
    parts = []
    for k in LAYER_KEYS:
        parts.append(sd[k + ".weight"].reshape(-1).float())
        parts.append(sd[k + ".bias"].reshape(-1).float())
    flat = torch.cat(parts)
    return flat.to(device) if device is not None else flat
