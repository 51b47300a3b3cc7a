"""Unit tests of the bf16 tcgen05 GEMM (gemm_tc.cu) against torch matmul on the same bf16 operands."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from nerf_tiny_b200 import build, ops
    build.build()
    c = ops.Context(0)
    yield c
    c.close()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _run(ctx, mn_major, a, b, m, n, k, out_f32, mask=None, ldc=None):
    from nerf_tiny_b200 import _lib
    ldc = ldc or n
    c = torch.zeros(m, ldc, device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(ctx.lib.nt_gemm_bf16_debug(ctx.h, mn_major, m, n, k, _p(a), a.shape[1], _p(b), b.shape[1], _p(c), ldc,
                                          1 if out_f32 else 0, _p(mask), mask.shape[1] if mask is not None else 0, st))
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("m,n,k", [(128, 256, 256), (384, 256, 256), (1000, 128, 128), (128 * 200, 256, 256),
                                   (128 * 150 + 64, 64, 256), (4096, 60, 256)])
def test_k_major(ctx, m, n, k):
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn(m, k, device="cuda", generator=g).bfloat16()
    b = (torch.randn(n, k, device="cuda", generator=g) * 0.1).bfloat16()
    ref = a.float() @ b.float().t()
    c = _run(ctx, 0, a, b, m, n, k, out_f32=False)
    err = (c.float() - ref).abs().max() / ref.abs().max()
    assert err < 1e-2, err                       # bf16 output rounding
    mask = torch.randn(m, n, device="cuda", generator=g).bfloat16()
    cm = _run(ctx, 0, a, b, m, n, k, out_f32=False, mask=mask)
    refm = torch.where(mask > 0, ref, torch.zeros_like(ref))
    assert (cm.float() - refm).abs().max() / ref.abs().max() < 1e-2


@pytest.mark.parametrize("m,n,k", [(256, 256, 64 * 40), (128, 256, 64 * 300), (256, 60, 64 * 77), (8, 256, 64 * 50),
                                   (128, 24, 64 * 33), (256, 256, 196608)])
def test_mn_major_split_k(ctx, m, n, k):
    g = torch.Generator(device="cuda").manual_seed(m * 3 + n + k)
    a = (torch.randn(k, m, device="cuda", generator=g) * 0.05).bfloat16()      # stored [K][M]
    b = torch.randn(k, n if n % 8 == 0 else (n + 7) // 8 * 8, device="cuda", generator=g).bfloat16()    # stored [K][N] (row pitch % 8)
    ref = a.float().t() @ b.float()[:, :n]
    ldc = 316 if n == 60 else n
    c = _run(ctx, 1, a, b, m, n, k, out_f32=True, ldc=ldc)
    err = (c[:, :n] - ref).abs().max() / ref.abs().max()
    assert err < 2e-3, err                       # fp32 accumulation, split-K atomics
    if ldc > n:
        assert float(c[:, n:].abs().max()) == 0.0    # nothing written past n_valid
