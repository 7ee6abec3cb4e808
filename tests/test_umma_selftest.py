"""Pins the tcgen05 descriptor conventions of csrc/rl_umma.cuh against torch matmul (fp32 reference of
bf16 inputs).  Every tensor-core kernel in the library relies on exactly these operand layouts."""
import ctypes

import pytest
import torch

from sim2real_lane_segment_b200 import _lib


def _run(a, a_mn, a_shift, b, b_mn, b_shift, N, K, bulk_b=False, b_packed=None):
    L = _lib.lib()
    D = torch.full((128, N), float("nan"), device="cuda", dtype=torch.float32)
    bsrc = b_packed if bulk_b else b
    st = L.rl_selftest_umma(_lib.ptr(a), a.shape[0], a.shape[1], a_mn, a_shift,
                            _lib.ptr(bsrc), b.shape[0], b.shape[1], b_mn, b_shift,
                            N, K, _lib.ptr(D), int(bulk_b), _lib.cur_stream())
    _lib.check(st)
    torch.cuda.synchronize()
    return D


def _chunk_pack(m):
    """row-major [R, C] -> chunk layout [C/8][R][8] flattened"""
    R, C = m.shape
    return m.reshape(R, C // 8, 8).permute(1, 0, 2).contiguous()


@pytest.mark.gpu
@pytest.mark.parametrize("N", [16, 48, 144, 256])
@pytest.mark.parametrize("K", [16, 64, 144])
def test_kmajor(N, K):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    D = _run(a, 0, 0, b, 0, 0, N, K)
    ref = a.float() @ b.float().t()
    torch.testing.assert_close(D, ref, rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
def test_kmajor_bulk_b():
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn(128, 64, device="cuda", generator=g).bfloat16()
    b = torch.randn(144, 64, device="cuda", generator=g).bfloat16()
    D = _run(a, 0, 0, b, 0, 0, 144, 64, bulk_b=True, b_packed=_chunk_pack(b))
    torch.testing.assert_close(D, a.float() @ b.float().t(), rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("shift_a,shift_b", [(0, 0), (1, 0), (0, 3), (21, 162), (163, 1)])
def test_kmajor_row_shift(shift_a, shift_b):
    g = torch.Generator(device="cuda").manual_seed(3)
    K, N = 32, 128
    a = torch.randn(128 + 170, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N + 170, K, device="cuda", generator=g).bfloat16()
    D = _run(a, 0, shift_a, b, 0, shift_b, N, K)
    ref = a[shift_a:shift_a + 128].float() @ b[shift_b:shift_b + N].float().t()
    torch.testing.assert_close(D, ref, rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [16, 144])
@pytest.mark.parametrize("shift", [0, 5])
def test_mnmajor(N, shift):
    """A given as [K rows, 128 cols], B as [K rows, N cols]: D = A^T B (the wgrad form)."""
    g = torch.Generator(device="cuda").manual_seed(4)
    K = 64
    a = torch.randn(K + 8, 128, device="cuda", generator=g).bfloat16()
    b = torch.randn(K + 8, N, device="cuda", generator=g).bfloat16()
    D = _run(a, 1, shift, b, 1, 0, N, K)
    ref = a[shift:shift + K].float().t() @ b[:K].float()
    torch.testing.assert_close(D, ref, rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
def test_mixed_major():
    """A K-major [128, K], B MN-major [K, N]."""
    g = torch.Generator(device="cuda").manual_seed(5)
    K, N = 48, 64
    a = torch.randn(128, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(K, N, device="cuda", generator=g).bfloat16()
    D = _run(a, 0, 0, b, 1, 0, N, K)
    torch.testing.assert_close(D, a.float() @ b.float(), rtol=1e-4, atol=1e-3)
