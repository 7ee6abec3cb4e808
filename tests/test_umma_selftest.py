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


@pytest.mark.gpu
def test_sw128_mnmajor_a_operand_from_tensor_map_box_layout():
    """MN-major A in the 128-byte-swizzle layout a [rows = K][64 elements = M] tensor-map box lands in: SBO = 1024 between
    8-row K groups, LBO = distance of the next 64-element M atom (the next box), K step of 16 rows = +2048 B
    (desc_sw128_mnmajor in csrc/rl_umma.cuh); B = the un-swizzled chunk layout read MN-major."""
    import ctypes
    import numpy as np
    L = _lib.lib()
    L.rl_selftest_umma_raw.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_uint, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p]

    def desc(lbo, sbo, layout):
        return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | layout << 61

    torch.manual_seed(0)
    K, M, N = 64, 128, 16
    A = torch.randn(K, M).bfloat16()
    B = torch.randn(K, N).bfloat16()
    ref = A.float().t() @ B.float()
    a16 = A.view(torch.int16).numpy()
    img = np.zeros((2, K, 64), dtype=np.int16)
    for atom in range(2):
        for p in range(K):
            for j in range(8):
                jj = j ^ (p & 7)
                img[atom, p, jj * 8:jj * 8 + 8] = a16[p, atom * 64 + j * 8: atom * 64 + j * 8 + 8]
    a_img = torch.from_numpy(img.reshape(-1).copy()).cuda()
    b_img = B.view(torch.int16).reshape(K, 2, 8).permute(1, 0, 2).contiguous().cuda()
    idesc = (1 << 4) | (1 << 7) | (1 << 10) | (1 << 15) | (1 << 16) | ((N >> 3) << 17) | ((128 >> 4) << 24)
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = L.rl_selftest_umma_raw(a_img.data_ptr(), a_img.numel() * 2, desc(K * 128, 1024, 2), 2048, b_img.data_ptr(),
                                b_img.numel() * 2, desc(128, 16 * K, 0), 256, idesc, K // 16, N, D.data_ptr(), _lib.cur_stream())
    _lib.check(rc)
    torch.cuda.synchronize()
    torch.testing.assert_close(D.cpu(), ref, rtol=1e-4, atol=1e-3)


@pytest.mark.gpu
def test_tmem_ld_16x256b_layout():
    """fragment layout of tcgen05.ld.16x256b.x4 that TransitionDown's pooling epilogue (csrc/rl_conv_fwd.cu) relies on:
    thread t of a warp receives lanes t/4 and t/4 + 8 (first load) / t/4 + 16 and t/4 + 24 (second load, lane base + 16) of
    the warp's 32-lane quarter, columns 8k + 2(t%4) + {0,1}"""
    import ctypes
    L = _lib.lib()
    out = torch.zeros(128, 2, 16, dtype=torch.int32, device="cuda")
    _lib.check(L.rl_selftest_tmem_shapes(ctypes.c_void_p(out.data_ptr()), _lib.cur_stream()))
    torch.cuda.synchronize()
    o = out.cpu()
    for t in range(128):
        w, l = t // 32, t % 32
        for half in range(2):
            for k in range(4):
                for j in range(4):
                    lane = 32 * w + 16 * half + l // 4 + (8 if j >= 2 else 0)
                    col = 8 * k + 2 * (l % 4) + (j & 1)
                    v = int(o[t, half, 4 * k + j])
                    assert (v >> 8, v & 255) == (lane, col), (t, half, k, j, v >> 8, v & 255)


@pytest.mark.gpu
def test_im2col_tensor_map_with_overlapping_dimensions():
    """the tensor map of the weight-gradient kernel's B operand (csrc/rl_bwd_fused.cu encode_dz_im2col_tmap): dimensions
    (64 positions x 8 ch as one 1 KB row, 3 x-offsets, 3 y-offsets, 2 planes) where the x-offset dimension overlaps the rows.  One
    box must deliver tile[plane][iy][ix][row][:] = planes[plane][c1 + row + iy*(W+2) + ix][:], zero where the position
    coordinate c1 + row is outside [0, positions)"""
    import ctypes
    L = _lib.lib()
    W, npos = 24, 4000
    positions = npos - 2 * (W + 2) - 2
    planes = torch.arange(2 * npos * 8, dtype=torch.int32, device="cuda").to(torch.int16).view(2, npos, 8)
    pc = planes.cpu()
    out = torch.zeros(18, 64, 8, dtype=torch.int16, device="cuda")
    for c1 in (100, -5, npos - 100):
        _lib.check(L.rl_selftest_im2col_tmap(ctypes.c_void_p(planes.data_ptr()), ctypes.c_long(npos), ctypes.c_long(positions), W, c1,
                                             ctypes.c_void_p(out.data_ptr()), _lib.cur_stream()))
        torch.cuda.synchronize()
        want = torch.zeros(18, 64, 8, dtype=torch.int16)
        rows = torch.arange(64) + c1
        ok = (rows >= 0) & (rows < positions)
        for half in range(2):
            for iy in range(3):
                for ix in range(3):
                    src = (rows + iy * (W + 2) + ix).clamp(0, npos - 1)
                    want[half * 9 + iy * 3 + ix] = torch.where(ok[:, None], pc[half, src], torch.zeros(1, dtype=torch.int16))
        assert torch.equal(out.cpu(), want), c1
