"""descriptor probe: MN-major A in the 128-byte-swizzle layout a [pixels][64 ch] tensor-map box lands in"""
import ctypes, itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sim2real_lane_segment_b200 import _lib
L = _lib.lib()
L.rl_selftest_umma_raw.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_ulonglong, ctypes.c_int, ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
def desc(lbo, sbo, layout):
    return ((lbo >> 4) & 0x3FFF) << 16 | ((sbo >> 4) & 0x3FFF) << 32 | 1 << 46 | layout << 61
def idesc(N, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((128 >> 4) << 24)
torch.manual_seed(0)
K, M, N = 64, 128, 16          # K = pixels, M = channels
A = torch.randn(K, M).bfloat16()   # [px][ch] as it sits in the arena
B = torch.randn(K, N).bfloat16()   # [px][co]
ref = A.float().t() @ B.float()    # [M][N]
# A image: two 64-channel atoms (TMA boxes), each [K rows][128 B] with the 16-byte chunk index XORed with (row & 7)
a16 = A.view(torch.int16).numpy()
img = np.zeros((2, K, 64), dtype=np.int16)
for atom in range(2):
    for p in range(K):
        for j in range(8):
            img[atom, p, ((j ^ (p & 7)) * 8):((j ^ (p & 7)) * 8 + 8)] = a16[p, atom * 64 + j * 8: atom * 64 + j * 8 + 8]
a_img = torch.from_numpy(img.reshape(-1).copy()).cuda()
# B image: chunk layout T[j][r][8] (2 chunks of 8 co), MN-major no swizzle: SBO = 16*R, LBO = 128
b16 = B.view(torch.int16).numpy()
bimg = np.zeros((2, K, 8), dtype=np.int16)
for j in range(2):
    bimg[j] = b16[:, j * 8:(j + 1) * 8]
b_img = torch.from_numpy(bimg.reshape(-1).copy()).cuda()
b_t = desc(128, 16 * K, 0)
atom_bytes = K * 128
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for (lbo, sbo, kstep) in [(atom_bytes, 1024, 2048), (1024, atom_bytes, 2048), (atom_bytes, 1024, 256), (1024, atom_bytes, 256),
                          (atom_bytes, 128, 2048), (128, atom_bytes, 2048)]:
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = L.rl_selftest_umma_raw(a_img.data_ptr(), a_img.numel() * 2, desc(lbo, sbo, 2), kstep, b_img.data_ptr(), b_img.numel() * 2,
                                b_t, 256, idesc(N, 1, 1), K // 16, N, D.data_ptr(), st)
    torch.cuda.synchronize()
    err = float((D.cpu() - ref).abs().max())
    print(f"LBO {lbo:6d} SBO {sbo:6d} kstep {kstep:5d}: rc {rc} max err {err:.4f}")
