import ctypes, torch, sys
sys.path.insert(0, '.')
from sim2real_lane_segment_b200 import _lib
L = _lib.lib()
out = torch.zeros(128, 2, 16, dtype=torch.int32, device='cuda')
rc = L.rl_selftest_tmem_shapes(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize(); print('rc', rc)
o = out.cpu()
for t in (0, 1, 2, 3, 4, 5, 31, 32, 33, 127):
    for h in (0, 1):
        print(t, h, [(int(v) >> 8, int(v) & 255) for v in o[t, h]])
