"""what does HBM deliver for row-strided channel windows of an NHWC arena?  (B=64, 120x160 level-0 arena, Ct=288)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sim2real_lane_segment_b200 import _lib
L = _lib.lib()
L.rl_selftest_stream.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
npx, Ct = 64 * 120 * 160, 288
arena = torch.randn(npx, Ct, device="cuda").to(torch.bfloat16)
sink = torch.zeros(4, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(c0, nch, mode, grid):
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert L.rl_selftest_stream(arena.data_ptr(), npx, Ct, c0, nch, mode, grid, sink.data_ptr(), st) == 0
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for mode in (0, 1):
    for c0, nch in ((0, 16), (0, 64), (64, 64), (0, 128), (0, 208), (0, 272), (0, 288), (272, 16)):
        for grid in (148 * 8, 148 * 32):
            ms = run(c0, nch, mode, grid)
            gb = npx * nch * 2 * (2 if mode else 1) / 1e9
            print(f"mode {mode} window [{c0:3d},{c0+nch:3d}) grid {grid:5d}: {ms*1e3:7.1f} us  {gb/ms*1e3:7.0f} GB/s")
