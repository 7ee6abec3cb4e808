"""ctypes loader for ``librightlane.so`` (the C-ABI boundary declared in ``include/rightlane.h``).

The product path has no CPU fallback: if the shared library is missing, importing any op that
needs it raises.  ``build()`` (nvcc, sm_100a) lives in ``_build.py``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librightlane.so")
_lib = None


class RightLaneError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RightLaneError(
                f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the CUDA path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.rl_last_error.restype = ctypes.c_char_p
    return _lib


def check(status):
    if status != 0:
        raise RightLaneError(f"librightlane status {status}: {lib().rl_last_error().decode()}")


def ptr(t):
    """device (or host) pointer of a torch tensor as c_void_p; None -> NULL"""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
