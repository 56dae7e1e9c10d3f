"""ctypes binding of libroitr_b200.so (the C-ABI in include/roitr_b200.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, this raises. torch is used only to
own device memory and to name the current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libroitr_b200.so")
_lib = None

c_int, c_float, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_void_p


class RoitrError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RoitrError("libroitr_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
                             "g.build()'`; there is no CPU/PyTorch fallback for the hot path." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.roitr_last_error.restype = ctypes.c_char_p
        _lib.roitr_abi_version.restype = c_int
    return _lib


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    if not t.is_cuda:
        raise RoitrError("expected a CUDA tensor, got %s" % t.device)
    if not t.is_contiguous():
        raise RoitrError("expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise RoitrError("expected dtype %s, got %s" % (dtype, t.dtype))
    return c_void_p(t.data_ptr())


def f32(t):
    return ptr(t, torch.float32)


def i32(t):
    return ptr(t, torch.int32)


def call(name, *args):
    """Call an int-returning entry point; raise with the library's message on failure."""
    fn = getattr(lib(), name)
    fn.restype = c_int
    rc = fn(*args)
    if rc != 0:
        raise RoitrError("%s failed (rc=%d): %s" % (name, rc, lib().roitr_last_error().decode()))
