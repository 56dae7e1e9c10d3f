"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

ctypes binding of oracle/pointops_ref.c, exposed with the exact names and signatures of the
reference's pybind module ``pointops_cuda`` (cpp_wrappers/pointops/src/pointops_api.cpp:13-14)
so the unmodified reference Python (cpp_wrappers/pointops/functions/pointops.py:23,42) can run on
CPU tensors with this object injected as ``sys.modules['pointops_cuda']``.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpointops_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pointops_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        P, I = ctypes.c_void_p, ctypes.c_int
        L.oracle_knnquery.argtypes = [I, I, P, P, P, P, P, P]
        L.oracle_knnquery.restype = I
        L.oracle_furthestsampling.argtypes = [I, I, P, P, P, P, P]
        L.oracle_furthestsampling.restype = I
        L.oracle_fps_block_size.argtypes = [I]
        L.oracle_fps_block_size.restype = I
        L.oracle_set_threads.argtypes = [I]
        L.oracle_set_threads.restype = I
        _lib = L
    return _lib


_REF_SO = os.path.join(_HERE, "_ref", "libpointops_ref_cuda.so")
_ref = None


def ref_cuda():
    """The reference's OWN kernels (oracle/build_ref.sh: knnquery_cuda_kernel.cu / sampling_cuda_kernel.cu compiled
    unmodified for sm_100a). Used when the oracle forward is evaluated on CUDA tensors: that is the reference's algorithm
    with the reference's kernels on this GPU (bench.py `reference_gpu`, scripts/bench_ops.py). Launches on the legacy
    default stream, as the reference's pybind wrappers do."""
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            raise RuntimeError("oracle/_ref/libpointops_ref_cuda.so is not built (oracle/build_ref.sh needs /root/reference)")
        L = ctypes.CDLL(_REF_SO)
        P, I = ctypes.c_void_p, ctypes.c_int
        L.knnquery_cuda_launcher.argtypes = [I, I, P, P, P, P, P, P]
        L.furthestsampling_cuda_launcher.argtypes = [I, I, P, P, P, P, P]
        _ref = L
    return _ref


def have_ref_cuda():
    return os.path.exists(_REF_SO)


def _dev(t, dtype):
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), (t.device, t.dtype)
    return ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype):
    assert t.device.type == "cpu" and t.dtype == dtype and t.is_contiguous(), (t.device, t.dtype)
    return ctypes.c_void_p(t.data_ptr())


def knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, dist2):
    """Same contract as knnquery_cuda (knnquery_cuda_kernel.h:7): caller allocates idx/dist2."""
    if xyz.is_cuda:
        ref_cuda().knnquery_cuda_launcher(int(m), int(nsample), _dev(xyz, torch.float32), _dev(new_xyz, torch.float32),
                                          _dev(offset, torch.int32), _dev(new_offset, torch.int32),
                                          _dev(idx, torch.int32), _dev(dist2, torch.float32))
        return
    rc = lib().oracle_knnquery(int(m), int(nsample), _chk(xyz, torch.float32), _chk(new_xyz, torch.float32),
                               _chk(offset, torch.int32), _chk(new_offset, torch.int32),
                               _chk(idx, torch.int32), _chk(dist2, torch.float32))
    if rc:
        raise RuntimeError("oracle_knnquery failed rc=%d" % rc)


def furthestsampling_cuda(b, n, xyz, offset, new_offset, tmp, idx):
    """Same contract as furthestsampling_cuda (sampling_cuda_kernel.h:7); n is the max segment length."""
    if xyz.is_cuda:
        ref_cuda().furthestsampling_cuda_launcher(int(b), int(n), _dev(xyz, torch.float32), _dev(offset, torch.int32),
                                                  _dev(new_offset, torch.int32), _dev(tmp, torch.float32),
                                                  _dev(idx, torch.int32))
        return
    rc = lib().oracle_furthestsampling(int(b), int(n), _chk(xyz, torch.float32), _chk(offset, torch.int32),
                                       _chk(new_offset, torch.int32), _chk(tmp, torch.float32),
                                       _chk(idx, torch.int32))
    if rc:
        raise RuntimeError("oracle_furthestsampling failed rc=%d" % rc)


def set_threads(n: int) -> int:
    """Host threads for the OpenMP loops (torchrun exports OMP_NUM_THREADS=1); returns the previous maximum."""
    return int(lib().oracle_set_threads(int(n)))


def fps_block_size(n_max: int) -> int:
    return lib().oracle_fps_block_size(int(n_max))


# ---- convenience wrappers (allocate like pointops.py:18-23,39-43) ----
def knn(nsample, xyz, new_xyz, offset, new_offset):
    m = new_xyz.shape[0]
    idx = torch.zeros(m, nsample, dtype=torch.int32, device=xyz.device)
    d2 = torch.zeros(m, nsample, dtype=torch.float32, device=xyz.device)
    knnquery_cuda(m, nsample, xyz, new_xyz, offset, new_offset, idx, d2)
    return idx, d2


def fps(xyz, offset, new_offset):
    b = offset.shape[0]
    ends = offset.tolist()
    n_max = max(e - s for s, e in zip([0] + ends[:-1], ends))
    idx = torch.zeros(int(new_offset[-1]), dtype=torch.int32, device=xyz.device)
    tmp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=xyz.device)
    furthestsampling_cuda(b, n_max, xyz, offset, new_offset, tmp, idx)
    return idx
