"""ctypes binding of libroitr_b200.so (the C-ABI in include/roitr_b200.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, this raises. torch is used only to
own device memory and to name the current stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ROITR_B200_LIB") or os.path.join(_HERE, "lib", "libroitr_b200.so")   # override: A/B runs of a variant build
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


_ARG_DEVICE = None    # device of the most recent tensor argument (every entry point takes its tensors before the stream)


def _note(t):
    global _ARG_DEVICE
    _ARG_DEVICE = t.device


def stream_ptr():
    """The current stream OF THE DEVICE THE ARGUMENTS LIVE ON (always the last argument of an entry point, so the tensor
    arguments have been seen): a model on cuda:1 must not launch on cuda:0's stream just because cuda:0 is current."""
    return c_void_p(torch.cuda.current_stream(_ARG_DEVICE).cuda_stream)


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
    _note(t)
    return c_void_p(t.data_ptr())


def f32(t):
    return ptr(t, torch.float32)


def i32(t):
    return ptr(t, torch.int32)


# kernels launched by one call of each entry point (for bench.py's gpu_launches claim; memsets not counted)
KERNELS_PER_CALL = {
    "roitr_knnquery_n": 2, "roitr_knn_ppf_n": 2, "roitr_knn_ppf_grid": 2, "roitr_knn_ppf_grid_q": 2, "roitr_knn_grid_build": 4, "roitr_knn_grid_build_target": 4, "roitr_furthestsampling_cfg": 1, "roitr_interpolate": 1,
    "roitr_gather_rows": 1, "roitr_linear": 1, "roitr_linear_tc_packed": 1, "roitr_linear_ln_tc_packed": 1, "roitr_row_epilogue": 1, "roitr_segment_mean": 1,
    "roitr_concat_segment": 1, "roitr_local_attention": 1, "roitr_local_attention_ordered": 1, "roitr_geo_knn": 1, "roitr_geo_embedding_tc": 1, "roitr_geo_embedding_tc_batched": 1, "roitr_geo_knn_batched": 1, "roitr_geo_embedding_table": 1, "roitr_gemm_tc_batched": 1, "roitr_geo_self_scores": 1, "roitr_geo_self_scores_ld": 1, "roitr_softmax_rows": 1,
    "roitr_point_to_node_batched": 4, "roitr_compact_flags": 3, "roitr_compact_flags_batched": 3, "roitr_coarse_matching_batched": 5, "roitr_coarse_matching_adaptive_batched": 6,
    "roitr_fine_matching_batched": 1, "roitr_fine_gather_batched": 1, "roitr_pad_transform": 1, "roitr_pad_transform_batched": 1, "roitr_node_occlusion_batched": 1,
    "roitr_node_overlaps_batched": 3, "roitr_corr_gather_batched": 1, "roitr_ransac_correspondences": 2, "roitr_weighted_procrustes": 1, "roitr_estimate_normals": 1,
}
STATS = {"launches": 0, "calls": {}}
TIMED = {}        # entry-point name -> list of (start_event, end_event); filled only for names present as keys
RECORD_ARGS = False
ARGS = {}         # entry-point name -> list of leading integer arguments per call (bench.py's work model)
PTRS = {}         # entry-point name -> per call, which pointer arguments were non-NULL (optional operands: residuals, ...)


def reset_stats():
    STATS["launches"] = 0
    STATS["calls"] = {}
    for k in TIMED:
        TIMED[k] = []
    ARGS.clear()
    PTRS.clear()


def call(name, *args):
    """Call an int-returning entry point; raise with the library's message on failure."""
    fn = getattr(lib(), name)
    fn.restype = c_int
    timed = TIMED.get(name)
    if timed is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dev = _ARG_DEVICE
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        with torch.cuda.device(dev):      # the launch itself must happen with the tensors' device current
            rc = fn(*args)
    else:
        rc = fn(*args)
    if timed is not None:
        e1.record()
        timed.append((e0, e1))
    if rc != 0:
        raise RoitrError("%s failed (rc=%d): %s" % (name, rc, lib().roitr_last_error().decode()))
    if RECORD_ARGS:
        ints = []
        for a in args:
            if isinstance(a, c_int):
                ints.append(a.value)
            else:
                break
        ARGS.setdefault(name, []).append(ints)
        PTRS.setdefault(name, []).append([bool(a.value) for a in args if isinstance(a, c_void_p)])
    STATS["launches"] += KERNELS_PER_CALL.get(name, 1)
    STATS["calls"][name] = STATS["calls"].get(name, 0) + 1
