"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Runs the UNMODIFIED reference Python (``/root/reference``) on CPU. Exists only in the build
container (the GPU box has no /root/reference): it is used to (a) generate the golden fixtures
under tests/golden/ (tests/golden/make_golden.py) and (b) pin oracle/forward_ref.py.

The reference is CUDA-only (pointops.py:7,21-23,40-42; lib/utils.py:451; modules.py:37). Four
process-local shims make it importable and runnable without a GPU; none edits a reference file:
  1. ``open3d`` stub module (imported at model/model.py:11, RIGA_v2.py:7, lib/utils.py:3; unused in forward)
  2. ``torch.Tensor.cuda`` -> identity
  3. ``torch.cuda.IntTensor/FloatTensor`` -> CPU constructors
  4. ``pointops_cuda`` -> oracle.native (C restatement of the two native kernels)
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("ROITR_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from oracle import native
    if "open3d" not in sys.modules:
        sys.modules["open3d"] = types.ModuleType("open3d")
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.IntTensor = torch.IntTensor
    torch.cuda.FloatTensor = torch.FloatTensor
    mod = types.ModuleType("pointops_cuda")
    mod.knnquery_cuda = native.knnquery_cuda
    mod.furthestsampling_cuda = native.furthestsampling_cuda
    sys.modules["pointops_cuda"] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


class Config(dict):
    """Minimal attribute dict (the reference uses easydict, which is not installed)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def default_config(benchmark="3DLoMatch", **kw):
    """Keys read by RIGA_v2.__init__ (model/RIGA_v2.py:18-51); values from configs/test/{tdmatch,fdmatch}.yaml."""
    four_d = benchmark in ("4DMatch", "4DLoMatch")
    cfg = Config(
        with_cross_pos_embed=True, benchmark=benchmark,
        num_est_coarse_corr=128 if four_d else 256,
        transformer_architecture=["self", "cross", "self", "cross", "self", "cross"],
        mode="test", point_per_patch=64, matching_radius=0.05, num_gt_coarse_corr=128,
        coarse_overlap_threshold=0.1, fine_matching_topk=2 if four_d else 3,
        fine_matching_mutual=True, fine_matching_confidence_threshold=0.05,
        fine_matching_use_dustbin=False, fine_matching_use_global_score=False,
        fine_matching_correspondence_threshold=3)
    cfg.update(kw)
    return cfg


def create_reference_model(config=None):
    install()
    from model.RIGA_v2 import create_model  # the reference's own factory (model/RIGA_v2.py:178)
    return create_model(config or default_config()).eval()


class Trace:
    """Records intermediate tensors of a reference forward by wrapping the reference's pointops API."""

    def __init__(self):
        self.fps = []
        self.knn = []

    def __enter__(self):
        install()
        from cpp_wrappers.pointops.functions import pointops
        self._p = pointops
        self._fps, self._knn = pointops.furthestsampling, pointops.knnquery

        def fps(xyz, o, no):
            r = self._fps(xyz, o, no)
            self.fps.append(r.clone())
            return r

        def knn(ns, xyz, new_xyz, o, no):
            r = self._knn(ns, xyz, new_xyz, o, no)
            self.knn.append((ns, r[0].clone(), r[1].clone()))
            return r

        pointops.furthestsampling, pointops.knnquery = fps, knn
        import lib.utils as lu
        self._lu, self._lu_knn = lu, lu.knnquery
        lu.knnquery = knn
        return self

    def __exit__(self, *a):
        self._p.furthestsampling, self._p.knnquery = self._fps, self._knn
        self._lu.knnquery = self._lu_knn
