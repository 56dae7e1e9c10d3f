"""Deterministic synthetic inputs: Gaussian-blob point-cloud pairs with unit normals and a
known rigid transform, plus seeded weights keyed by the reference's state_dict schema.

There is no network for datasets or checkpoints, so every test and benchmark in this repo runs on these.
Shapes/dtypes follow what lib/tester.py:45-50 feeds RIGA_v2.forward (contiguous f32, feats = ones (N,1),
rot (3,3), trans (3,1)).
"""
import math

import torch


def synthetic_pair(index: int, n_points: int, overlap_scale: float = 1.4, deform: bool = False):
    """Pair ``index`` with ``n_points`` per cloud. Returns a dict of CPU tensors.

    ``tgt`` = first N scene points, ``src`` = last N scene points moved by the inverse of (rot, trans), so that
    ``src @ rot.T + trans.T`` re-aligns with the scene (the convention of lib/utils.py:505,562).
    """
    g = torch.Generator().manual_seed(1000 + index)
    n_scene = int(math.ceil(overlap_scale * n_points))
    n_blobs = 32
    centres = (torch.rand(n_blobs, 3, generator=g) * 3.0 - 1.5)
    which = torch.randint(0, n_blobs, (n_scene,), generator=g)
    c = centres[which]
    pts = c + 0.15 * torch.randn(n_scene, 3, generator=g)
    nrm = pts - c + 0.05 * torch.randn(n_scene, 3, generator=g)
    nrm = nrm / nrm.norm(dim=1, keepdim=True).clamp_min(1e-12)

    q = torch.randn(4, generator=g)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    rot = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=torch.float32)
    trans = (torch.rand(3, 1, generator=g) - 0.5)

    tgt, tgt_n = pts[:n_points].clone(), nrm[:n_points].clone()
    src_scene, src_scene_n = pts[n_scene - n_points:], nrm[n_scene - n_points:]
    src = (src_scene - trans.T) @ rot          # src @ rot.T + trans.T == src_scene
    src_n = src_scene_n @ rot
    src = src + 0.002 * torch.randn(n_points, 3, generator=g)
    tgt = tgt + 0.002 * torch.randn(n_points, 3, generator=g)
    src_raw = src.clone()
    if deform:  # 4DMatch-style smooth non-rigid flow on the source
        a = torch.rand(3, 3, generator=g)
        src = src_raw + 0.05 * torch.sin(2 * math.pi * (src_raw @ a.T))
    f32 = lambda t: t.to(torch.float32).contiguous()
    return dict(src_pcd=f32(src), tgt_pcd=f32(tgt), src_feats=torch.ones(n_points, 1), tgt_feats=torch.ones(n_points, 1),
                src_normals=f32(src_n), tgt_normals=f32(tgt_n), rot=f32(rot), trans=f32(trans), src_raw_pcd=f32(src_raw))


FORWARD_ARG_ORDER = ("src_pcd", "tgt_pcd", "src_feats", "tgt_feats", "src_normals", "tgt_normals", "rot", "trans",
                     "src_raw_pcd")


def forward_args(pair: dict, device=None):
    """Positional arguments in the order of RIGA_v2.forward (model/RIGA_v2.py:58)."""
    return [pair[k].to(device) if device is not None else pair[k] for k in FORWARD_ARG_ORDER]


def seeded_state_dict(schema, seed: int = 42, fine_scale: float = 8.0):
    """Weights as a pure function of (schema, seed), independent of module construction order.

    ``schema``: list of (name, shape) in state_dict order (tests/golden/state_dict_schema_f{1,2}.json, dumped from the
    reference model). Linear weights ~ U(±1/sqrt(fan_in)); biases small; LayerNorm affine near (1, 0) but not
    trivial, so that parity tests exercise gamma/beta; ``div_term`` buffers keep their defining formula
    (positional_encoding.py:43-45); OT ``alpha`` = 1 (modules.py:18).
    """
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in schema:
        shape = tuple(shape)
        if name.endswith("div_term"):
            d_model = shape[0] * 2
            sd[name] = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        elif name.endswith("alpha"):
            sd[name] = torch.tensor(1.0)
        elif len(shape) == 2:
            bound = 1.0 / math.sqrt(shape[1])
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
            if name == 'fine_proj.weight':
                sd[name] *= fine_scale   # random features are nearly flat; sharpen the fine scores so correspondences exist
        elif len(shape) == 1 and (".norm." in name or ".bn2." in name or "pos_norm" in name or _is_ln(name)):
            if name.endswith("weight"):
                sd[name] = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
            else:
                sd[name] = 0.05 * (torch.rand(shape, generator=g) * 2 - 1)
        else:  # linear bias
            sd[name] = 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
    return sd


def _is_ln(name: str) -> bool:
    # LayerNorms inside nn.Sequential of TransitionUp: dec*.0.linear{1,2}.1.{weight,bias} (model/model.py:90-97)
    parts = name.split(".")
    return len(parts) >= 2 and parts[-2] == "1" and "linear" in parts[-3]
