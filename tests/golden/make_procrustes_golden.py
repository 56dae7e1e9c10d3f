"""Generates tests/golden/procrustes_golden.npz by running the UNMODIFIED reference weighted_procrustes
(/root/reference/lib/utils.py:159-218) on CPU through oracle/reference_shim.py. Run in the build container only."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim as rs  # noqa: E402


def cases():
    g = torch.Generator().manual_seed(7)
    out = []
    for B, N, noise, thr in [(1, 64, 0.0, 0.0), (3, 500, 0.01, 0.0), (2, 3000, 0.02, 0.3), (4, 16, 0.001, 0.0)]:
        src = torch.randn(B, N, 3, generator=g)
        q = torch.randn(B, 4, generator=g)
        q = q / q.norm(dim=1, keepdim=True)
        w, x, y, z = q.unbind(1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                         2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).view(B, 3, 3)
        t = torch.randn(B, 1, 3, generator=g)
        tgt = src @ R.transpose(1, 2) + t + noise * torch.randn(B, N, 3, generator=g)
        wts = torch.rand(B, N, generator=g)
        out.append((src, tgt, wts, thr))
    return out


if __name__ == "__main__":
    rs.install()
    from lib.utils import weighted_procrustes
    z = {}
    for i, (src, tgt, w, thr) in enumerate(cases()):
        R, t = weighted_procrustes(src, tgt, w, weight_thresh=thr)
        z.update({"src%d" % i: src.numpy(), "tgt%d" % i: tgt.numpy(), "w%d" % i: w.numpy(), "thr%d" % i: np.float32(thr),
                  "R%d" % i: R.numpy(), "t%d" % i: t.numpy()})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "procrustes_golden.npz"), n=np.int32(len(cases())), **z)
    print("wrote", len(cases()), "cases")
