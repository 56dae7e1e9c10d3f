import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import forward_ref as fr
from roitr_b200 import model
from roitr_b200.synthetic import synthetic_pair, forward_args
from tests import parity
from tests.helpers import CONFIG_3D, weights
pair = synthetic_pair(0, 20000); sd = weights(1)
m = model.create_model(CONFIG_3D); m.load_state_dict(sd); m = m.cuda().eval()
aux = {}
out = m(*forward_args(pair, "cuda:0"), _aux=aux); torch.cuda.synchronize()
with torch.no_grad(): ref = fr.riga_forward(sd, CONFIG_3D, *forward_args(pair), with_aux=True)
ra = ref["_aux"]
for side in ("src", "tgt"):
    G, R = aux[side + "_levels"], ra[side + "_levels"]
    for li in range(4):
        gi, ri = G[li]["idx"].cpu().long(), R[li]["idx"]
        bad = (gi != ri).any(1).nonzero().flatten()
        dx = (G[li]["x"].cpu() - R[li]["x"]).abs().amax(1)
        badx = (dx > 2e-4).nonzero().flatten()
        print(side, "L%d" % (li + 1), "knn rows differing:", bad.tolist()[:8], len(bad), "| feat rows >2e-4:", len(badx), badx.tolist()[:8], "max %.2e" % dx.max())
        if len(bad):
            r = int(bad[0]); p = R[li]["p"]
            print("   row", r, "gpu", gi[r].tolist(), "ref", ri[r].tolist())
            print("   d2 gpu", ((p[gi[r]] - p[r]) ** 2).sum(1).tolist()); print("   d2 ref", ((p[ri[r]] - p[r]) ** 2).sum(1).tolist())
    gk, rk = aux[side + "_node_knn_indices"].cpu().long(), ra[side + "_node_knn_indices"]
    print(side, "partition rows differing", int((gk != rk).any(1).sum()), "entries", int((gk != rk).sum()))
