"""The forward hot path of RoITr sequenced over libroitr_b200 kernels.

Mirrors RIPointTransformer.forward (model/model.py:187-237) and RIGA_v2.forward (model/RIGA_v2.py:58-175) stage by
stage; every numeric step is a CUDA kernel from roitr_b200/csrc (see ops.py). PyTorch is used for allocation, dtype
casts of the returned dict (int32 -> int64, uint8 -> bool), tiny index compositions and ONE device->host read of the
three data-dependent output lengths at the very end (the reference syncs ~15 times per forward: .item() per level,
every torch.nonzero, masks.sum()).

Differences from the reference that do not change results:
  * level-1 kNN+PPF is computed once (the reference recomputes the identical query, model/model.py:75 and :31);
  * the dead all-pairs PPF on level-4 nodes (model/model.py:208-212) is skipped;
  * positional projections are folded (csrc/local_attn.cu, csrc/geo.cu headers).
"""
import math

import torch

from . import ops

STRIDES = (1, 4, 4, 4)
NSAMPLE = (8, 16, 16, 16)
BLOCKS = (2, 3, 3, 3)
HEADS = 4


# ------------------------------------------------------------------------------------------------ weight packing
class Packed(dict):
    """name -> contiguous f32 CUDA tensor, plus derived (stacked / folded / transposed) weights."""


def pack_weights(state_dict, device, architecture):
    W = Packed()
    for k, v in state_dict.items():
        W[k] = v.detach().to(device=device, dtype=torch.float32).contiguous()
    hp = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        local = [k[: -len(".embedding.proj.weight")] for k in W if k.endswith(".embedding.proj.weight")]
        for p in local:  # LocalPPFTransformer prefixes
            a = p + ".transformer.attention"
            We, be = W[p + ".embedding.proj.weight"].double(), W[p + ".embedding.proj.bias"].double()
            for nm in ("p", "vp"):
                Wx, bx = W[a + ".proj_%s.weight" % nm].double(), W[a + ".proj_%s.bias" % nm].double()
                W[p + "#A" + nm] = (Wx @ We).float().contiguous()                 # (C,4)
                W[p + "#c" + nm] = (Wx @ be + bx).float().contiguous()            # (C,)
            W[p + "#Wqkv"] = torch.cat([W[a + ".proj_%s.weight" % t] for t in "qkv"], 0).contiguous()
            W[p + "#bqkv"] = torch.cat([W[a + ".proj_%s.bias" % t] for t in "qkv"], 0).contiguous()
        g = "backbone.global_transformer"
        C = W[g + ".in_proj.weight"].shape[0]
        c = C // HEADS
        for i, kind in enumerate(architecture):
            a = "%s.transformer.layers.%d.attention.attention" % (g, i)
            if kind == "self":
                W[a + "#Wqkv"] = torch.cat([W[a + ".proj_%s.weight" % t] for t in "qkv"], 0).contiguous()
                W[a + "#bqkv"] = torch.cat([W[a + ".proj_%s.bias" % t] for t in "qkv"], 0).contiguous()
                Wp = W[a + ".proj_p.weight"]                                       # (C, C): p = Wp e + bp
                # gq[n,h,:] = sum_{k in head h} q[n, h*c+k] * Wp[h*c+k, :]  ->  per head a (C x c) matrix, K = c
                W[a + "#WpT"] = torch.stack([Wp[h * c:(h + 1) * c, :].t().contiguous() for h in range(HEADS)], 0).contiguous()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = hp
    return W


# ------------------------------------------------------------------------------------------------ local layers
def _lin(W, p, x, **kw):
    return ops.linear(x, W[p + ".weight"], W[p + ".bias"], **kw)


def _ln(W, p, x, **kw):
    return ops.row_epilogue(x, gamma=W[p + ".weight"], beta=W[p + ".bias"], **kw)


def local_ppf_transformer(W, p, feats, node_idx, group_idx, ppf):
    """LocalPPFTransformer.forward (ppftransformer.py:243-253): (n,Cin) -> (m,Cout)."""
    C = W[p + ".in_proj.weight"].shape[0]
    f = _lin(W, p + ".in_proj", feats)
    qkv = ops.linear(f, W[p + "#Wqkv"], W[p + "#bqkv"])
    h = ops.local_attention(qkv, C, node_idx, group_idx, ppf, W[p + "#Ap"], W[p + "#cp"], W[p + "#Avp"], W[p + "#cvp"])
    t = _lin(W, p + ".transformer.linear", h)
    y = _ln(W, p + ".transformer.norm", t, res_pre=f, res_pre_index=node_idx, mode=ops.MODE_LN)
    return _lin(W, p + ".out_proj", y)


def block(W, p, x, idx, ppf):
    """RIPointTransformerBlock.forward (model/model.py:131-142) with cached (idx, ppf)."""
    y = local_ppf_transformer(W, p + ".transformer.transformer", x, None, idx, ppf)
    return _ln(W, p + ".bn2", y, res_post=x, mode=ops.MODE_LN | ops.MODE_RELU)


def _offsets(ends, device):
    return torch.tensor(ends, dtype=torch.int32, device=device)


def encode(W, pts, feats, nrm, ends, fps_cluster=0):
    """enc1..enc4 for a batch of clouds concatenated along dim 0 (``ends`` = host list of cumulative sizes)."""
    dev = pts.device
    levels = []
    o = _offsets(ends, dev)
    x = feats
    for li in range(4):
        p = "backbone.enc%d" % (li + 1)
        k = NSAMPLE[li]
        if STRIDES[li] != 1:
            sizes = [e - s for s, e in zip([0] + ends[:-1], ends)]
            new_sizes = [n // STRIDES[li] for n in sizes]
            new_ends = [sum(new_sizes[: i + 1]) for i in range(len(new_sizes))]
            no = _offsets(new_ends, dev)
            down_idx, n_p = ops.fps(pts, o, no, max(sizes), new_ends[-1], per_segment_rule=True, cluster=fps_cluster)
            n_n = ops.gather_rows(nrm, down_idx)
            gidx, gppf, _ = ops.knn_ppf(k, pts, nrm, n_p, n_n, o, no)
            x = local_ppf_transformer(W, p + ".0.transformer", x, down_idx, gidx, gppf)
            pts, nrm, o, ends = n_p, n_n, no, new_ends
            idx, ppf, _ = ops.knn_ppf(k, pts, nrm, pts, nrm, o, o)
        else:
            down_idx = None
            idx, ppf, _ = ops.knn_ppf(k, pts, nrm, pts, nrm, o, o)   # shared by the TD and the blocks of level 1
            x = local_ppf_transformer(W, p + ".0.transformer", x, None, idx, ppf)
        for bi in range(1, BLOCKS[li]):
            x = block(W, "%s.%d" % (p, bi), x, idx, ppf)
        levels.append(dict(p=pts, n=nrm, x=x, o=o, ends=list(ends), idx=idx, ppf=ppf, down_idx=down_idx))
    return levels


def decode(W, L):
    """dec4..dec1 (model/model.py:223-231): TransitionUp + one block per level, reusing the encoder's (idx, ppf)."""
    l4 = L[3]
    p = "backbone.dec4.0"
    g = _lin(W, p + ".linear2.0", ops.segment_mean(l4["x"], l4["o"]), relu=True)
    y = _ln(W, p + ".linear1.1", _lin(W, p + ".linear1.0", ops.concat_segment(l4["x"], g, l4["o"])),
            mode=ops.MODE_LN | ops.MODE_RELU)
    xs = [None, None, None, block(W, "backbone.dec4.1", y, l4["idx"], l4["ppf"])]
    for li in (2, 1, 0):
        p = "backbone.dec%d.0" % (li + 1)
        fine, coarse = L[li], L[li + 1]
        a = _ln(W, p + ".linear1.1", _lin(W, p + ".linear1.0", fine["x"]), mode=ops.MODE_LN | ops.MODE_RELU)
        b = _ln(W, p + ".linear2.1", _lin(W, p + ".linear2.0", xs[li + 1]), mode=ops.MODE_LN | ops.MODE_RELU)
        nn_idx, _, nn_dist = ops.knn_ppf(3, coarse["p"], None, fine["p"], None, coarse["o"], fine["o"], drop_first=0,
                                         want_ppf=False, want_dist=True)
        y = ops.interpolate(nn_idx, nn_dist, b, base=a)
        xs[li] = block(W, "backbone.dec%d.1" % (li + 1), y, fine["idx"], fine["ppf"])
    return xs


# ------------------------------------------------------------------------------------------------ global transformer
def _ffn(W, p, x):
    h = _lin(W, p + ".squeeze", _lin(W, p + ".expand", x, relu=True))
    return _ln(W, p + ".norm", h, res_pre=x, mode=ops.MODE_LN)


def _self_layer(W, lp, x, E):
    a = lp + ".attention.attention"
    C = x.shape[1]
    c = C // HEADS
    N = x.shape[0]
    qkv = ops.linear(x, W[a + "#Wqkv"], W[a + "#bqkv"])
    gq = torch.empty(N, HEADS * C, dtype=torch.float32, device=x.device)
    for h in range(HEADS):
        ops.linear(qkv[:, h * c:(h + 1) * c], W[a + "#WpT"][h], None, out=gq[:, h * C:(h + 1) * C], M=N, K=c)
    hidden, G = ops.geo_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], C, E=E, gq=gq, bp=W[a + ".proj_p.bias"])
    Wvp, bvp = W[a + ".proj_vp.weight"], W[a + ".proj_vp.bias"]
    G2 = G.view(N, HEADS * C)
    pos = torch.empty(N, C, dtype=torch.float32, device=x.device)
    for h in range(HEADS):
        ops.linear(G2[:, h * C:(h + 1) * C], Wvp[h * c:(h + 1) * c], bvp[h * c:(h + 1) * c], out=pos[:, h * c:(h + 1) * c],
                   M=N, K=C)
    y = _ln(W, lp + ".attention.norm", _lin(W, lp + ".attention.linear", hidden), res_pre=x, mode=ops.MODE_LN)
    pos = _ln(W, lp + ".attention.pos_norm", _lin(W, lp + ".attention.pos_linear", pos), mode=ops.MODE_LN)
    return _ffn(W, lp + ".output", y), _ffn(W, lp + ".pos_proj", pos)


def _cross_layer(W, lp, x, y, pos_x, pos_y):
    a = lp + ".attention.attention"
    C = x.shape[1]
    q = ops.linear(x, W[a + ".proj_q.weight"], W[a + ".proj_q.bias"], a_add=pos_x)
    k = ops.linear(y, W[a + ".proj_k.weight"], W[a + ".proj_k.bias"], a_add=pos_y)
    v = ops.linear(y, W[a + ".proj_v.weight"], W[a + ".proj_v.bias"])
    hidden = ops.geo_attention(q, k, v, C)
    z = _ln(W, lp + ".attention.norm", _lin(W, lp + ".attention.linear", hidden), res_pre=x, mode=ops.MODE_LN)
    return _ffn(W, lp + ".output", z)


def geometric_transformer(W, architecture, pts0, pts1, f0, f1, sigma_d=0.2, sigma_a=15.0):
    """GeometricTransformer.forward (geotransformer.py:94-133); '0' = src, '1' = tgt as called at model/model.py:214."""
    g = "backbone.global_transformer"
    e = g + ".embedding"
    embs = []
    for pts in (pts0, pts1):
        nn3 = ops.geo_knn(pts, 3)
        embs.append(ops.geo_embedding(pts, nn3, W[e + ".proj_d.weight"], W[e + ".proj_d.bias"], W[e + ".proj_a.weight"],
                                      W[e + ".proj_a.bias"], W[e + ".embedding.div_term"], sigma_d, sigma_a))
    f0, f1 = _lin(W, g + ".in_proj", f0), _lin(W, g + ".in_proj", f1)
    pos0 = pos1 = None
    for i, kind in enumerate(architecture):
        lp = "%s.transformer.layers.%d" % (g, i)
        if kind == "self":
            f0, pos0 = _self_layer(W, lp, f0, embs[0])
            f1, pos1 = _self_layer(W, lp, f1, embs[1])
        else:
            f0 = _cross_layer(W, lp, f0, f1, pos0, pos1)
            f1 = _cross_layer(W, lp, f1, f0, pos1, pos0)
    return _lin(W, g + ".out_proj", f0), _lin(W, g + ".out_proj", f1), embs


# ------------------------------------------------------------------------------------------------ backbone
def backbone_forward(W, architecture, s_pxon, t_pxon, src_deformed, aux=None):
    """RIPointTransformer.forward: -> (s_p4, s_g_x4, src_deformed_pcd, s_x1, t_p4, t_g_x4, t_p1, t_x1)."""
    s_p, s_x, s_o, s_n = s_pxon
    t_p, t_x, t_o, t_n = t_pxon
    S = encode(W, s_p, s_x, s_n, [int(s_p.shape[0])])
    T = encode(W, t_p, t_x, t_n, [int(t_p.shape[0])])
    s_g, t_g, embs = geometric_transformer(W, architecture, S[3]["p"], T[3]["p"], S[3]["x"], T[3]["x"])
    s_dec, t_dec = decode(W, S), decode(W, T)
    d3 = S[1]["down_idx"].long()[S[2]["down_idx"].long()]          # index-chain composition (model/model.py:233-234)
    d4 = d3[S[3]["down_idx"].long()]
    s_nodes = ops.gather_rows(src_deformed, d4)
    if aux is not None:
        aux.update(src_levels=S, tgt_levels=T, src_node_idx=d4, src_dec=s_dec, tgt_dec=t_dec, emb0=embs[0], emb1=embs[1])
    return s_nodes, s_g, src_deformed, s_dec[0], T[3]["p"], t_g, T[0]["p"], t_dec[0]


# ------------------------------------------------------------------------------------------------ pipeline
def riga_forward(W, cfg, src_pcd, tgt_pcd, src_feats, tgt_feats, src_normals, tgt_normals, rot, trans, src_raw_pcd,
                 aux=None):
    """RIGA_v2.forward (eval). Returns the reference's 22-key dict (exact-size tensors; one host sync at the end)."""
    four_d = cfg["benchmark"] not in ("3DMatch", "3DLoMatch")
    if four_d:
        raise NotImplementedError("AdaptiveSuperPointMatching (4DMatch head) is scheduled after the 3DMatch path; see DESIGN.md")
    dev = src_pcd.device
    K = int(cfg["point_per_patch"])
    Ns, Nt = src_raw_pcd.shape[0], tgt_pcd.shape[0]
    so, to = _offsets([Ns], dev), _offsets([Nt], dev)
    (src_nodes, src_nf, src_pts, src_pf, tgt_nodes, tgt_nf, tgt_pts, tgt_pf) = backbone_forward(
        W, cfg["transformer_architecture"], [src_raw_pcd, src_feats, so, src_normals],
        [tgt_pcd, tgt_feats, to, tgt_normals], src_pcd, aux)
    src_nf = ops.row_epilogue(_lin(W, "coarse_proj", src_nf), mode=ops.MODE_L2NORM)
    tgt_nf = ops.row_epilogue(_lin(W, "coarse_proj", tgt_nf), mode=ops.MODE_L2NORM)
    src_pf, tgt_pf = _lin(W, "fine_proj", src_pf), _lin(W, "fine_proj", tgt_pf)

    # 2. partition + ground-truth bookkeeping
    _, s_nm, s_ki, s_km = ops.point_to_node(src_pts, src_nodes, K)
    _, t_nm, t_ki, t_km = ops.point_to_node(tgt_pts, tgt_nodes, K)
    Ms, Mt = src_nodes.shape[0], tgt_nodes.shape[0]
    ov, ov_flag = ops.node_overlaps(tgt_nodes, src_nodes, t_ki, s_ki, t_km, s_km, t_nm, s_nm, tgt_pts, src_pts, rot,
                                    trans, float(cfg["matching_radius"]))
    gt_flat, gt_count = ops.compact_flags(ov_flag, Mt * Ms)
    gt_idx, gt_ov = ops.corr_gather(Mt * Ms, Ms, gt_flat, gt_count, ov)
    t_pad = ops.pad_transform(tgt_pts)
    s_pad_t = ops.pad_transform(src_pts, rot, trans)
    o_t, o_s = _offsets([Nt + 1], dev), _offsets([Ns + 1], dev)
    _, _, t_nn = ops.knn_ppf(1, s_pad_t, None, t_pad, None, o_s, o_t, drop_first=0, want_ppf=False, want_dist=True)
    _, _, s_nn = ops.knn_ppf(1, t_pad, None, s_pad_t, None, o_t, o_s, drop_first=0, want_ppf=False, want_dist=True)
    t_occ = ops.node_occlusion(t_ki, t_km, t_nm, t_nn.view(-1))
    s_occ = ops.node_occlusion(s_ki, s_km, s_nm, s_nn.view(-1))

    # 3. coarse matching   (called as (tgt, src), model/RIGA_v2.py:121)
    Pmax = int(cfg["num_est_coarse_corr"])
    t_ci, s_ci, node_sc, p_count = ops.coarse_matching(tgt_nf, src_nf, t_nm, s_nm, Pmax, dual=True)

    # 4-6. fine scoring + OT + fine matching
    scores, flags = ops.fine_matching(tgt_pf, src_pf, t_ki, s_ki, t_km, s_km, t_ci, s_ci, p_count,
                                      W["optimal_transport.alpha"].view(1), 100, int(cfg["fine_matching_topk"]),
                                      bool(cfg["fine_matching_mutual"]), float(cfg["fine_matching_confidence_threshold"]))
    cap = Pmax * K * int(cfg["fine_matching_topk"])
    c_flat, c_count = ops.compact_flags(flags, cap)
    t_cp, s_cp, c_sc = ops.fine_gather(cap, c_flat, c_count, scores, t_ci, s_ci, t_ki, s_ki, tgt_pts, src_pts)

    counts = torch.cat([p_count, gt_count, c_count]).tolist()       # the one host sync
    P, n_gt, n_c = counts
    n_c = min(n_c, cap)
    s_ci, t_ci = s_ci[:P].long(), t_ci[:P].long()
    s_cki, t_cki = s_ki.long()[s_ci], t_ki.long()[t_ci]
    out = dict(
        src_points=src_pts, tgt_points=tgt_pts, src_nodes=src_nodes, tgt_nodes=tgt_nodes,
        src_point_feats=src_pf, tgt_point_feats=tgt_pf, src_node_feats=src_nf, tgt_node_feats=tgt_nf,
        gt_node_corr_indices=gt_idx[:n_gt], gt_node_corr_overlaps=gt_ov[:n_gt], gt_tgt_node_occ=t_occ, gt_src_node_occ=s_occ,
        src_node_corr_indices=s_ci, tgt_node_corr_indices=t_ci,
        src_node_corr_knn_points=ops.gather_rows(src_pts, s_cki, pad_row=Ns),
        tgt_node_corr_knn_points=ops.gather_rows(tgt_pts, t_cki, pad_row=Nt),
        src_node_corr_knn_masks=s_km.bool()[s_ci], tgt_node_corr_knn_masks=t_km.bool()[t_ci],
        matching_scores=scores[:P], tgt_corr_points=t_cp[:n_c], src_corr_points=s_cp[:n_c], corr_scores=c_sc[:n_c])
    if aux is not None:
        aux.update(node_corr_scores=node_sc[:P], src_node_knn_indices=s_ki, tgt_node_knn_indices=t_ki,
                   src_node_masks=s_nm.bool(), tgt_node_masks=t_nm.bool(), src_node_knn_masks=s_km.bool(),
                   tgt_node_knn_masks=t_km.bool(), corr_flat=c_flat[:n_c])
    return out
