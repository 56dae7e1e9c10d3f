"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md). Never imported by roitr_b200/.

Plain-PyTorch fp32 restatement of the reference forward hot path (RIGA_v2.forward), written as pure functions of
(state_dict, inputs). It travels to the GPU box (where /root/reference does not exist) and is the checker for
the CUDA path. It deliberately keeps the reference's *unfolded* formulation (explicit (m,k,C) positional tensors,
explicit (N,N,3,C) sinusoid tensors, two einsums then add then divide) so that its rounding behaviour is the
reference's, not the optimised kernels'.

Parity pin: tests/test_oracle_pinned.py compares every stage of this file against (a) the unmodified reference
executed under oracle/reference_shim.py in this container and (b) the committed golden vectors in tests/golden/
that were generated from the reference by tests/golden/make_golden.py. Native kNN/FPS come from
oracle/pointops_ref.c (oracle.native).

Section map (reference file:line each function follows):
  ppf                        lib/utils.py:358-389
  local_ppf_transformer      model/transformer/ppftransformer.py:243-253, attention.py:166-200,308-320,
                             positional_encoding.py:77-79
  transition_down / block    model/model.py:56-80, 28-44, 131-142
  transition_up              model/model.py:99-117, cpp_wrappers/pointops/functions/pointops.py:168-182
  geometric_embedding        model/transformer/positional_encoding.py:9-34,48-62,111-154
  rpe_self_layer             model/transformer/geoattention.py:101-136,216-232,251-261,186-192
  cross_layer                model/transformer/geoattention.py:43-66,154-173,281-292
  backbone                   model/model.py:187-237 (dead all-pairs PPF at :208-212 skipped)
  partition                  lib/utils.py:448-463, 139-156
  node_occlusion / node_correspondences   lib/utils.py:474-527, 530-614
  coarse_matching_3d / _4d   model/modules.py:135-178, 75-132
  optimal_transport          model/modules.py:21-68
  fine_matching              model/modules.py:216-324
  riga_forward               model/RIGA_v2.py:58-175
"""
import math

import torch
import torch.nn.functional as F

from oracle import native

NUM_HEADS = 4
STRIDES = (1, 4, 4, 4)
NSAMPLE = (8, 16, 16, 16)
BLOCKS = (2, 3, 3, 3)


# --------------------------------------------------------------------------------------------- helpers
def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _i32(vals, device=None):
    return torch.tensor(list(vals), dtype=torch.int32, device=device)


def knn(nsample, xyz, new_xyz, offset, new_offset):
    """pointops.knnquery (pointops.py:30-45): returns (idx int32, sqrt(d2))."""
    idx, d2 = native.knn(nsample, xyz.contiguous(), new_xyz.contiguous(), offset, new_offset)
    return idx, torch.sqrt(d2)


def group_indices(nsample, xyz, new_xyz, offset, new_offset):
    """queryandgroup(..., return_idx=True) (pointops.py:87-92): kNN(nsample+1), drop column 0."""
    idx, _ = knn(nsample + 1, xyz, new_xyz, offset, new_offset)
    return idx[:, 1:].contiguous().long()


def ppf(points, normals, patches, patch_normals):
    p = points.unsqueeze(1).expand_as(patches)
    n1 = normals.unsqueeze(1).expand_as(patches)
    d = patches - p

    def angle(a, b):
        y = (a * b).sum(-1, keepdim=True)
        x = torch.cross(a, b, dim=-1)
        x = torch.sqrt((x ** 2).sum(-1, keepdim=True))
        return torch.atan2(x, y) / math.pi

    dist = torch.sqrt((d ** 2).sum(-1, keepdim=True))
    return torch.cat([dist, angle(n1, d), angle(patch_normals, d), angle(n1, patch_normals)], dim=-1)


# --------------------------------------------------------------------------------------------- local attention
def local_ppf_transformer(sd, p, feats, node_idx, group_idx, ppfs):
    """(n,Cin) -> (m,Cout). ``p`` is the prefix of a LocalPPFTransformer (…'.transformer')."""
    H = NUM_HEADS
    e = _lin(sd, p + ".embedding.proj", ppfs)                        # (m,k,C)
    f = _lin(sd, p + ".in_proj", feats)                              # (n,C)
    a = p + ".transformer.attention"
    q, k, v = _lin(sd, a + ".proj_q", f), _lin(sd, a + ".proj_k", f), _lin(sd, a + ".proj_v", f)
    pp, vp = _lin(sd, a + ".proj_p", e), _lin(sd, a + ".proj_vp", e)
    m, kk, C = pp.shape
    c = C // H
    Q = q[node_idx].view(m, H, 1, c)
    K = k[group_idx].view(m, kk, H, c).permute(0, 2, 1, 3)
    V = v[group_idx].view(m, kk, H, c).permute(0, 2, 1, 3)
    P = pp.view(m, kk, H, c).permute(0, 2, 1, 3)
    VP = vp.view(m, kk, H, c).permute(0, 2, 1, 3)
    s_p = torch.einsum("bhnc,bhmc->bhnm", Q, P)
    s_e = torch.einsum("bhnc,bhmc->bhnm", Q, K)
    s = (s_e + s_p) / c ** 0.5
    w = F.softmax(s, dim=-1)
    h = torch.matmul(w, V + VP)                                      # (m,H,1,c)
    h = h.permute(0, 2, 1, 3).reshape(m, C)
    h = _lin(sd, p + ".transformer.linear", h)
    y = _ln(sd, p + ".transformer.norm", h + f[node_idx])
    return _lin(sd, p + ".out_proj", y)


def transition_down(sd, p, stride, nsample, pts, x, o, nrm):
    """-> (new_pts, new_x, new_o, new_nrm, fps_idx(long), aux)."""
    if stride != 1:
        ends = o.tolist()
        cnt, new_o = 0, []
        for s, e in zip([0] + ends[:-1], ends):
            cnt += (e - s) // stride
            new_o.append(cnt)
        new_o = _i32(new_o, pts.device)
        idx = native.fps(pts.contiguous(), o, new_o).long()
        n_p, n_n = pts[idx], nrm[idx]
    else:
        new_o, n_p, n_n = o, pts, nrm
        idx = torch.arange(pts.shape[0], device=pts.device)
    g = group_indices(nsample, pts, n_p.contiguous(), o, new_o)
    pf = ppf(n_p, n_n, pts[g], nrm[g])
    x = local_ppf_transformer(sd, p + ".transformer", x, idx, g, pf)
    return n_p.contiguous(), x, new_o, n_n.contiguous(), idx, dict(group_idx=g, ppf=pf)


def block(sd, p, nsample, pts, x, o, nrm, idx, ppf_r):
    """RIPointTransformerBlock: returns (x, idx, ppf)."""
    if idx is None:
        idx = group_indices(nsample, pts, pts, o, o)
    if ppf_r is None:
        ppf_r = ppf(pts, nrm, pts[idx], nrm[idx])
    node_idx = torch.arange(pts.shape[0], device=pts.device)
    y = local_ppf_transformer(sd, p + ".transformer.transformer", x, node_idx, idx, ppf_r)
    y = _ln(sd, p + ".bn2", y)
    return F.relu(y + x), idx, ppf_r


def interpolate3(coarse_xyz, fine_xyz, feat, o_coarse, o_fine, k=3):
    idx, dist = knn(k, coarse_xyz, fine_xyz, o_coarse, o_fine)
    r = 1.0 / (dist + 1e-8)
    w = r / r.sum(dim=1, keepdim=True)
    out = torch.zeros(fine_xyz.shape[0], feat.shape[1], device=feat.device)
    for i in range(k):
        out += feat[idx[:, i].long()] * w[:, i].unsqueeze(-1)
    return out


def transition_up_head(sd, p, x, o):
    outs = []
    ends = o.tolist()
    for s, e in zip([0] + ends[:-1], ends):
        xb = x[s:e]
        g = F.relu(_lin(sd, p + ".linear2.0", xb.sum(0, True) / (e - s)))
        outs.append(torch.cat([xb, g.repeat(e - s, 1)], 1))
    y = torch.cat(outs, 0)
    return F.relu(_ln(sd, p + ".linear1.1", _lin(sd, p + ".linear1.0", y)))


def transition_up(sd, p, p1, x1, o1, p2, x2, o2):
    a = F.relu(_ln(sd, p + ".linear1.1", _lin(sd, p + ".linear1.0", x1)))
    b = F.relu(_ln(sd, p + ".linear2.1", _lin(sd, p + ".linear2.0", x2)))
    return a + interpolate3(p2, p1, b, o2, o1)


# --------------------------------------------------------------------------------------------- global transformer
def sinusoid(sd_div_term, t, d_model):
    om = t.reshape(-1, 1, 1) * sd_div_term.view(1, -1, 1)
    emb = torch.cat([torch.sin(om), torch.cos(om)], dim=2)
    return emb.view(*t.shape, d_model)


def geometric_indices(points, sigma_d=0.2, sigma_a=15.0, angle_k=3):
    """points (B,N,3) -> d_indices (B,N,N), a_indices (B,N,N,k), knn (B,N,k)."""
    xy = torch.matmul(points, points.transpose(-1, -2))
    x2 = (points ** 2).sum(-1).unsqueeze(-1)
    y2 = (points ** 2).sum(-1).unsqueeze(-2)
    dist = torch.sqrt((x2 - 2 * xy + y2).clamp(min=0.0))
    d_idx = dist / sigma_d
    B, N, _ = points.shape
    nn_idx = dist.topk(k=angle_k + 1, dim=2, largest=False)[1][:, :, 1:]
    nn_pts = torch.gather(points.unsqueeze(1).expand(B, N, N, 3), 2, nn_idx.unsqueeze(3).expand(B, N, angle_k, 3))
    ref = (nn_pts - points.unsqueeze(2)).unsqueeze(2).expand(B, N, N, angle_k, 3)
    anc = (points.unsqueeze(1) - points.unsqueeze(2)).unsqueeze(3).expand(B, N, N, angle_k, 3)
    sin_v = torch.linalg.norm(torch.cross(ref, anc, dim=-1), dim=-1)
    cos_v = (ref * anc).sum(-1)
    a_idx = torch.atan2(sin_v, cos_v) * (180.0 / (sigma_a * math.pi))
    return d_idx, a_idx, nn_idx


def geometric_embedding(sd, p, points):
    C = sd[p + ".proj_d.weight"].shape[0]
    div = sd[p + ".embedding.div_term"]
    d_idx, a_idx, _ = geometric_indices(points)
    d_emb = _lin(sd, p + ".proj_d", sinusoid(div, d_idx, C))
    a_emb = _lin(sd, p + ".proj_a", sinusoid(div, a_idx, C)).max(dim=3)[0]
    return d_emb + a_emb


def _heads(x, H):
    b, n, C = x.shape
    return x.view(b, n, H, C // H).permute(0, 2, 1, 3)


def _ffn(sd, p, x):
    h = _lin(sd, p + ".squeeze", F.relu(_lin(sd, p + ".expand", x)))
    return _ln(sd, p + ".norm", x + h)


def rpe_self_layer(sd, p, x, emb):
    """x (1,N,C), emb (1,N,N,C) -> (feats, pos)."""
    H = NUM_HEADS
    a = p + ".attention.attention"
    q, k, v = (_heads(_lin(sd, a + ".proj_" + t, x), H) for t in "qkv")
    b, N, C = x.shape
    c = C // H
    pe = _lin(sd, a + ".proj_p", emb).view(b, N, N, H, c).permute(0, 3, 1, 2, 4)
    vpe = _lin(sd, a + ".proj_vp", emb).view(b, N, N, H, c).permute(0, 3, 1, 2, 4)
    s_p = torch.einsum("bhnc,bhnmc->bhnm", q, pe)
    s_e = torch.einsum("bhnc,bhmc->bhnm", q, k)
    s = (s_e + s_p) / c ** 0.5
    eye = torch.eye(N, dtype=torch.bool, device=x.device).view(1, 1, N, N)
    s_noself = s.masked_fill(eye, float("-inf"))
    w = F.softmax(s, dim=-1)
    h = torch.matmul(w, v).permute(0, 2, 1, 3).reshape(b, N, C)
    w2 = F.softmax(s_noself, dim=-1)
    pos = (w2.unsqueeze(-1) * vpe).sum(dim=-2).permute(0, 2, 1, 3).reshape(b, N, C)
    y = _ln(sd, p + ".attention.norm", _lin(sd, p + ".attention.linear", h) + x)
    pos = _ln(sd, p + ".attention.pos_norm", _lin(sd, p + ".attention.pos_linear", pos))
    return _ffn(sd, p + ".output", y), _ffn(sd, p + ".pos_proj", pos)


def cross_layer(sd, p, x, y, pos_x, pos_y):
    H = NUM_HEADS
    a = p + ".attention.attention"
    q = _heads(_lin(sd, a + ".proj_q", x + pos_x), H)
    k = _heads(_lin(sd, a + ".proj_k", y + pos_y), H)
    v = _heads(_lin(sd, a + ".proj_v", y), H)
    c = q.shape[-1]
    w = F.softmax(torch.einsum("bhnc,bhmc->bhnm", q, k) / c ** 0.5, dim=-1)
    h = torch.matmul(w, v).permute(0, 2, 1, 3).reshape(x.shape)
    z = _ln(sd, p + ".attention.norm", _lin(sd, p + ".attention.linear", h) + x)
    return _ffn(sd, p + ".output", z)


def geometric_transformer(sd, p, pts0, pts1, f0, f1, architecture):
    """Note the reference calls it as (s_p4, t_p4, s_x4, t_x4) (model/model.py:214): '0' = src, '1' = tgt."""
    e0 = geometric_embedding(sd, p + ".embedding", pts0)
    e1 = geometric_embedding(sd, p + ".embedding", pts1)
    f0, f1 = _lin(sd, p + ".in_proj", f0), _lin(sd, p + ".in_proj", f1)
    pos0 = pos1 = None
    for i, kind in enumerate(architecture):
        lp = "%s.transformer.layers.%d" % (p, i)
        if kind == "self":
            f0, pos0 = rpe_self_layer(sd, lp, f0, e0)
            f1, pos1 = rpe_self_layer(sd, lp, f1, e1)
        else:
            f0 = cross_layer(sd, lp, f0, f1, pos0, pos1)
            f1 = cross_layer(sd, lp, f1, f0, pos1, pos0)
    return _lin(sd, p + ".out_proj", f0), _lin(sd, p + ".out_proj", f1), dict(emb0=e0, emb1=e1)


# --------------------------------------------------------------------------------------------- backbone
def encode(sd, pts, x, o, nrm, trace):
    """One cloud through enc1..enc4. Returns per-level dicts."""
    levels = []
    for li in range(4):
        p = "backbone.enc%d" % (li + 1)
        pts, x, o, nrm, fps_idx, aux = transition_down(sd, p + ".0", STRIDES[li], NSAMPLE[li], pts, x, o, nrm)
        td_x = x
        idx = ppf_r = None
        for bi in range(1, BLOCKS[li]):
            x, idx, ppf_r = block(sd, "%s.%d" % (p, bi), NSAMPLE[li], pts, x, o, nrm, idx, ppf_r)
        levels.append(dict(p=pts, x=x, o=o, n=nrm, idx=idx, ppf=ppf_r, down_idx=fps_idx, td_x=td_x,
                           td_group_idx=aux["group_idx"], td_ppf=aux["ppf"]))
    return levels


def decode(sd, L):
    x4 = transition_up_head(sd, "backbone.dec4.0", L[3]["x"], L[3]["o"])
    x4, _, _ = block(sd, "backbone.dec4.1", NSAMPLE[3], L[3]["p"], x4, L[3]["o"], L[3]["n"], L[3]["idx"], L[3]["ppf"])
    xs = [None, None, None, x4]
    for li in (2, 1, 0):
        p = "backbone.dec%d" % (li + 1)
        y = transition_up(sd, p + ".0", L[li]["p"], L[li]["x"], L[li]["o"], L[li + 1]["p"], xs[li + 1], L[li + 1]["o"])
        y, _, _ = block(sd, p + ".1", NSAMPLE[li], L[li]["p"], y, L[li]["o"], L[li]["n"], L[li]["idx"], L[li]["ppf"])
        xs[li] = y
    return xs


def backbone(sd, s_pxon, t_pxon, src_deformed, architecture):
    S = encode(sd, *s_pxon, trace=None)
    T = encode(sd, *t_pxon, trace=None)
    s_g, t_g, gaux = geometric_transformer(sd, "backbone.global_transformer", S[3]["p"].unsqueeze(0),
                                           T[3]["p"].unsqueeze(0), S[3]["x"].unsqueeze(0), T[3]["x"].unsqueeze(0),
                                           architecture)
    s_dec, t_dec = decode(sd, S), decode(sd, T)
    d3 = S[1]["down_idx"][S[2]["down_idx"]]
    d4 = d3[S[3]["down_idx"]]
    s_nodes = src_deformed[d4]
    aux = dict(src_levels=S, tgt_levels=T, src_node_idx=d4, src_dec=s_dec, tgt_dec=t_dec, **gaux)
    return s_nodes, s_g[0], src_deformed, s_dec[0], T[3]["p"], t_g[0], T[0]["p"], t_dec[0], aux


# --------------------------------------------------------------------------------------------- matching head
def square_distance(a, b, normalized=False):
    if normalized:
        d = 2.0 - 2.0 * torch.matmul(a, b.transpose(-1, -2).contiguous())
    else:
        d = -2.0 * torch.matmul(a, b.transpose(-1, -2).contiguous())
        d = d + (a ** 2).sum(-1).unsqueeze(-1)
        d = d + (b ** 2).sum(-1).unsqueeze(-2)
    return d.clamp(min=1e-12)


def partition(points, nodes, point_limit):
    d = square_distance(nodes[None], points[None])[0]               # (M,N)
    owner = d.min(dim=0)[1]
    M, N = d.shape
    node_masks = torch.zeros(M, dtype=torch.bool, device=d.device)
    node_masks[owner] = True
    own = torch.zeros_like(d, dtype=torch.bool)
    own[owner, torch.arange(N, device=d.device)] = True
    d = d.masked_fill(~own, 1e12)
    knn_idx = d.topk(k=point_limit, dim=1, largest=False)[1]
    knn_masks = owner[knn_idx] == torch.arange(M, device=d.device).unsqueeze(1)
    knn_idx = knn_idx.masked_fill(~knn_masks, N)
    return owner, node_masks, knn_idx, knn_masks


def node_occlusion(ref_knn_ids, src_knn_ids, ref_pts, src_pts, rot, trans, ref_masks, src_masks, ref_knn_masks,
                   src_knn_masks, thres=0.0375):
    src_pts = torch.matmul(src_pts, rot.T) + trans.T
    ro, so = _i32([ref_pts.shape[0]], ref_pts.device), _i32([src_pts.shape[0]], ref_pts.device)
    _, rd = knn(1, src_pts, ref_pts, so, ro)
    _, sdist = knn(1, ref_pts, src_pts, ro, so)
    r_ov = (rd < thres).float().squeeze(1)
    s_ov = (sdist < thres).float().squeeze(1)
    r = (r_ov[ref_knn_ids] * ref_knn_masks).sum(1) / (ref_knn_masks.sum(1) + 1e-10)
    s = (s_ov[src_knn_ids] * src_knn_masks).sum(1) / (src_knn_masks.sum(1) + 1e-10)
    return r * ref_masks, s * src_masks


def node_correspondences(ref_nodes, src_nodes, ref_knn_pts, src_knn_pts, rot, trans, radius, ref_masks, src_masks,
                         ref_knn_masks, src_knn_masks):
    src_nodes = torch.matmul(src_nodes, rot.T) + trans.T
    src_knn_pts = torch.matmul(src_knn_pts, rot.T) + trans.T[None]
    pair_ok = ref_masks.unsqueeze(1) & src_masks.unsqueeze(0)
    rd = torch.linalg.norm(ref_knn_pts - ref_nodes.unsqueeze(1), dim=-1).masked_fill(~ref_knn_masks, 0.0).max(1)[0]
    sdm = torch.linalg.norm(src_knn_pts - src_nodes.unsqueeze(1), dim=-1).masked_fill(~src_knn_masks, 0.0).max(1)[0]
    dm = torch.sqrt(square_distance(ref_nodes[None], src_nodes[None])[0])
    hit = ((rd.unsqueeze(1) + sdm.unsqueeze(0) + radius - dm) > 0) & pair_ok
    ri, si = torch.nonzero(hit, as_tuple=True)
    rkm, skm = ref_knn_masks[ri], src_knn_masks[si]
    pd = square_distance(ref_knn_pts[ri], src_knn_pts[si])
    pd = pd.masked_fill(~(rkm.unsqueeze(2) & skm.unsqueeze(1)), 1e12)
    close = pd < radius ** 2
    r_cnt = torch.count_nonzero(close.sum(-1), dim=-1).float()
    s_cnt = torch.count_nonzero(close.sum(-2), dim=-1).float()
    ov = (r_cnt / rkm.sum(-1).float() + s_cnt / skm.sum(-1).float()) / 2
    keep = ov > 0
    return torch.stack([ri[keep], si[keep]], dim=1), ov[keep]


def coarse_scores_3d(ref_feats, src_feats):
    s = torch.exp(-square_distance(ref_feats[None], src_feats[None]))[0]
    return (s / (s.sum(dim=1, keepdim=True) + 1e-8)) * (s / (s.sum(dim=0, keepdim=True) + 1e-8))


def coarse_matching_3d(ref_feats, src_feats, ref_masks, src_masks, num):
    ri = torch.nonzero(ref_masks, as_tuple=True)[0]
    si = torch.nonzero(src_masks, as_tuple=True)[0]
    s = coarse_scores_3d(ref_feats[ri], src_feats[si])
    k = min(num, s.numel())
    sc, flat = s.view(-1).topk(k=k, largest=True)
    return ri[flat // s.shape[1]], si[flat % s.shape[1]], sc


def coarse_matching_4d(a_feats, b_feats, a_masks, b_masks, min_num, thr=0.75):
    ai = torch.nonzero(a_masks, as_tuple=True)[0]
    bi = torch.nonzero(b_masks, as_tuple=True)[0]
    sim = torch.sqrt(square_distance(a_feats[ai][None], b_feats[bi][None], normalized=True)[0])
    k = min(min_num, sim.numel())
    ok = sim <= thr
    if ok.sum() < k:
        dist, flat = sim.view(-1).topk(k=k, largest=False)
        ra, rb = torch.div(flat, sim.shape[1], rounding_mode="floor"), flat % sim.shape[1]
    else:
        ra, rb = torch.nonzero(ok, as_tuple=True)
        dist = sim[ra, rb]
    return ai[ra], bi[rb], torch.exp(-dist)


def optimal_transport(alpha, scores, row_masks, col_masks, num_iter=100, inf=1e6):
    B, R, Cn = scores.shape
    dev = scores.device
    prm = torch.zeros(B, R + 1, dtype=torch.bool, device=dev)
    prm[:, :R] = ~row_masks
    pcm = torch.zeros(B, Cn + 1, dtype=torch.bool, device=dev)
    pcm[:, :Cn] = ~col_masks
    z = torch.cat([torch.cat([scores, alpha.expand(B, R, 1)], dim=-1), alpha.expand(B, 1, Cn + 1)], dim=1)
    z = z.masked_fill(prm.unsqueeze(2) | pcm.unsqueeze(1), -inf)
    nr, nc = row_masks.to(scores.dtype).sum(1), col_masks.to(scores.dtype).sum(1)
    norm = -torch.log(nr + nc)
    log_mu = torch.empty(B, R + 1, dtype=scores.dtype, device=dev)
    log_mu[:, :R] = norm.unsqueeze(1)
    log_mu[:, R] = torch.log(nc) + norm
    log_mu[prm] = -inf
    log_nu = torch.empty(B, Cn + 1, dtype=scores.dtype, device=dev)
    log_nu[:, :Cn] = norm.unsqueeze(1)
    log_nu[:, Cn] = torch.log(nr) + norm
    log_nu[pcm] = -inf
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(num_iter):
        u = log_mu - torch.logsumexp(z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(z + u.unsqueeze(2), dim=1)
    return z + u.unsqueeze(2) + v.unsqueeze(1) - norm.view(B, 1, 1)


def fine_matching(ref_pts, src_pts, ref_masks, src_masks, log_scores, k, thr=0.05, mutual=True):
    s = torch.exp(log_scores)
    B, R, Cn = s.shape
    rv, ri = s.topk(k=k, dim=2)
    rmat = torch.zeros_like(s).scatter_(2, ri, rv) > thr
    cv, ci = s.topk(k=k, dim=1)
    cmat = torch.zeros_like(s).scatter_(1, ci, cv) > thr
    corr = (rmat & cmat) if mutual else (rmat | cmat)
    corr = corr & (ref_masks.unsqueeze(2) & src_masks.unsqueeze(1))
    b, r, c = torch.nonzero(corr, as_tuple=True)
    return ref_pts[b, r], src_pts[b, c], (s * corr.float())[b, r, c], torch.stack([b, r, c], 1)


def fine_stage_fp64(alpha, tgt_pf, src_pf, t_ki, s_ki, t_km, s_km, t_ci, s_ci, num_iter=100):
    """The fine stage of RIGA_v2.forward (RIGA_v2.py:139-160: gather patch descriptors, einsum / sqrt(C), log optimal
    transport) evaluated in float64 from fp32 descriptors: the yardstick for how far TWO correct fp32 evaluations of the
    same formula may sit from exact arithmetic (tests/parity.py). Returns the (P, 65, 65) log-assignment in float64."""
    tgt_pf, src_pf = tgt_pf.double(), src_pf.double()
    s_pad = torch.cat([src_pf, torch.zeros_like(src_pf[:1])], 0)
    t_pad = torch.cat([tgt_pf, torch.zeros_like(tgt_pf[:1])], 0)
    s_f, t_f = s_pad[s_ki[s_ci]], t_pad[t_ki[t_ci]]
    ms = torch.einsum("bnd,bmd->bnm", t_f, s_f) / src_pf.shape[1] ** 0.5
    return optimal_transport(alpha.double(), ms, t_km[t_ci], s_km[s_ci], num_iter)


# --------------------------------------------------------------------------------------------- pipeline
def riga_forward(sd, cfg, src_pcd, tgt_pcd, src_feats, tgt_feats, src_normals, tgt_normals, rot, trans, src_raw_pcd,
                 with_aux=False):
    """Restates RIGA_v2.forward (eval mode). ``cfg`` needs: benchmark, num_est_coarse_corr,
    transformer_architecture, point_per_patch, matching_radius, fine_matching_topk, fine_matching_mutual,
    fine_matching_confidence_threshold."""
    four_d = cfg["benchmark"] not in ("3DMatch", "3DLoMatch")
    so, to = _i32([src_raw_pcd.shape[0]], src_raw_pcd.device), _i32([tgt_pcd.shape[0]], tgt_pcd.device)
    (src_nodes, src_nf, src_pts, src_pf, tgt_nodes, tgt_nf, tgt_pts, tgt_pf, aux) = backbone(
        sd, [src_raw_pcd, src_feats, so, src_normals], [tgt_pcd, tgt_feats, to, tgt_normals], src_pcd,
        cfg["transformer_architecture"])
    src_nf = F.normalize(_lin(sd, "coarse_proj", src_nf), p=2, dim=1)
    tgt_nf = F.normalize(_lin(sd, "coarse_proj", tgt_nf), p=2, dim=1)
    src_pf, tgt_pf = _lin(sd, "fine_proj", src_pf), _lin(sd, "fine_proj", tgt_pf)
    out = dict(src_points=src_pts, tgt_points=tgt_pts, src_nodes=src_nodes, tgt_nodes=tgt_nodes,
               src_point_feats=src_pf, tgt_point_feats=tgt_pf, src_node_feats=src_nf, tgt_node_feats=tgt_nf)
    K = cfg["point_per_patch"]
    _, s_nm, s_ki, s_km = partition(src_pts, src_nodes, K)
    _, t_nm, t_ki, t_km = partition(tgt_pts, tgt_nodes, K)
    s_pad = torch.cat([src_pts, torch.zeros_like(src_pts[:1])], 0)
    t_pad = torch.cat([tgt_pts, torch.zeros_like(tgt_pts[:1])], 0)
    s_kp, t_kp = s_pad[s_ki], t_pad[t_ki]
    gt_idx, gt_ov = node_correspondences(tgt_nodes, src_nodes, t_kp, s_kp, rot, trans, cfg["matching_radius"], t_nm,
                                         s_nm, t_km, s_km)
    t_occ, s_occ = node_occlusion(t_ki, s_ki, t_pad, s_pad, rot, trans, t_nm, s_nm, t_km, s_km)
    out.update(gt_node_corr_indices=gt_idx, gt_node_corr_overlaps=gt_ov, gt_tgt_node_occ=t_occ, gt_src_node_occ=s_occ)
    if four_d:
        t_ci, s_ci, node_sc = coarse_matching_4d(tgt_nf, src_nf, t_nm, s_nm, cfg["num_est_coarse_corr"])
    else:
        t_ci, s_ci, node_sc = coarse_matching_3d(tgt_nf, src_nf, t_nm, s_nm, cfg["num_est_coarse_corr"])
    out.update(src_node_corr_indices=s_ci, tgt_node_corr_indices=t_ci)
    s_cki, t_cki = s_ki[s_ci], t_ki[t_ci]
    s_ckm, t_ckm = s_km[s_ci], t_km[t_ci]
    s_ckp, t_ckp = s_kp[s_ci], t_kp[t_ci]
    s_pf_pad = torch.cat([src_pf, torch.zeros_like(src_pf[:1])], 0)
    t_pf_pad = torch.cat([tgt_pf, torch.zeros_like(tgt_pf[:1])], 0)
    s_f, t_f = s_pf_pad[s_cki], t_pf_pad[t_cki]
    out.update(src_node_corr_knn_points=s_ckp, tgt_node_corr_knn_points=t_ckp, src_node_corr_knn_masks=s_ckm,
               tgt_node_corr_knn_masks=t_ckm)
    ms = torch.einsum("bnd,bmd->bnm", t_f, s_f) / src_pf.shape[1] ** 0.5
    ms = optimal_transport(sd["optimal_transport.alpha"], ms, t_ckm, s_ckm)
    out["matching_scores"] = ms
    t_cp, s_cp, sc, brc = fine_matching(t_ckp, s_ckp, t_ckm, s_ckm, ms[:, :-1, :-1], cfg["fine_matching_topk"],
                                        cfg["fine_matching_confidence_threshold"], cfg["fine_matching_mutual"])
    out.update(tgt_corr_points=t_cp, src_corr_points=s_cp, corr_scores=sc)
    if with_aux:
        aux.update(node_corr_scores=node_sc, corr_brc=brc, src_node_knn_indices=s_ki, tgt_node_knn_indices=t_ki,
                   src_node_masks=s_nm, tgt_node_masks=t_nm, src_node_knn_masks=s_km, tgt_node_knn_masks=t_km)
        out["_aux"] = aux
    return out
