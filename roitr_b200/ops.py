"""Thin typed wrappers over the C ABI (include/roitr_b200.h). torch only allocates the output tensors and names the
stream; every computation happens in libroitr_b200. No fallbacks: errors raise RoitrError."""
import ctypes

import torch

from . import _lib
from ._lib import c_float, c_int, f32, i32, ptr, stream_ptr

c_ll = ctypes.c_longlong
MODE_LN, MODE_RELU, MODE_L2NORM = 1, 2, 4


def _u8(t):
    return ptr(t, torch.uint8)


def empty(*shape, dtype=torch.float32, like=None, device=None):
    return torch.empty(*shape, dtype=dtype, device=like.device if like is not None else device)


def linear(a, w, bias=None, relu=False, a_index=None, a_add=None, out=None, M=None, K=None, lda=None, ldw=None,
           ldc=None, wpack=None):
    """out[M,N] = (a [+ a_add])[rows, :K] @ w[:N, :K]^T + bias. ``a``/``out`` may be column slices of wider buffers
    (pass lda/ldc); ``a_index`` gathers rows of ``a``."""
    N = w.shape[0]
    K = w.shape[1] if K is None else K
    if M is None:
        M = a_index.shape[0] if a_index is not None else a.shape[0]
    lda = a.stride(0) if lda is None else lda
    ldw = w.stride(0) if ldw is None else ldw
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    ldc = out.stride(0) if ldc is None else ldc
    if wpack is not None:       # (packed weight, tile rows) from engine.pack_linear_tc: persistent tcgen05 kernel
        _lib.call("roitr_linear_tc_packed", c_int(M), c_int(N), c_int(K), c_void(a), c_void(a_add), c_int(lda), i32(a_index),
                  f32(wpack[0]), c_int(wpack[1]), c_void(bias), c_void(out), c_int(ldc), c_int(1 if relu else 0), stream_ptr())
        return out
    _lib.call("roitr_linear", c_int(M), c_int(N), c_int(K), c_void(a), c_void(a_add), c_int(lda), i32(a_index),
              c_void(w), c_int(ldw), c_void(bias), c_void(out), c_int(ldc), c_int(1 if relu else 0), stream_ptr())
    return out


def linear_ln(a, w, bias, wpack, gamma, beta, res_pre=None, res_pre_index=None, res_post=None, relu=False):
    """act(LN(a @ w^T + bias + res_pre[res_pre_index]) * gamma + beta + res_post) in ONE kernel when the layer fits the fused
    epilogue (N a multiple of 32 within one weight tile - or two 128-row tiles, N <= 256 -, plain aligned input); otherwise
    linear + row_epilogue."""
    M, K = a.shape
    N = w.shape[0]
    fused = (wpack is not None and N % 32 == 0 and (N <= wpack[1] or (wpack[1] == 128 and N <= 256)) and a.stride(0) % 4 == 0 and K % 4 == 0 and
             a.data_ptr() % 16 == 0 and a.stride(1) == 1 and
             (res_pre is None or res_post is None or res_pre.stride(0) == res_post.stride(0)))      # one residual pitch
    if not fused:
        t = linear(a, w, bias, wpack=wpack)
        res_pre = res_pre.contiguous() if res_pre is not None else None
        res_post = res_post.contiguous() if res_post is not None else None
        return row_epilogue(t, res_pre=res_pre, res_pre_index=res_pre_index, gamma=gamma, beta=beta, res_post=res_post,
                            mode=MODE_LN | (MODE_RELU if relu else 0))
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    ldr = (res_pre if res_pre is not None else res_post if res_post is not None else out).stride(0)
    _lib.call("roitr_linear_ln_tc_packed", c_int(M), c_int(N), c_int(K), c_void(a), c_int(a.stride(0)), f32(wpack[0]),
              c_int(wpack[1]), c_void(bias), f32(gamma), f32(beta), c_void(res_pre), i32(res_pre_index), c_void(res_post), c_int(ldr),
              c_int(1 if relu else 0), f32(out), c_int(N), stream_ptr())
    return out


def set_linear_variant(v):
    """Configuration of the streaming dense-layer kernel for the launches issued from now on (baked into a graph at capture):
    0 = deep rings, one CTA per SM; 3 = light footprint that shares an SM with other kernels' CTAs."""
    _lib.lib().roitr_set_linear_config(c_int(int(v)))


def c_void(t):
    """Device pointer of a (possibly strided-view) f32 tensor: views into wider buffers are allowed here because the
    leading dimension is passed explicitly."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.dtype == torch.float32
    _lib._note(t)
    return ctypes.c_void_p(t.data_ptr())


def row_epilogue(x, res_pre=None, res_pre_index=None, gamma=None, beta=None, res_post=None, mode=0, out=None):
    M, C = x.shape
    if out is None:
        out = torch.empty_like(x)
    _lib.call("roitr_row_epilogue", c_int(M), c_int(C), f32(x), f32(res_pre), i32(res_pre_index), f32(gamma), f32(beta),
              f32(res_post), f32(out), c_int(mode), stream_ptr())
    return out


def segment_mean(x, offset):
    b, C = offset.shape[0], x.shape[1]
    out = torch.empty(b, C, dtype=torch.float32, device=x.device)
    _lib.call("roitr_segment_mean", c_int(b), c_int(C), f32(x), i32(offset), f32(out), stream_ptr())
    return out


def concat_segment(x, g, offset):
    M, C = x.shape
    out = torch.empty(M, 2 * C, dtype=torch.float32, device=x.device)
    _lib.call("roitr_concat_segment", c_int(M), c_int(C), c_int(offset.shape[0]), f32(x), f32(g), i32(offset), f32(out),
              stream_ptr())
    return out


GRID_MIN_SEGMENT = 1024     # reference sets with at least this many points per segment get the grid-accelerated kNN


def knn_grid_build(xyz, offset, target=None):
    """Uniform-grid acceleration structure over a (segmented) reference set; reusable by every query against it.
    ``target`` = average points per cell over the bounding box (None: the library default, 0.5)."""
    b, n = offset.shape[0], xyz.shape[0]
    fn = _lib.lib().roitr_knn_grid_workspace_bytes
    fn.restype = c_ll
    ws = torch.empty(int(fn(c_int(b), c_int(n))), dtype=torch.uint8, device=xyz.device)
    if target is None:
        _lib.call("roitr_knn_grid_build", c_int(b), c_int(n), f32(xyz), i32(offset), ptr(ws), stream_ptr())
    else:
        _lib.call("roitr_knn_grid_build_target", c_int(b), c_int(n), f32(xyz), i32(offset), c_float(float(target)), ptr(ws), stream_ptr())
    return ws


def knn_ppf(k, xyz, nrm, new_xyz, new_nrm, offset, new_offset, drop_first=1, want_ppf=True, want_dist=False, grid=None,
            qgrid=None):
    """Exact kNN (+PPF). ``grid``: uniform grid over the reference set (knn_grid_build); ``qgrid``: the query set's own
    grid, used only as a cache-friendly visiting order of the queries (same results with or without it)."""
    m = new_xyz.shape[0]
    idx = torch.empty(m, k, dtype=torch.int32, device=xyz.device)
    ppf = torch.empty(m, k, 4, dtype=torch.float32, device=xyz.device) if want_ppf else None
    dist = torch.empty(m, k, dtype=torch.float32, device=xyz.device) if want_dist else None
    if grid is not None:
        _lib.call("roitr_knn_ppf_grid_q", c_int(offset.shape[0]), c_int(m), c_int(k), c_int(drop_first), c_int(xyz.shape[0]),
                  f32(xyz), f32(nrm) if want_ppf else None, f32(new_xyz), f32(new_nrm) if want_ppf else None, i32(offset),
                  i32(new_offset), ptr(grid), ptr(qgrid), i32(idx), f32(dist), f32(ppf), stream_ptr())
        return idx, ppf, dist
    _lib.call("roitr_knn_ppf_n", c_int(offset.shape[0]), c_int(m), c_int(k), c_int(drop_first), c_int(xyz.shape[0]),
              f32(xyz), f32(nrm) if want_ppf else None, f32(new_xyz), f32(new_nrm) if want_ppf else None, i32(offset),
              i32(new_offset), i32(idx), f32(dist), f32(ppf), stream_ptr())
    return idx, ppf, dist


def fps(xyz, offset, new_offset, n_seg_max, m_total, per_segment_rule=True, cluster=0):
    idx = torch.empty(m_total, dtype=torch.int32, device=xyz.device)
    new_xyz = torch.empty(m_total, 3, dtype=torch.float32, device=xyz.device)
    _lib.call("roitr_furthestsampling_cfg", c_int(offset.shape[0]), c_int(0 if per_segment_rule else n_seg_max),
              c_int(n_seg_max), f32(xyz), i32(offset), i32(new_offset), i32(idx), f32(new_xyz), c_int(cluster),
              stream_ptr())
    return idx, new_xyz


def gather_rows(src, index, pad_row=-1):
    c = src.shape[1]
    out = torch.empty(*index.shape, c, dtype=torch.float32, device=src.device)
    _lib.call("roitr_gather_rows", c_ll(index.numel()), c_int(c), ptr(index), c_int(1 if index.dtype == torch.int64 else 0),
              f32(src), f32(out), c_ll(pad_row), stream_ptr())
    return out


def interpolate(idx, dist, feat, base=None):
    n, k = idx.shape
    c = feat.shape[1]
    out = torch.empty(n, c, dtype=torch.float32, device=feat.device)
    _lib.call("roitr_interpolate", c_int(n), c_int(c), c_int(k), i32(idx), f32(dist), f32(feat), f32(base), f32(out),
              stream_ptr())
    return out


def grid_order_ptr(grid, b):
    """Device address of the cell-sorted (x, y, z, index) array inside a grid workspace of b segments (None -> NULL)."""
    if grid is None:
        return ctypes.c_void_p(0)
    fn = _lib.lib().roitr_knn_grid_sorted_offset
    fn.restype = c_ll
    return ctypes.c_void_p(grid.data_ptr() + int(fn(c_int(b))))


def local_attention(qkv, C, node_idx, group_idx, ppf, Ap, cp, Avp, cvp, order=None):
    """``order`` = (grid workspace of the QUERY set, number of segments): visit the queries in cell order."""
    m, knb = group_idx.shape
    P = ctypes.c_void_p
    if isinstance(qkv, tuple):      # (q, k, v) as separate (strided) views: q rows are addressed through node_idx, k / v by group_idx
        q, k, v = qkv
        dev = q.device
        qa, ka, va = (P(q.data_ptr()), c_int(q.stride(0))), (P(k.data_ptr()), c_int(k.stride(0))), (P(v.data_ptr()), c_int(v.stride(0)))
    else:
        dev, ld, base = qkv.device, qkv.stride(0), qkv.data_ptr()
        qa, ka, va = (P(base), c_int(ld)), (P(base + 4 * C), c_int(ld)), (P(base + 8 * C), c_int(ld))
    out = torch.empty(m, C, dtype=torch.float32, device=dev)
    _lib.call("roitr_local_attention_ordered", c_int(m), c_int(C), c_int(4), c_int(knb), qa[0], qa[1], ka[0], ka[1], va[0], va[1],
              i32(node_idx), i32(group_idx), f32(ppf), f32(Ap), f32(cp), f32(Avp),
              f32(cvp), grid_order_ptr(*order) if order is not None else P(0), f32(out), stream_ptr())
    return out


def geo_knn(pts, k=3):
    N = pts.shape[0]
    nn = torch.empty(N, k, dtype=torch.int32, device=pts.device)
    _lib.call("roitr_geo_knn", c_int(N), c_int(k), f32(pts), i32(nn), stream_ptr())
    return nn


def geo_knn_batched(batch, N, pts, k=3):
    nn = torch.empty(batch * N, k, dtype=torch.int32, device=pts.device)
    _lib.call("roitr_geo_knn_batched", c_int(batch), c_int(N), c_int(k), f32(pts), i32(nn), stream_ptr())
    return nn


def geo_embedding_tc_batched(batch, N, pts, nn3, wpack, bd, ba, div_term, sigma_d, sigma_a):
    C = bd.shape[0]
    E = torch.empty(batch, N, N, C, dtype=torch.float32, device=pts.device)
    _lib.call("roitr_geo_embedding_tc_batched", c_int(batch), c_int(N), c_int(C), f32(pts), i32(nn3), f32(wpack), f32(bd),
              f32(ba), f32(div_term), c_float(sigma_d), c_float(sigma_a), f32(E), stream_ptr())
    return E


def geo_embedding_table(batch, N, pts, nn3, tables, Wd, bd, Wa, ba, div_term, sigma_d, sigma_a):
    """tables = engine.build_geo_tables(...). E (batch, N, N, C)."""
    C = bd.shape[0]
    ta, td = tables["tab_a"], tables["tab_d"]
    E = torch.empty(batch, N, N, C, dtype=torch.float32, device=pts.device)
    _lib.call("roitr_geo_embedding_table", c_int(batch), c_int(N), c_int(C), f32(pts), i32(nn3), f32(ta), c_int(ta.shape[1]),
              f32(td), c_int(td.shape[1]), c_float(tables["inv_h"]), f32(Wd), f32(bd), f32(Wa), f32(ba), f32(div_term),
              c_float(sigma_d), c_float(sigma_a), f32(E), stream_ptr())
    return E


def geo_embedding_tc(pts, nn3, wpack, bd, ba, div_term, sigma_d, sigma_a, out=None):
    N, C = pts.shape[0], bd.shape[0]
    E = torch.empty(N, N, C, dtype=torch.float32, device=pts.device) if out is None else out
    _lib.call("roitr_geo_embedding_tc", c_int(N), c_int(C), f32(pts), i32(nn3), f32(wpack), f32(bd), f32(ba), f32(div_term),
              c_float(sigma_d), c_float(sigma_a), f32(E), stream_ptr())
    return E


def gemm_tc_batched(outer, inner, M, N, K, A, lda, sA, W, ldw, sW, C, ldc, sC, w_transposed=False):
    """C_oi = A_oi W_oi^T on tcgen05 for outer x inner operand triples; sX = (outer stride, inner stride) in elements.
    A / W / C may be views (pointers + explicit leading dimensions)."""
    _lib.call("roitr_gemm_tc_batched", c_int(outer), c_int(inner), c_int(M), c_int(N), c_int(K), c_void(A), c_int(lda),
              c_ll(sA[0]), c_ll(sA[1]), c_void(W), c_int(ldw), c_ll(sW[0]), c_ll(sW[1]), c_int(1 if w_transposed else 0),
              c_void(C), c_int(ldc), c_ll(sC[0]), c_ll(sC[1]), stream_ptr())
    return C


def attention_tc(batch, N, M, C, q, k, v, heads=4, E=None, gq=None, bp=None):
    """Attention core with Q K^T and P V on the tensor cores. q: rows of `batch` clouds of N queries (views with explicit
    leading dimension), k / v: `batch` clouds of M keys. E (batch,N,N,C), gq (batch*N,heads,C), bp (C) select the RPE
    self-attention. Returns hidden (batch*N, C) [, G (batch*N, heads, C)]."""
    dev = q.device
    c = C // heads
    ldq, ldk, ldv = q.stride(0), k.stride(0), v.stride(0)
    qk = torch.empty(batch, heads, N, M, dtype=torch.float32, device=dev)
    gemm_tc_batched(batch, heads, N, M, c, q, ldq, (N * ldq, c), k, ldk, (M * ldk, c), qk, M, (heads * N * M, N * M))
    G = None
    if E is not None:
        P = torch.empty_like(qk)
        G = torch.empty(batch * N, heads, C, dtype=torch.float32, device=dev)
        _lib.call("roitr_geo_self_scores_ld", c_int(batch), c_int(N), c_int(C), c_int(heads), f32(qk), c_void(q), c_int(ldq),
                  c_ll(N * ldq), f32(E), c_void(gq), c_int(gq.stride(0)), f32(bp), f32(P), f32(G), stream_ptr())
    else:
        P = qk
        _lib.call("roitr_softmax_rows", c_ll(batch * heads * N), c_int(M), f32(qk), c_float(float(c) ** 0.5), f32(P),
                  stream_ptr())
    hidden = torch.empty(batch * N, C, dtype=torch.float32, device=dev)
    gemm_tc_batched(batch, heads, N, c, M, P, M, (heads * N * M, N * M), v, ldv, (M * ldv, c), hidden, C, (N * C, c),
                    w_transposed=True)
    return (hidden, G) if E is not None else hidden


# ---- matching head: every op takes B equally sized problems laid back to back and issues ONE launch per kernel ----
def point_to_node_batched(B, pts, nodes, limit):
    """pts (B*N,3), nodes (B*M,3) -> owner (B,N), node_mask (B,M) u8, knn_idx (B,M,limit) int32 (pad = N), knn_mask (B,M,limit) u8."""
    N, M = pts.shape[0] // B, nodes.shape[0] // B
    dev = pts.device
    owner = torch.empty(B, N, dtype=torch.int32, device=dev)
    dmin = torch.empty(B, N, dtype=torch.float32, device=dev)
    count = torch.empty(B, M, dtype=torch.int32, device=dev)
    knn_idx = torch.empty(B, M, limit, dtype=torch.int32, device=dev)
    knn_mask = torch.empty(B, M, limit, dtype=torch.uint8, device=dev)
    node_mask = torch.empty(B, M, dtype=torch.uint8, device=dev)
    fn = _lib.lib().roitr_point_to_node_workspace_bytes
    fn.restype = c_ll
    ws = torch.empty(int(fn(c_int(B), c_int(N), c_int(M))), dtype=torch.uint8, device=dev)
    _lib.call("roitr_point_to_node_batched", c_int(B), c_int(N), c_int(M), c_int(limit), f32(pts), f32(nodes), i32(owner),
              f32(dmin), i32(count), ptr(ws), i32(knn_idx), _u8(knn_mask), _u8(node_mask), stream_ptr())
    return owner, node_mask, knn_idx, knn_mask


def point_to_node(pts, nodes, limit):
    owner, node_mask, knn_idx, knn_mask = point_to_node_batched(1, pts, nodes, limit)
    return owner[0], node_mask[0], knn_idx[0], knn_mask[0]


def compact_flags_batched(B, flags, capacity):
    """flags: B equal segments of bytes. Ascending flat indices (within the segment) of the non-zero bytes, (B, max(capacity,1))
    int32 padded, and the true totals (B,) int32 on the device (torch.nonzero order per segment)."""
    n = flags.numel() // B
    fn = _lib.lib().roitr_compact_scratch_ints
    fn.restype = c_ll
    scratch = torch.empty(B * int(fn(c_ll(n))), dtype=torch.int32, device=flags.device)
    out = torch.empty(B, max(capacity, 1), dtype=torch.int32, device=flags.device)
    count = torch.empty(B, dtype=torch.int32, device=flags.device)
    _lib.call("roitr_compact_flags_batched", c_int(B), c_ll(n), _u8(flags), i32(scratch), i32(out), c_int(capacity), i32(count),
              stream_ptr())
    return out, count


def compact_flags(flags, capacity):
    """Ascending flat indices of the non-zero bytes (torch.nonzero order), padded to ``capacity``; device count."""
    out, count = compact_flags_batched(1, flags, capacity)
    return out[0], count


def coarse_matching_batched(B, ref_feats, src_feats, ref_mask, src_mask, k, dual=True, xy=None):
    """ref_feats (B*Mr,C), src_feats (B*Ms,C), masks (B,Mr) / (B,Ms) u8, xy (B,Mr,Ms) -> (B,k) ref idx, src idx, scores; (B,) counts."""
    Mr, Ms, C = ref_feats.shape[0] // B, src_feats.shape[0] // B, ref_feats.shape[1]
    dev = ref_feats.device
    if xy is None:
        xy = torch.empty(B, Mr, Ms, dtype=torch.float32, device=dev)
        for b in range(B):
            linear(ref_feats[b * Mr:(b + 1) * Mr], src_feats[b * Ms:(b + 1) * Ms], out=xy[b])     # ref @ src^T
    work = torch.empty(B * (Mr * Ms + 2 * (Mr + Ms)), dtype=torch.float32, device=dev)
    out_ref = torch.zeros(B, k, dtype=torch.int32, device=dev)
    out_src = torch.zeros(B, k, dtype=torch.int32, device=dev)
    out_score = torch.zeros(B, k, dtype=torch.float32, device=dev)
    count = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.call("roitr_coarse_matching_batched", c_int(B), c_int(Mr), c_int(Ms), c_int(C), c_int(k), c_int(1 if dual else 0),
              f32(ref_feats), f32(src_feats), _u8(ref_mask), _u8(src_mask), f32(xy), f32(work), i32(out_ref), i32(out_src),
              f32(out_score), i32(count), stream_ptr())
    return out_ref, out_src, out_score, count


def coarse_matching(ref_feats, src_feats, ref_mask, src_mask, k, dual=True, xy=None):
    r, s_, sc, c = coarse_matching_batched(1, ref_feats, src_feats, ref_mask, src_mask, k, dual, None if xy is None else xy[None])
    return r[0], s_[0], sc[0], c


def coarse_matching_adaptive_batched(B, a_feats, b_feats, a_mask, b_mask, min_num, threshold, cap, xy=None):
    Ma, Mb = a_feats.shape[0] // B, b_feats.shape[0] // B
    dev = a_feats.device
    if xy is None:
        xy = torch.empty(B, Ma, Mb, dtype=torch.float32, device=dev)
        for b in range(B):
            linear(a_feats[b * Ma:(b + 1) * Ma], b_feats[b * Mb:(b + 1) * Mb], out=xy[b])
    n = Ma * Mb
    work = torch.empty(B * (2 * n + (n + 3) // 4), dtype=torch.float32, device=dev)
    fn = _lib.lib().roitr_compact_scratch_ints
    fn.restype = c_ll
    iwork = torch.empty(B * (cap + 3 * min_num + 4 + int(fn(c_ll(n)))), dtype=torch.int32, device=dev)
    out_a = torch.zeros(B, cap, dtype=torch.int32, device=dev)
    out_b = torch.zeros(B, cap, dtype=torch.int32, device=dev)
    out_score = torch.zeros(B, cap, dtype=torch.float32, device=dev)
    count = torch.empty(B, dtype=torch.int32, device=dev)
    _lib.call("roitr_coarse_matching_adaptive_batched", c_int(B), c_int(Ma), c_int(Mb), c_int(min_num), c_float(threshold),
              _u8(a_mask), _u8(b_mask), f32(xy), f32(work), i32(iwork), c_int(cap), i32(out_a), i32(out_b), f32(out_score),
              i32(count), stream_ptr())
    return out_a, out_b, out_score, count


def coarse_matching_adaptive(a_feats, b_feats, a_mask, b_mask, min_num, threshold, cap, xy=None):
    a, b, sc, c = coarse_matching_adaptive_batched(1, a_feats, b_feats, a_mask, b_mask, min_num, threshold, cap,
                                                   None if xy is None else xy[None])
    return a[0], b[0], sc[0], c


def fine_matching_batched(B, tgt_feat, src_feat, tgt_knn, src_knn, tgt_kmask, src_kmask, corr_t, corr_s, corr_count, alpha,
                          num_iter, topk, mutual, threshold):
    """tgt_feat (B*Nt,C), src_feat (B*Ns,C), knn / kmask (B,M,64), corr_t / corr_s (B,Pmax), corr_count (B) ->
    scores (B,Pmax,65,65), flags (B,Pmax,64,64) u8."""
    Pmax = corr_t.shape[1]
    Nt, Ns, C = tgt_feat.shape[0] // B, src_feat.shape[0] // B, tgt_feat.shape[1]
    Mt, Ms = tgt_knn.shape[1], src_knn.shape[1]
    dev = tgt_feat.device
    if tgt_knn.shape[2] != 64 or src_knn.shape[2] != 64:
        raise _lib.RoitrError("fine_matching: patches must hold 64 points (got %d / %d)" % (tgt_knn.shape[2], src_knn.shape[2]))
    scores = torch.zeros(B, Pmax, 65, 65, dtype=torch.float32, device=dev)
    flags = torch.empty(B, Pmax, 64, 64, dtype=torch.uint8, device=dev)
    _lib.call("roitr_fine_matching_batched", c_int(B), c_int(Pmax), c_int(Mt), c_int(Ms), c_int(Nt), c_int(Ns), c_int(C),
              f32(tgt_feat), f32(src_feat), i32(tgt_knn), i32(src_knn), _u8(tgt_kmask), _u8(src_kmask), i32(corr_t), i32(corr_s),
              i32(corr_count), f32(alpha), c_int(num_iter), c_int(topk), c_int(1 if mutual else 0), c_float(threshold),
              f32(scores), _u8(flags), stream_ptr())
    return scores, flags


def fine_matching(tgt_feat, src_feat, tgt_knn, src_knn, tgt_kmask, src_kmask, corr_t, corr_s, corr_count, alpha,
                  num_iter, topk, mutual, threshold):
    sc, fl = fine_matching_batched(1, tgt_feat, src_feat, tgt_knn[None], src_knn[None], tgt_kmask[None], src_kmask[None],
                                   corr_t[None], corr_s[None], corr_count.view(1), alpha, num_iter, topk, mutual, threshold)
    return sc[0], fl[0]


def fine_gather_batched(B, capacity, flat, count, scores, corr_t, corr_s, tgt_knn, src_knn, tgt_pts, src_pts):
    """flat (B,capacity), count (B), tgt_pts (B*Nt1,3), src_pts (B*Ns1,3) -> (B,capacity,3) x2 and (B,capacity) scores."""
    dev = scores.device
    cap = max(capacity, 1)
    out_t = torch.empty(B, cap, 3, dtype=torch.float32, device=dev)
    out_s = torch.empty(B, cap, 3, dtype=torch.float32, device=dev)
    out_sc = torch.empty(B, cap, dtype=torch.float32, device=dev)
    _lib.call("roitr_fine_gather_batched", c_int(B), c_int(cap), c_int(corr_t.shape[1]), c_int(tgt_knn.shape[1]),
              c_int(src_knn.shape[1]), c_int(tgt_pts.shape[0] // B), c_int(src_pts.shape[0] // B), i32(flat), i32(count), f32(scores),
              i32(corr_t), i32(corr_s), i32(tgt_knn), i32(src_knn), f32(tgt_pts), f32(src_pts), f32(out_t), f32(out_s), f32(out_sc),
              stream_ptr())
    return out_t, out_s, out_sc


def fine_gather(capacity, flat, count, scores, corr_t, corr_s, tgt_knn, src_knn, tgt_pts, src_pts):
    t, s_, sc = fine_gather_batched(1, capacity, flat.view(1, -1), count.view(1), scores[None], corr_t[None], corr_s[None],
                                    tgt_knn[None], src_knn[None], tgt_pts, src_pts)
    return t[0], s_[0], sc[0]


def pad_transform(pts, rot=None, trans=None):
    N = pts.shape[0]
    out = torch.empty(N + 1, 3, dtype=torch.float32, device=pts.device)
    _lib.call("roitr_pad_transform", c_int(N), f32(pts), f32(rot), f32(trans), f32(out), stream_ptr())
    return out


def pad_transform_batched(B, N, pts, rot=None, trans=None):
    """pts (B*N,3), rot (B,3,3), trans (B,3,1) -> (B*(N+1),3): every cloud followed by its (transformed) zero pad row."""
    out = torch.empty(B * (N + 1), 3, dtype=torch.float32, device=pts.device)
    _lib.call("roitr_pad_transform_batched", c_int(B), c_int(N), f32(pts), f32(rot), f32(trans), f32(out), stream_ptr())
    return out


def node_occlusion_batched(B, knn, kmask, nmask, nn_dist, thr=0.0375):
    """knn / kmask (B,M,K), nmask (B,M), nn_dist (B,N+1) -> occ (B,M)."""
    M, K = knn.shape[1], knn.shape[2]
    occ = torch.empty(B, M, dtype=torch.float32, device=knn.device)
    _lib.call("roitr_node_occlusion_batched", c_int(B), c_int(M), c_int(K), c_int(nn_dist.shape[1]), i32(knn), _u8(kmask),
              _u8(nmask), f32(nn_dist), c_float(thr), f32(occ), stream_ptr())
    return occ


def node_occlusion(knn, kmask, nmask, nn_dist, thr=0.0375):
    return node_occlusion_batched(1, knn[None], kmask[None], nmask[None], nn_dist.view(1, -1), thr)[0]


def node_overlaps_batched(B, ref_nodes, src_nodes, ref_knn, src_knn, ref_kmask, src_kmask, ref_mask, src_mask, ref_pts, src_pts,
                          rot, trans, radius):
    """ref_nodes (B*Mr,3), src_nodes (B*Ms,3), knn / kmask (B,M,64), masks (B,M), pts (B*N,3), rot (B,3,3), trans (B,3[,1]) ->
    overlap (B,Mr,Ms) f32, flag (B,Mr,Ms) u8."""
    Mr, Ms, K = ref_nodes.shape[0] // B, src_nodes.shape[0] // B, ref_knn.shape[2]
    dev = ref_nodes.device
    work = torch.empty(B * (4 * Ms + 4 * Mr), dtype=torch.float32, device=dev)
    overlap = torch.empty(B, Mr, Ms, dtype=torch.float32, device=dev)
    flag = torch.empty(B, Mr, Ms, dtype=torch.uint8, device=dev)
    _lib.call("roitr_node_overlaps_batched", c_int(B), c_int(Mr), c_int(Ms), c_int(K), c_int(ref_pts.shape[0] // B),
              c_int(src_pts.shape[0] // B), f32(ref_nodes), f32(src_nodes), i32(ref_knn), i32(src_knn), _u8(ref_kmask),
              _u8(src_kmask), _u8(ref_mask), _u8(src_mask), f32(ref_pts), f32(src_pts), f32(rot), f32(trans), c_float(radius),
              f32(work), f32(overlap), _u8(flag), stream_ptr())
    return overlap, flag


def node_overlaps(ref_nodes, src_nodes, ref_knn, src_knn, ref_kmask, src_kmask, ref_mask, src_mask, ref_pts, src_pts,
                  rot, trans, radius):
    ov, fl = node_overlaps_batched(1, ref_nodes, src_nodes, ref_knn[None], src_knn[None], ref_kmask[None], src_kmask[None],
                                   ref_mask[None], src_mask[None], ref_pts, src_pts, rot.reshape(1, 3, 3), trans.reshape(1, 3), radius)
    return ov[0], fl[0]


def corr_gather_batched(B, capacity, Mr, Ms, flat, count, overlap):
    dev = overlap.device
    cap = max(capacity, 1)
    out_idx = torch.empty(B, cap, 2, dtype=torch.int64, device=dev)
    out_ov = torch.empty(B, cap, dtype=torch.float32, device=dev)
    _lib.call("roitr_corr_gather_batched", c_int(B), c_int(cap), c_int(Mr), c_int(Ms), i32(flat), i32(count), f32(overlap),
              ptr(out_idx), f32(out_ov), stream_ptr())
    return out_idx, out_ov


def corr_gather(capacity, Ms, flat, count, overlap):
    Mr = overlap.shape[0]
    i, o = corr_gather_batched(1, capacity, Mr, Ms, flat.view(1, -1), count.view(1), overlap[None])
    return i[0], o[0]
