"""Host-side mirror of the reference model interface for the forward hot path.

Same class names, constructor arguments, ``forward`` signatures and — key for ``load_state_dict(strict=True)``
(lib/trainer.py:106,110) — the same state_dict schema (522 tensors for factor 1, tests/test_model_schema.py) as
model/model.py, model/transformer/*, model/modules.py and model/RIGA_v2.py. The modules here only OWN parameters;
every ``forward`` runs the CUDA path in roitr_b200.engine (libroitr_b200). There is no PyTorch implementation of the
math to fall back to: without the CUDA library these modules raise.

Inference only: the kernels have no backward (training is outside the hot path).
"""
import math

import torch
import torch.nn as nn

from . import engine


def _holder(**children):
    m = nn.Module()
    for k, v in children.items():
        setattr(m, k, v)
    return m


class SinusoidalPositionalEmbedding(nn.Module):
    """Holds the ``div_term`` buffer (model/transformer/positional_encoding.py:38-46)."""

    def __init__(self, d_model):
        super().__init__()
        self.d_model = d_model
        self.register_buffer("div_term", torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model)))


class LocalPPFTransformer(nn.Module):
    """ppftransformer.py:202-253. Argument order (input_dim, output_dim, hidden_dim, num_heads) as in the reference."""

    def __init__(self, input_dim, output_dim, hidden_dim, num_heads, dropout=None):
        super().__init__()
        if dropout:
            raise ValueError("dropout is not supported (inference path; the reference uses dropout=None)")
        self.num_heads = num_heads
        self.embedding = _holder(embedding=SinusoidalPositionalEmbedding(hidden_dim), proj=nn.Linear(4, hidden_dim))
        self.in_proj = nn.Linear(input_dim, hidden_dim)
        self.transformer = _holder(
            attention=_holder(**{"proj_" + n: nn.Linear(hidden_dim, hidden_dim) for n in ("q", "k", "v", "p", "vp")}),
            linear=nn.Linear(hidden_dim, hidden_dim), norm=nn.LayerNorm(hidden_dim))
        self.out_proj = nn.Linear(hidden_dim, output_dim)


class RIPointTransformerLayer(nn.Module):
    def __init__(self, in_planes, out_planes, num_heads=4, nsample=16, factor=1):
        super().__init__()
        self.nsample = nsample
        self.transformer = LocalPPFTransformer(in_planes, out_planes, min(out_planes, 256 * factor), num_heads)


class TransitionDown(nn.Module):
    def __init__(self, in_planes, out_planes, num_heads=4, stride=1, nsample=16, factor=1):
        super().__init__()
        self.stride, self.nsample = stride, nsample
        self.transformer = LocalPPFTransformer(in_planes, out_planes, min(out_planes, 256 * factor), num_heads)


class TransitionUp(nn.Module):
    def __init__(self, in_planes, out_planes=None):
        super().__init__()
        if out_planes is None:
            self.linear1 = nn.Sequential(nn.Linear(2 * in_planes, in_planes), nn.LayerNorm(in_planes), nn.ReLU(inplace=True))
            self.linear2 = nn.Sequential(nn.Linear(in_planes, in_planes), nn.ReLU(inplace=True))
        else:
            self.linear1 = nn.Sequential(nn.Linear(out_planes, out_planes), nn.LayerNorm(out_planes), nn.ReLU(inplace=True))
            self.linear2 = nn.Sequential(nn.Linear(in_planes, out_planes), nn.LayerNorm(out_planes), nn.ReLU(inplace=True))


class RIPointTransformerBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, num_heads=4, nsample=16, factor=1):
        super().__init__()
        self.transformer = RIPointTransformerLayer(in_planes, planes, num_heads, nsample, factor)
        self.bn2 = nn.LayerNorm(planes)


def _attention_output(d):
    return _holder(expand=nn.Linear(d, 2 * d), squeeze=nn.Linear(2 * d, d), norm=nn.LayerNorm(d))


class GeometricTransformer(nn.Module):
    """geotransformer.py:56-133 (parameters only; evaluated by engine.geometric_transformer)."""

    def __init__(self, input_dim, output_dim, hidden_dim, num_heads, blocks, sigma_d, sigma_a, angle_k, dropout=None,
                 activation_fn="ReLU", reduction_a="max"):
        super().__init__()
        if reduction_a != "max" or activation_fn != "ReLU" or angle_k != 3 or num_heads != 4 or dropout:
            raise ValueError("only the configuration RoITr instantiates is supported (model/model.py:165)")
        for b in blocks:
            if b not in ("self", "cross"):
                raise ValueError('Unsupported block type "{}".'.format(b))
        if blocks and blocks[0] != "self":
            raise ValueError("the first block must be 'self' (cross layers consume the self layers' position states)")
        self.blocks, self.sigma_d, self.sigma_a, self.angle_k = list(blocks), sigma_d, sigma_a, angle_k
        d = hidden_dim
        self.embedding = _holder(embedding=SinusoidalPositionalEmbedding(d), proj_d=nn.Linear(d, d), proj_a=nn.Linear(d, d))
        self.in_proj = nn.Linear(input_dim, d)
        layers = []
        for b in self.blocks:
            if b == "self":
                att = _holder(attention=_holder(**{"proj_" + n: nn.Linear(d, d) for n in ("q", "k", "v", "p", "vp")}),
                              linear=nn.Linear(d, d), norm=nn.LayerNorm(d), pos_linear=nn.Linear(d, d), pos_norm=nn.LayerNorm(d))
                layers.append(_holder(attention=att, output=_attention_output(d), pos_proj=_attention_output(d)))
            else:
                att = _holder(attention=_holder(**{"proj_" + n: nn.Linear(d, d) for n in ("q", "k", "v")}),
                              linear=nn.Linear(d, d), norm=nn.LayerNorm(d))
                layers.append(_holder(attention=att, output=_attention_output(d)))
        self.transformer = _holder(layers=nn.ModuleList(layers))
        self.out_proj = nn.Linear(d, output_dim)


class _PackedMixin:
    """Caches the device-resident packed/folded weights; rebuilt when parameters move or change."""

    def repack(self):
        """Drop the packed / folded / TF32-split copy of the weights; the next forward rebuilds it. Needed only after an
        update the cache key cannot see: an in-place write through ``p.data`` (which does not bump the version counter)."""
        self._pack_key = None

    def _apply(self, fn, *a, **kw):            # .to() / .cuda() / .float(): new storage
        self._pack_key = None
        self.__dict__.pop("_graph_cache", None)
        return super()._apply(fn, *a, **kw)

    def load_state_dict(self, *a, **kw):
        self._pack_key = None
        return super().load_state_dict(*a, **kw)

    def _packed(self, prefix, architecture):
        params = list(self.parameters())
        every = params + list(self.buffers())   # buffers too: the geometric-embedding tables are built from div_term
        key = (params[0].device, tuple(t.data_ptr() for t in every), sum(t._version for t in every))
        if getattr(self, "_pack_key", None) != key:
            if params[0].device.type != "cuda":
                raise RuntimeError("roitr_b200 runs on CUDA only: call .to('cuda') / .cuda() first (no CPU fallback)")
            sd = {prefix + k: v for k, v in self.state_dict().items()}
            self._pack = engine.pack_weights(sd, params[0].device, architecture)
            self._pack_key = key
            self.__dict__.pop("_graph_cache", None)      # captured graphs hold the old packed weights
        return self._pack


class RIPointTransformer(_PackedMixin, nn.Module):
    """model/model.py:145-237: same constructor, same forward(s_pxon, t_pxon, src_deformed_pcd) -> 8-tuple."""

    def __init__(self, blocks=[2, 3, 3, 3], block=RIPointTransformerBlock, c=1, transformer_architecture=None,
                 with_cross_pos_embed=None, factor=1, occ_thres=0.):
        super().__init__()
        if list(blocks) != [2, 3, 3, 3] or block is not RIPointTransformerBlock:
            raise ValueError("only blocks=[2,3,3,3] with RIPointTransformerBlock (what RIGA_v2 builds) is supported")
        self.c, self.num_heads = c, 4
        self.in_planes, planes = c, [64 * factor, 128 * factor, 256 * factor, 256 * factor]
        stride, nsample = [1, 4, 4, 4], [8, 16, 16, 16]
        for i in range(4):
            setattr(self, "enc%d" % (i + 1), self._make_enc(block, planes[i], blocks[i], 4, stride[i], nsample[i], factor))
        for i in (3, 2, 1, 0):
            setattr(self, "dec%d" % (i + 1), self._make_dec(block, planes[i], 2, 4, nsample[i], factor, is_head=(i == 3)))
        self.nsample = nsample
        self.transformer_architecture = transformer_architecture
        self.global_transformer = GeometricTransformer(256 * factor, 256 * factor, 256 * factor, 4,
                                                       transformer_architecture, sigma_d=0.2, sigma_a=15, angle_k=3)
        self.occ_proj = nn.Linear(256 * factor, 1)   # unused by the forward; kept for strict weight loading
        self.with_cross_pos_embed = with_cross_pos_embed

    def _make_enc(self, block, planes, blocks, num_heads, stride, nsample, factor):
        layers = [TransitionDown(self.in_planes, planes, num_heads, stride, nsample, factor)]
        self.in_planes = planes
        layers += [block(planes, planes, num_heads, nsample=nsample, factor=factor) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def _make_dec(self, block, planes, blocks, share_planes, nsample, factor, is_head=False):
        layers = [TransitionUp(self.in_planes, None if is_head else planes * block.expansion)]
        self.in_planes = planes * block.expansion
        layers += [block(self.in_planes, self.in_planes, share_planes, nsample=nsample, factor=factor) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    @torch.no_grad()
    def forward(self, s_pxon, t_pxon, src_deformed_pcd):
        with torch.cuda.device(s_pxon[0].device):      # streams / events below are those of the tensors' device
            W = self._packed("backbone.", self.transformer_architecture)
            return engine.backbone_forward(W, self.transformer_architecture, s_pxon, t_pxon, src_deformed_pcd)


class LearnableLogOptimalTransport(nn.Module):
    """model/modules.py:10-72 (parameter ``alpha``; the iterations run inside csrc/matching.cu fine_patch_kernel)."""

    def __init__(self, num_iter, inf=1e6):
        super().__init__()
        self.num_iter, self.inf = num_iter, inf
        self.register_parameter("alpha", nn.Parameter(torch.tensor(1.)))

    def __repr__(self):
        return self.__class__.__name__ + "(num_iter={})".format(self.num_iter)


class CoarseMatching(nn.Module):
    def __init__(self, num_correspondences, dual_normalization=True):
        super().__init__()
        self.num_correspondences, self.dual_normalization = num_correspondences, dual_normalization


class AdaptiveSuperPointMatching(nn.Module):
    def __init__(self, min_num_correspondences, similarity_threshold=0.75):
        super().__init__()
        self.min_num_correspondences, self.similarity_threshold = min_num_correspondences, similarity_threshold


class FineMatching(nn.Module):
    def __init__(self, k, mutual=True, confidence_threshold=0.05, use_dustbin=False, use_global_score=False,
                 correspondence_threshold=3):
        super().__init__()
        if use_dustbin or use_global_score:
            raise ValueError("use_dustbin / use_global_score are False in every reference config and are not supported")
        self.k, self.mutual, self.confidence_threshold = k, mutual, confidence_threshold
        self.use_dustbin, self.use_global_score = use_dustbin, use_global_score
        self.correspondence_threshold = correspondence_threshold


class RIGA_v2(_PackedMixin, nn.Module):
    """model/RIGA_v2.py:10-175: the RoITr pipeline. ``config`` needs the 17 keys RIGA_v2.__init__ reads (attribute or
    item access). forward(...) -> dict with the reference's 22 keys."""

    KEYS = ("with_cross_pos_embed", "benchmark", "num_est_coarse_corr", "transformer_architecture", "mode",
            "point_per_patch", "matching_radius", "num_gt_coarse_corr", "coarse_overlap_threshold", "fine_matching_topk",
            "fine_matching_mutual", "fine_matching_confidence_threshold", "fine_matching_use_dustbin",
            "fine_matching_use_global_score", "fine_matching_correspondence_threshold")

    def __init__(self, config):
        super().__init__()
        get = (lambda k: config[k]) if isinstance(config, dict) else (lambda k: getattr(config, k))
        self.config = config
        self.cfg = {k: get(k) for k in self.KEYS}
        self.with_cross_pos_embed = self.cfg["with_cross_pos_embed"]
        self.benchmark = self.cfg["benchmark"]
        if self.benchmark in ("3DMatch", "3DLoMatch"):
            self.coarse_matching = CoarseMatching(self.cfg["num_est_coarse_corr"], dual_normalization=True)
            self.factor = 1
        else:
            self.coarse_matching = AdaptiveSuperPointMatching(self.cfg["num_est_coarse_corr"], similarity_threshold=0.75)
            self.factor = 2
        self.backbone = RIPointTransformer(transformer_architecture=self.cfg["transformer_architecture"],
                                           with_cross_pos_embed=self.with_cross_pos_embed, factor=self.factor)
        self.OT = LearnableLogOptimalTransport(num_iter=100)      # unused by forward; kept for strict weight loading
        self.mode = self.cfg["mode"]
        self.point_per_patch = self.cfg["point_per_patch"]
        self.matching_radius = self.cfg["matching_radius"]
        self.coarse_proj = nn.Linear(256 * self.factor, 256 * self.factor)
        self.fine_proj = nn.Linear(64 * self.factor, 256 * self.factor)
        self.fine_matching = FineMatching(self.cfg["fine_matching_topk"], mutual=self.cfg["fine_matching_mutual"],
                                          confidence_threshold=self.cfg["fine_matching_confidence_threshold"],
                                          use_dustbin=self.cfg["fine_matching_use_dustbin"],
                                          use_global_score=self.cfg["fine_matching_use_global_score"],
                                          correspondence_threshold=self.cfg["fine_matching_correspondence_threshold"])
        self.fine_matching_use_dustbin = self.cfg["fine_matching_use_dustbin"]
        self.optimal_transport = LearnableLogOptimalTransport(num_iter=100)

    @torch.no_grad()
    def forward(self, src_pcd, tgt_pcd, src_feats, tgt_feats, src_normals, tgt_normals, rot, trans, src_raw_pcd, _aux=None):
        if self.training:
            raise RuntimeError("roitr_b200 implements the inference forward only: call .eval() (training is out of scope)")
        if not src_pcd.is_cuda:
            raise RuntimeError("roitr_b200 runs on CUDA only (no CPU fallback): move the inputs to the model's device")
        with torch.cuda.device(src_pcd.device):        # streams / events below are those of the tensors' device
            W = self._packed("", self.cfg["transformer_architecture"])
            args = [t.contiguous().float() for t in (src_pcd, tgt_pcd, src_feats, tgt_feats, src_normals, tgt_normals, rot,
                                                     trans, src_raw_pcd)]
            if _aux is None and self.graph_cache_size > 0:
                out = self._forward_cached_graph(W, args)
                if out is not None:
                    return out
            return engine.riga_forward(W, self.cfg, *args, aux=_aux)

    # lib/tester.py:53 calls forward once per pair (batch_size 1): ~700 kernel launches whose HOST cost (~38 ms through ctypes)
    # exceeds their device time (~12 ms at 2 x 20 000 points). Cloud sizes that REPEAT - fixed-size inference, benchmarks,
    # voxel-capped datasets - are served from a per-shape CUDA graph instead: the second forward with a given (n_src, n_tgt)
    # captures a one-pair engine.BatchRunner, later ones replay it (inputs copied into its static buffers, outputs cloned out
    # of them, so results stay valid after the next call). Results are bit-identical to the eager path (same kernels).
    # ``graph_cache_size = 0`` disables it; least-recently-used shapes are dropped beyond the limit.
    graph_cache_size = 4

    def _forward_cached_graph(self, W, args):
        src_pcd, tgt_pcd, src_feats, tgt_feats, src_normals, tgt_normals, rot, trans, src_raw_pcd = args
        key = (int(src_raw_pcd.shape[0]), int(tgt_pcd.shape[0]), src_pcd.device, id(W))
        cache = self.__dict__.setdefault("_graph_cache", {})
        seen = self.__dict__.setdefault("_graph_seen", {})
        r = cache.get(key)
        if r is None:
            seen[key] = seen.get(key, 0) + 1
            if seen[key] < 2:                 # first sight of a shape: eager (a capture costs several forwards)
                if len(seen) > 64:
                    seen.clear()
                return None
            r = engine.BatchRunner(W, self.cfg, 1, key[0], key[1], src_pcd.device, graph=True)
            cache[key] = r
            while len(cache) > self.graph_cache_size:
                cache.pop(next(iter(cache)))
        else:
            cache[key] = cache.pop(key)        # most recently used last
        r.load([dict(src_pcd=src_pcd, tgt_pcd=tgt_pcd, src_feats=src_feats, tgt_feats=tgt_feats, src_normals=src_normals,
                     tgt_normals=tgt_normals, rot=rot, trans=trans.reshape(3, 1), src_raw_pcd=src_raw_pcd)])
        r.run()
        out = r.results()[0]                   # the one host sync (counts)
        return {k: v.clone() for k, v in out.items()}


    def batch_runner(self, batch_pairs, n_src, n_tgt, graph=True, fps_cluster=0, serial=False):
        """Throughput API (not in the reference, which is batch_size=1 only): a engine.BatchRunner that evaluates
        ``batch_pairs`` independent pairs of fixed size per step, as one CUDA graph when ``graph`` is True."""
        if self.training:
            raise RuntimeError("inference only: call .eval()")
        W = self._packed("", self.cfg["transformer_architecture"])
        dev = next(self.parameters()).device
        return engine.BatchRunner(W, self.cfg, batch_pairs, n_src, n_tgt, dev, graph=graph, fps_cluster=fps_cluster,
                                  serial=serial)


    def pipelined_runner(self, batch_pairs, n_src, n_tgt, depth=2, mid_level=1, fps_cluster=0):
        """engine.PipelinedRunner: ``depth`` batch runners whose steps overlap (the next step's bandwidth-heavy front runs
        beside the current step's latency-bound back)."""
        if self.training:
            raise RuntimeError("inference only: call .eval()")
        W = self._packed("", self.cfg["transformer_architecture"])
        dev = next(self.parameters()).device
        return engine.PipelinedRunner(W, self.cfg, batch_pairs, n_src, n_tgt, dev, depth=depth, mid_level=mid_level,
                                      fps_cluster=fps_cluster)


def create_model(config):
    """model/RIGA_v2.py:178-180."""
    return RIGA_v2(config)
