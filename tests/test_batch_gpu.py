"""The batched / CUDA-graph execution path must give exactly what the single-pair forward gives."""
import pytest
import torch

from roitr_b200 import model
from roitr_b200.synthetic import forward_args, synthetic_pair
from tests.helpers import CONFIG_3D, weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("graph", [False, True])
def test_batch_runner_equals_single_pair_forward(graph):
    N, B = 2048, 3
    m = model.create_model(CONFIG_3D)
    m.load_state_dict(weights(1))
    m = m.to(DEV).eval()
    pairs = [synthetic_pair(10 + i, N) for i in range(B)]
    singles = [m(*forward_args(p, DEV)) for p in pairs]
    r = m.batch_runner(B, N, N, graph=graph)
    for rep in range(2):                      # second round replays the captured graph with fresh inputs
        order = pairs if rep == 0 else pairs[::-1]
        r.load([{k: v.to(DEV) for k, v in p.items()} for p in order])
        r.run()
        outs = r.results()
        ref = singles if rep == 0 else singles[::-1]
        for o, s in zip(outs, ref):
            assert set(o) == set(s)
            for k in s:
                assert o[k].shape == s[k].shape and o[k].dtype == s[k].dtype, k
                assert torch.equal(o[k], s[k]), k      # same kernels, same arithmetic: bit-identical


def test_collated_load_and_packed_correspondences_equal_results():
    """The e2e path of bench.py (collated pinned host batch in, packed correspondences out) returns what results() returns."""
    from roitr_b200.engine import BatchRunner
    N, B = 2048, 3
    m = model.create_model(CONFIG_3D)
    m.load_state_dict(weights(1))
    m = m.to(DEV).eval()
    pairs = [synthetic_pair(20 + i, N) for i in range(B)]
    r = m.batch_runner(B, N, N, graph=True)
    r.load([{k: v.to(DEV) for k, v in p.items()} for p in pairs])
    r.run()
    ref = r.results()
    r.load_batched(BatchRunner.collate(pairs))
    r.run()
    got = r.correspondences()
    for (t, s, c), o in zip(got, ref):
        assert torch.equal(t, o["tgt_corr_points"].cpu()) and torch.equal(s, o["src_corr_points"].cpu())
        assert torch.equal(c, o["corr_scores"].cpu())


@pytest.mark.parametrize("name,n_src,n_tgt,B", [("3dmatch_30k", 30000, 28000, 2), ("4dmatch_8k", 8000, 7000, 2)])
def test_batch_runner_other_baseline_configs(name, n_src, n_tgt, B):
    """BASELINE.json configs 3 and 5 (3DMatch-shaped ~30k clouds; 4DMatch-shaped ~8k non-rigid pair, factor-2 backbone and
    the adaptive head) with UNEQUAL cloud sizes: the batched CUDA-graph path equals the single-pair forward bit for bit."""
    from tests.helpers import CONFIG_4D
    four_d = name.startswith("4d")
    cfg = CONFIG_4D if four_d else CONFIG_3D
    m = model.create_model(cfg)
    m.load_state_dict(weights(2 if four_d else 1))
    m = m.to(DEV).eval()
    pairs = []
    for i in range(B):
        p = synthetic_pair(30 + i, n_src, deform=four_d)
        p = dict(p)
        for k in ("tgt_pcd", "tgt_feats", "tgt_normals"):
            p[k] = p[k][:n_tgt].contiguous()
        pairs.append(p)
    singles = [m(*forward_args(p, DEV)) for p in pairs]
    r = m.batch_runner(B, n_src, n_tgt, graph=True)
    r.load([{k: v.to(DEV) for k, v in p.items()} for p in pairs])
    r.run()
    for o, s in zip(r.results(), singles):
        assert set(o) == set(s)
        for k in s:
            assert o[k].shape == s[k].shape and torch.equal(o[k], s[k]), k
    assert sum(int(s["corr_scores"].shape[0]) for s in singles) > 0


def test_pipelined_runner_equals_single_pair_forward():
    """engine.PipelinedRunner: consecutive steps overlap in time (two runners, two graphs, staggered by an external event
    recorded inside the graph); every step must still return exactly the single-pair forward of its own pairs."""
    from roitr_b200.engine import BatchRunner
    N, B = 2048, 2
    m = model.create_model(CONFIG_3D)
    m.load_state_dict(weights(1))
    m = m.to(DEV).eval()
    batches = [[synthetic_pair(40 + 2 * s + i, N) for i in range(B)] for s in range(5)]
    singles = [[m(*forward_args(p, DEV)) for p in batch] for batch in batches]
    pr = m.pipelined_runner(B, N, N, depth=2, mid_level=1)
    slots = []
    got = [None] * len(batches)
    clone = lambda outs: [{k: v.clone() for k, v in o.items()} for o in outs]     # outputs are views of buffers the slot reuses
    for s, batch in enumerate(batches):
        slots.append(pr.submit(BatchRunner.collate(batch)))
        if s >= 1:                       # read step s-1 while step s runs (its slot is reused only at step s+1)
            pr.wait(slots[s - 1])
            got[s - 1] = clone(pr.runner(slots[s - 1]).results())
    pr.wait(slots[-1])
    got[-1] = clone(pr.runner(slots[-1]).results())
    for outs, ref in zip(got, singles):
        for o, r in zip(outs, ref):
            assert set(o) == set(r)
            for k in r:
                assert o[k].shape == r[k].shape and torch.equal(o[k], r[k]), k


def test_single_pair_forward_graph_cache_is_transparent():
    """RIGA_v2.forward serves a REPEATED cloud shape from a per-shape CUDA graph (model.graph_cache_size): same dict, same
    values bit for bit as the eager forward, outputs stay valid after later calls, other shapes still work."""
    N = 2048
    m = model.create_model(CONFIG_3D)
    m.load_state_dict(weights(1))
    m = m.to(DEV).eval()
    pairs = [synthetic_pair(60 + i, N) for i in range(4)]
    m.graph_cache_size = 0
    eager = [m(*forward_args(p, DEV)) for p in pairs]
    m.graph_cache_size = 4
    outs = [m(*forward_args(p, DEV)) for p in pairs]          # 1st eager, 2nd captures, 3rd / 4th replay
    assert len(m._graph_cache) == 1
    other = m(*forward_args(synthetic_pair(70, 1500), DEV))    # a different shape in between (eager: first sight)
    again = m(*forward_args(pairs[0], DEV))                    # replay with the first pair's data
    for o, e in zip(outs + [again], eager + [eager[0]]):
        assert set(o) == set(e)
        for k in e:
            assert o[k].shape == e[k].shape and o[k].dtype == e[k].dtype and torch.equal(o[k], e[k]), k
    assert other["src_points"].shape[0] == 1500
    m.float()                                                  # _apply drops the captured graphs (they hold the old weights)
    assert "_graph_cache" not in m.__dict__
