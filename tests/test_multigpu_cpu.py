"""The N>1 path on CPU: two `gloo` ranks exercise exactly the host logic bench.py / a multi-GPU deployment uses
(roitr_b200/sharding.py): pair ownership, the max-over-ranks step time and the result-count gather. No GPU, no kernels:
the data path has no collective, so this IS all the inter-rank traffic there is."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from roitr_b200 import sharding
from roitr_b200.synthetic import synthetic_pair


def test_ownership_is_a_partition():
    for world in (1, 2, 4, 8):
        for batch, nb in ((1, 1), (16, 2), (5, 3)):
            seen = []
            for r in range(world):
                own = sharding.owned_pairs(r, world, batch, nb)
                assert len(own) == nb and all(len(o) == batch for o in own)
                for j, o in enumerate(own):
                    for i, g in enumerate(o):
                        assert sharding.owner_of(g, world) == (r, j * batch + i)
                seen += [g for o in own for g in o]
            assert sorted(seen) == list(range(world * batch * nb))
    with pytest.raises(ValueError):
        sharding.owned_pairs(2, 2, 1)


def test_weak_scaling_throughput_formula():
    assert sharding.job_throughput(16, 8, 10, 500.0) == pytest.approx(8 * 16 * 10 / 0.5)


def test_single_process_paths_need_no_process_group():
    assert sharding.max_over_ranks([3.0, 4.0]) == [3.0, 4.0]
    assert sharding.gather_counts(7) == [7]
    one = sharding.gather_correspondences([(3, torch.ones(2, 7)), (5, torch.zeros(0, 7))])
    assert sorted(one) == [3, 5] and one[3].shape == (2, 7) and one[5].shape == (0, 7)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        own = sharding.owned_pairs(rank, world, 2, 2)
        # the shard's inputs are a pure function of the GLOBAL pair index: any rank can regenerate any pair
        first = synthetic_pair(own[0][0], 64)
        checksum = float(first["src_pcd"].double().sum())
        dist.barrier()
        ms = sharding.max_over_ranks([10.0 + rank, 20.0 - rank], dist)          # slowest rank wins, per element
        counts = sharding.gather_counts(100 + rank, dist)                       # rank order
        flat = torch.tensor([g for o in own for g in o], dtype=torch.int64)
        allidx = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(allidx, flat)
        # variable-length result gather: pair g contributes (g % 5) + rank rows whose first column encodes (g, row)
        local = [(g, torch.arange((g % 5) + rank, dtype=torch.float32)[:, None].repeat(1, 7) + 100.0 * g) for o in own for g in o]
        gathered = sharding.gather_correspondences(local, dist)
        dist.barrier()
        summary = None if gathered is None else {g: (t.shape[0], float(t.sum())) for g, t in gathered.items()}
        out.put((rank, ms, counts, [t.tolist() for t in allidx], own[0][0], checksum, summary))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_shard_reduce_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, counts, allidx, g0, checksum, summary in got:
        if rank == 0:      # rank 0 holds every pair's correspondences, each with its own length
            assert sorted(summary) == list(range(8))
            for g, (rows, total) in summary.items():
                r_own, _ = sharding.owner_of(g, world)
                n = (g % 5) + r_own
                assert rows == n and total == pytest.approx(7 * (n * (n - 1) / 2 + 100.0 * g * n))
        else:
            assert summary is None
        assert ms == [11.0, 20.0]
        assert counts == [100, 101]
        assert sorted(i for l in allidx for i in l) == list(range(8))          # disjoint and complete across ranks
        assert allidx[rank][0] == g0 == rank
        assert checksum == pytest.approx(float(synthetic_pair(g0, 64)["src_pcd"].double().sum()))
    assert sharding.job_throughput(2, world, 2, max(m[1][0] for m in got)) == pytest.approx(8 / 0.011)


def test_collate_layout_matches_runner_inputs():
    """BatchRunner.collate (the host-side batch layout of the e2e path): [src_0..src_{B-1}, tgt_0..tgt_{B-1}] per key."""
    import torch
    from roitr_b200.engine import BatchRunner
    from roitr_b200.synthetic import synthetic_pair
    pairs = [synthetic_pair(i, 256) for i in range(3)]
    h = BatchRunner.collate(pairs, pin=False)
    assert set(h) == set(BatchRunner.INPUT_KEYS)
    assert h["pts"].shape == (6 * 256, 3) and h["src_pcd"].shape == (3 * 256, 3) and h["rot"].shape == (3, 3, 3) and h["trans"].shape == (3, 3, 1)
    assert torch.equal(h["pts"][256:512], pairs[1]["src_raw_pcd"]) and torch.equal(h["pts"][3 * 256 + 512:], pairs[2]["tgt_pcd"])
    assert torch.equal(h["nrm"][:256], pairs[0]["src_normals"]) and torch.equal(h["feats"][3 * 256:4 * 256], pairs[0]["tgt_feats"])
    assert torch.equal(h["src_pcd"][512:], pairs[2]["src_pcd"]) and torch.equal(h["trans"][1].flatten(), pairs[1]["trans"].flatten())
