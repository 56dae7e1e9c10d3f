"""Multi-GPU sharding of the hot path: point-cloud pairs are independent, so rank r of W owns pairs r, r+W, r+2W, ...
and there is NO data-path collective. torch.distributed is used for exactly three things:
the barrier either side of a timed region, a MAX-reduce of the per-rank device time, and the gather of the
per-rank result counts (what the reference's tester does with its per-pair metrics, reference test.py / lib/tester.py).

Backend-agnostic on purpose: `nccl` on the GPUs, `gloo` in the CPU tests (tests/test_multigpu_cpu.py).
"""
import torch


def owned_pairs(rank, world, batch, n_batches=1):
    """Global pair indices owned by `rank`, as n_batches lists of `batch` indices. Step j of rank r processes pairs
    r + world*(j*batch + i), i < batch: every global index in [0, world*batch*n_batches) is owned by exactly one rank."""
    if not (0 <= rank < world) or batch < 1 or n_batches < 1:
        raise ValueError("owned_pairs: need 0 <= rank < world, batch >= 1, n_batches >= 1")
    return [[rank + world * (j * batch + i) for i in range(batch)] for j in range(n_batches)]


def owner_of(pair_index, world):
    """Inverse of owned_pairs: (rank, local position) of a global pair index."""
    return pair_index % world, pair_index // world


def max_over_ranks(values, dist=None, device="cpu"):
    """Element-wise MAX over ranks of a list of floats (device-timed milliseconds): the job is as slow as its slowest rank."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_counts(count, dist=None, device="cpu"):
    """Per-rank integer result (e.g. correspondences found in the last step) gathered to every rank, rank order."""
    c = torch.tensor([int(count)], dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.zeros_like(c) for _ in range(dist.get_world_size())]
        dist.all_gather(out, c)
        return [int(x) for x in out]
    return [int(c)]


def gather_correspondences(local, dist=None, device="cpu", dst=0):
    """The result gather of a sharded run: ``local`` = this rank's list of (global_pair_index, tensor (c_i, 7)) - per pair the
    rows [tgt_xyz | src_xyz | score] of its correspondences, c_i data dependent. Returns on rank ``dst`` a dict
    {global_pair_index: tensor (c_i, 7)} holding every rank's pairs, and None on the other ranks. Two collectives: an
    all_gather of the per-rank (pair index, count) tables (padded to the longest), then one all_gather of the payloads padded
    to the largest per-rank total - a gatherv expressed with the collectives both nccl and gloo implement."""
    idx = torch.tensor([[int(g), int(t.shape[0])] for g, t in local], dtype=torch.int64, device=device).reshape(-1, 2)
    payload = torch.cat([t.reshape(-1, 7).to(device=device, dtype=torch.float32) for _, t in local]) if local else \
        torch.zeros(0, 7, dtype=torch.float32, device=device)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        out, o = {}, 0
        for g, c in idx.tolist():
            out[g] = payload[o:o + c]
            o += c
        return out
    world = dist.get_world_size()
    sizes = torch.tensor([idx.shape[0], payload.shape[0]], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    max_pairs, max_rows = max(int(s[0]) for s in all_sizes), max(int(s[1]) for s in all_sizes)
    idx_pad = torch.zeros(max(max_pairs, 1), 2, dtype=torch.int64, device=device)
    idx_pad[:idx.shape[0]] = idx
    pay_pad = torch.zeros(max(max_rows, 1), 7, dtype=torch.float32, device=device)
    pay_pad[:payload.shape[0]] = payload
    all_idx = [torch.zeros_like(idx_pad) for _ in range(world)]
    all_pay = [torch.zeros_like(pay_pad) for _ in range(world)]
    dist.all_gather(all_idx, idx_pad)
    dist.all_gather(all_pay, pay_pad)
    if dist.get_rank() != dst:
        return None
    out = {}
    for r in range(world):
        o = 0
        for g, c in all_idx[r][:int(all_sizes[r][0])].tolist():
            out[g] = all_pay[r][o:o + c]
            o += c
    return out


def job_throughput(pairs_per_step_per_rank, world, steps, ms_max):
    """Whole-job pairs/s: units all ranks processed / the max-over-ranks device time."""
    return world * pairs_per_step_per_rank * steps / (ms_max * 1e-3)
