"""The on-disk result record of the reference's test runner and what its evaluators read back.

``lib/tester.py:56-69`` turns the 22-key output dict of ``model.forward`` plus five of its inputs into a 16-key CPU dict
(17 with ``metric_index_list`` for 4DMatch) and ``torch.save``s it as ``{snapshot_dir}/{benchmark}/{idx}.pth``;
``registration/evaluate_registration_c2f.py:68-75`` (3DMatch) and ``registration/evaluate_fdmatch.py`` (4DMatch) load those
files. This module writes exactly that record - same keys, same order, same dtypes, CPU tensors - so the unchanged
evaluators consume the output of ``roitr_b200``. Two things differ on purpose: files are named by the GLOBAL pair index
(the reference uses the local loop index, which collides when pairs are sharded over ranks), and a batch of pairs is moved
to the host through one staging buffer per dtype instead of ~16 blocking ``.cpu()`` calls per pair.
"""
import os

import torch

# (record key, source, key in source) in the reference's order (lib/tester.py:57-66)
RECORD_FIELDS = (
    ("src_raw_pcd", "inputs", "src_raw_pcd"), ("src_pcd", "inputs", "src_pcd"), ("tgt_pcd", "inputs", "tgt_pcd"),
    ("src_nodes", "outputs", "src_nodes"), ("tgt_nodes", "outputs", "tgt_nodes"),
    ("src_node_desc", "outputs", "src_node_feats"), ("tgt_node_desc", "outputs", "tgt_node_feats"),
    ("src_point_desc", "outputs", "src_point_feats"), ("tgt_point_desc", "outputs", "tgt_point_feats"),
    ("src_corr_pts", "outputs", "src_corr_points"), ("tgt_corr_pts", "outputs", "tgt_corr_points"),
    ("confidence", "outputs", "corr_scores"),
    ("gt_tgt_node_occ", "outputs", "gt_tgt_node_occ"), ("gt_src_node_occ", "outputs", "gt_src_node_occ"),
    ("rot", "inputs", "rot"), ("trans", "inputs", "trans"),
)
FOUR_D = ("4DMatch", "4DLoMatch")


_STAGING = {}     # (dtype, device) -> pinned host buffer, grown on demand and reused from step to step


def _staging(dtype, dev, numel):
    buf = _STAGING.get((dtype, dev))
    if buf is None or buf.numel() < numel:
        buf = torch.empty(int(numel * 1.25) + 16, dtype=dtype, pin_memory=True)
        _STAGING[(dtype, dev)] = buf
    return buf[:numel]


def tester_records(inputs_list, outputs_list, benchmark="3DMatch", metric_index_list=None):
    """tester_record for a BATCH of pairs with ONE device->host copy per dtype for the whole batch (reused pinned staging
    buffers; a record's tensors are views into them, valid until the next call - save_record clones). inputs_list /
    outputs_list: per pair, the forward inputs and the forward's output dict (e.g. BatchRunner.results())."""
    flat = {}
    for b, (inp, out) in enumerate(zip(inputs_list, outputs_list)):
        src = {"inputs": inp, "outputs": out}
        for key, where, name in RECORD_FIELDS:
            flat[(b, key)] = src[where][name]
    host = _to_host(flat, reuse=True)
    recs = []
    for b in range(len(outputs_list)):
        data = {key: host[(b, key)] for key, _, _ in RECORD_FIELDS}
        if benchmark in FOUR_D:
            data["metric_index_list"] = None if metric_index_list is None else metric_index_list[b]
        recs.append(data)
    return recs


def _to_host(tensors, reuse=False):
    """dict name -> tensor (any device) -> dict name -> CPU tensor. Device tensors of one dtype travel together: one
    concatenation on the device, one copy into pinned memory, views on the host."""
    out, groups = {}, {}
    for k, t in tensors.items():
        t = t.detach()
        if t.is_cuda:
            groups.setdefault((t.dtype, t.device), []).append((k, t))
        else:
            out[k] = t
    for (dtype, dev), items in groups.items():
        flat = torch.cat([t.reshape(-1) for _, t in items])
        host = _staging(dtype, dev, flat.numel()) if reuse else torch.empty(flat.shape, dtype=dtype, pin_memory=True)
        host.copy_(flat, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        o = 0
        for k, t in items:
            out[k] = host[o:o + t.numel()].view(t.shape)
            o += t.numel()
    return out


def tester_record(inputs, outputs, benchmark="3DMatch", metric_index=None):
    """inputs: the tensors handed to forward (``src_pcd, tgt_pcd, src_raw_pcd, rot, trans``; names of RIGA_v2.forward),
    outputs: forward's dict. Returns the reference's record: an insertion-ordered dict of CPU tensors (lib/tester.py:56-68)."""
    src = {"inputs": inputs, "outputs": outputs}
    host = _to_host({key: src[where][name] for key, where, name in RECORD_FIELDS})
    data = {key: host[key] for key, _, _ in RECORD_FIELDS}
    if benchmark in FOUR_D:
        data["metric_index_list"] = metric_index                      # passed through untouched (lib/tester.py:67-68)
    return data


def record_path(snapshot_dir, benchmark, global_pair_index):
    return os.path.join(snapshot_dir, benchmark, "%d.pth" % int(global_pair_index))


def save_record(data, snapshot_dir, benchmark, global_pair_index):
    """torch.save as ``{snapshot_dir}/{benchmark}/{global index}.pth`` (lib/tester.py:69). Views into a shared staging
    buffer are cloned so that each file holds only its own tensors."""
    path = record_path(snapshot_dir, benchmark, global_pair_index)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({k: (v.clone() if torch.is_tensor(v) else v) for k, v in data.items()}, path)
    return path


def load_for_registration(path):
    """What registration/evaluate_registration_c2f.py:68-75 reads from a record."""
    data = torch.load(path)
    keys = ("src_pcd", "tgt_pcd", "src_nodes", "tgt_nodes", "src_node_desc", "tgt_node_desc", "src_point_desc", "tgt_point_desc",
            "rot", "trans", "src_corr_pts", "tgt_corr_pts", "confidence")
    return {k: data[k] for k in keys}
