"""Region sharding across the GPUs of one box.

Sessions are independent units (fresh dictionaries per anonymize() call, anonymizer_methods.py:433-439,534), so
the genome-ordered session list is cut into contiguous ranges, one per rank; a read that overlaps sessions of two
ranks is simply present in both shards (it is evidence in both sessions).  There is no collective on the data path.
The only exchange is one all-reduce (sum) of the masking counters - the values AnonymizedVariantsStatistics
accumulates (short_read_tumor_normal_anonymizer.py:198-204) - and the host merge of the per-rank records in genome
order, which is a concatenation because session indices are global.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .batch import MaskResult

N_VARIANT_TYPES = 8          # VariantType: SNV, DEL, INS, DUP, INV, CNV, TRA, SGL (SR.py:218-219)


def shard_sessions(weights: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) session ranges balanced by weight (e.g. session reads); the reference's own
    precedent balances its experimental region split by window base pairs (SR.py:795-812)."""
    w = np.asarray(weights, dtype=np.int64)
    n = len(w)
    if n == 0:
        return [(0, 0)] * world_size
    cum = np.cumsum(w)
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        k = int(np.searchsorted(cum, target, side="left")) + 1
        cuts.append(min(max(k, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world_size)]


def counters_vector(sess_counts: np.ndarray) -> np.ndarray:
    """Per-shard masking counters in the reference's 8-column VariantType layout."""
    out = np.zeros(N_VARIANT_TYPES, dtype=np.int64)
    if sess_counts is not None and len(sess_counts):
        out[:3] = np.asarray(sess_counts)[:, :3].sum(axis=0)
    return out


def all_reduce_counters(local: np.ndarray, device=None) -> np.ndarray:
    """Sum of the per-rank counter vectors (NCCL on GPUs, gloo on CPU); identity when not distributed."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(np.asarray(local, dtype=np.int64))
    if dist.is_available() and dist.is_initialized():
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu()
    return t.numpy()


def merge_shard_results(shards: Sequence[Tuple[MaskResult, int, np.ndarray]]) -> MaskResult:
    """shards: (result, first global session index, map shard read index -> global read index) per rank, in rank
    (= genome) order.  Returns one MaskResult keyed by global (session, read)."""
    out = MaskResult()
    counts = []
    totals = {"n_modified": 0, "seq16_used": 0, "qual16_used": 0, "session_reads": 0, "session_bases": 0,
              "indel_records": 0, "masked": [0, 0, 0], "error": 0, "error_detail": 0}
    for res, s0, read_map in shards:
        for (s, r), v in res.records.items():
            key = (s0 + s, int(read_map[r]))
            if key in out.records:
                raise ValueError(f"(session, read) {key} produced by two shards")
            out.records[key] = v
        counts.append(np.asarray(res.sess_counts))
        for k in ("n_modified", "seq16_used", "qual16_used", "session_reads", "session_bases", "indel_records"):
            totals[k] += res.totals[k]
        totals["masked"] = [a + b for a, b in zip(totals["masked"], res.totals["masked"])]
    out.sess_counts = np.concatenate(counts) if counts else np.zeros((0, 4), np.uint32)
    out.totals = totals
    return out


def bind_to_gpu_numa_node(device_index: int):
    """Pins the calling process (one process per GPU) to the CPUs of the NUMA node its GPU hangs on, so that the
    pinned host batches it allocates afterwards are local to that GPU's PCIe root (first-touch placement): with
    several ranks uploading at once the uploads then stop crossing the socket interconnect.  Returns
    {"node": n, "cpus": count} or None when the topology cannot be read (single-node hosts, containers without /sys)."""
    import os
    import torch
    try:
        p = torch.cuda.get_device_properties(device_index)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except (OSError, ValueError, AttributeError):
        return None
