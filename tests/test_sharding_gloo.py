"""Multi-rank host logic on CPU (gloo, world_size 2): window sharding, the counter all-reduce and the genome-order
merge give exactly the single-rank result.  The CPU oracle stands in for the device."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from genomeanonymizer_b200 import sharding
from genomeanonymizer_b200 import synthdev as SD


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, out_dir):
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = SD.WORKLOADS[name]
    w0, nw = SD.shard_windows(cfg.total_windows, world, rank)
    batch, sessions, ref = SD.generate_host(cfg, w0, nw)
    res, st = oracle.run(batch, sessions, ref)
    assert st == 0
    total = sharding.all_reduce_counters(sharding.counters_vector(res.sess_counts))
    gathered = [None] * world
    dist.all_gather_object(gathered, (res, w0, batch.n_reads, batch.n_tumor))
    if rank == 0:
        np.save(os.path.join(out_dir, "counters.npy"), total)
        import pickle
        with open(os.path.join(out_dir, "shards.pkl"), "wb") as f:
            pickle.dump(gathered, f)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["tiny", "tiny-stress"])
def test_two_rank_sharding_equals_single_rank(tmp_path, name):
    import pickle
    from oracle import oracle
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    cfg = SD.WORKLOADS[name]
    batch, sessions, ref = SD.generate_host(cfg)
    full, st = oracle.run(batch, sessions, ref)
    assert st == 0
    counters = np.load(tmp_path / "counters.npy")
    assert list(counters[:3]) == full.totals["masked"] and counters[3:].sum() == 0
    shards = pickle.load(open(tmp_path / "shards.pkl", "rb"))
    # synthetic shards hold whole windows: reads of window w sit at the same offset inside their dataset
    pl = cfg.plan()
    per_t, per_n = pl.reads_per_window
    parts = []
    for res, w0, n_reads, n_tumor in shards:
        read_map = np.concatenate([np.arange(n_tumor) + w0 * per_t,
                                   np.arange(n_reads - n_tumor) + int(pl.n_tumor) + w0 * per_n])
        parts.append((res, w0, read_map))
    merged = sharding.merge_shard_results(parts)
    assert merged.totals == full.totals
    assert np.array_equal(merged.sess_counts, full.sess_counts)
    assert sorted(merged.records) == sorted(full.records)
    for k, v in full.records.items():
        assert np.array_equal(v["seq"], merged.records[k]["seq"])
        assert (v["qual"] is None) == (merged.records[k]["qual"] is None)
        if v["qual"] is not None:
            assert np.array_equal(v["qual"], merged.records[k]["qual"])


def test_shard_sessions_balances_by_weight():
    w = [10] * 100
    assert sharding.shard_sessions(w, 4) == [(0, 25), (25, 50), (50, 75), (75, 100)]
    spans = sharding.shard_sessions([1000] + [1] * 99, 2)
    assert spans[0] == (0, 1) and spans[1] == (1, 100)
    assert sharding.shard_sessions([], 3) == [(0, 0)] * 3
    spans = sharding.shard_sessions([5, 5, 5], 8)
    assert spans[0][0] == 0 and spans[-1][1] == 3 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
