#!/usr/bin/env python
"""Where the end-to-end ceiling of an 8-GPU box comes from: pinned host -> device copy bandwidth of every GPU alone and of
all GPUs together (and the same for device -> host, and both directions at once).  One process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_ceiling.py
Rank 0 prints one JSON line."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 1 << 30
    host_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host_a.fill_(1); host_b.fill_(2)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def timed(fn, reps=4):
        fn(); torch.cuda.synchronize(dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        barrier()
        return reps * n / dt / 1e9

    h2d = lambda: d_a.copy_(host_a, non_blocking=True)
    d2h = lambda: host_b.copy_(d_b, non_blocking=True)

    def both():
        d_a.copy_(host_a, non_blocking=True)
        with torch.cuda.stream(s2):
            host_b.copy_(d_b, non_blocking=True)
        torch.cuda.current_stream(dev).wait_stream(s2)

    out = {}
    for name, fn in (("h2d", h2d), ("d2h", d2h), ("both_directions", both)):
        alone = torch.zeros(world, dtype=torch.float64, device=dev)
        for r in range(world):                                           # one GPU at a time
            if r == rank:
                alone[r] = _solo(fn, dev, n)
            barrier()
        together = torch.zeros(world, dtype=torch.float64, device=dev)
        together[rank] = timed(fn)
        if world > 1:
            dist.all_reduce(alone); dist.all_reduce(together)
        out[name] = {"alone_gbs_per_gpu": [round(x, 1) for x in alone.tolist()], "together_gbs_per_gpu": [round(x, 1) for x in together.tolist()],
                     "together_aggregate_gbs": round(float(together.sum().item()), 1), "bytes_per_copy": n,
                     "note": "both_directions counts the bytes of ONE direction (each direction moves as many)" if name == "both_directions" else ""}
    # host memory bandwidth of one rank (a plain host copy), for scale
    t0 = time.perf_counter()
    for _ in range(4):
        host_b.copy_(host_a)
    out["host_memcpy_gbs_rank0"] = round(4 * n / (time.perf_counter() - t0) / 1e9, 1)
    if rank == 0:
        out["n_gpus"] = world
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _solo(fn, dev, n, reps=4):
    fn(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(dev)
    return reps * n / (time.perf_counter() - t0) / 1e9


if __name__ == "__main__":
    main()
