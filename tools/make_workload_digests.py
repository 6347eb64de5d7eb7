#!/usr/bin/env python
"""Oracle digests of the benchmark workloads -> tests/golden/workload_digests.json.

For every workload of bench.py the CPU oracle (oracle/ga_oracle.c) masks EVERY session and its modified records are
reduced to the order- and shard-independent digest of include/ga_digest.h, block by block (blocks = the genome-ordered
session list cut at contig boundaries and every 50,000 sessions).  bench.py compares the engine's digest with these
at every GPU count, so "parity at scale" is a comparison against the oracle's records, not against an earlier engine
run.  Runs on a GPU box only because the synthetic input generator is fastest there (the inputs are generated in HBM
and copied to the host); the engine library is never loaded.

    python tools/make_workload_digests.py [workload ...]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "workload_digests.json")
REGION = 50_000


def blocks_of(contigs):
    out, base = [], 0
    for k, c in enumerate(contigs):
        w = 0
        while w < c.total_windows:
            g = base + w
            nw = min(c.total_windows - w, REGION - g % REGION)
            out.append((k, w, nw, g))
            w += nw
        base += c.total_windows
    return out


def main():
    import numpy as np
    import torch
    from genomeanonymizer_b200 import synthdev as SD
    from oracle import oracle
    oracle.build()
    names = sys.argv[1:] or ["chr22-1k", "chr1-30x-50k", "cigar-stress", "dense-60x30x", "wgs-30x", "wgs-60x30x"]
    dev = torch.device("cuda", 0)
    try:
        with open(OUT) as f:
            table = json.load(f)
    except Exception:
        table = {}
    threads = oracle.n_threads()
    for name in names:
        contigs = SD.genome_contigs(name)
        t0 = time.perf_counter()
        blocks, total = [], np.zeros(4, np.uint64)
        reads = 0
        ref_k, ref = -1, None
        for k, w0, nw, g in blocks_of(contigs):
            cfg = contigs[k]
            if k != ref_k:
                ref = SD.reference_device(cfg, dev).cpu().numpy().tobytes()
                ref_k = k
            db, ds = SD.generate_device(cfg, dev, w0, nw)
            hb, hs = db.to_host(), ds.to_host()
            del db, ds
            torch.cuda.empty_cache()
            raw, st = oracle.run(hb, hs, ref, threads=threads, decode=False, cap_frac=0.25)
            if st != 0:
                raise SystemExit(f"oracle failed on {name} block {k}:{w0}+{nw} with status {st}")
            pl = cfg.plan(0, 0)
            d = oracle.digest(raw["result"], int(raw["totals"].n_modified), session_base=g, tumor_base=w0 * int(pl.reads_per_window[0]),
                              normal_base=w0 * int(pl.reads_per_window[1]), n_tumor=hb.n_tumor, contig=k)
            with np.errstate(over="ignore"):
                total += d
            reads += int(raw["totals"].session_reads)
            blocks.append([g, g + nw] + [int(x) for x in d])
            del hb, hs, raw
        table[name] = {"sessions": int(sum(c.total_windows for c in contigs)), "session_reads": reads, "total": [int(x) for x in total],
                       "blocks": blocks, "layout": "[first session, end session, sum hash lo, sum hash hi, records, sum of new lengths]",
                       "source": f"oracle/ga_oracle.c on every session ({threads} host threads), tools/make_workload_digests.py"}
        print(f"{name}: {len(blocks)} blocks, {reads} session reads, {int(total[2])} records, {time.perf_counter() - t0:.1f} s", flush=True)
        with open(OUT, "w") as f:
            json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()
