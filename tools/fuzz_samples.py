#!/usr/bin/env python
"""Randomised whole-sample parity on the GPU: driver.anonymize_sample (plan, ga_run, ga_fastq_render, the host's second
application for the reads the plan flags - quirk Q12) against plan_sample + oracle + the oracle's renderer, on seeded
samples with random coverage, window layout, orphans and placed-unmapped mates (the same shapes tools/fuzz_genome.py runs
against the reference's own driver on the CPU).  usage: tools/fuzz_samples.py [first seed] [cases]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeanonymizer_b200 import synth                             # noqa: E402
from genomeanonymizer_b200.engine import Engine                     # noqa: E402
from tests.test_genome_files import device_sample_equals_plan_and_oracle   # noqa: E402
from tests.test_plan_native import unmap_some                       # noqa: E402


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 760000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    eng = Engine(0)
    bad = flagged = 0
    for seed in range(seed0, seed0 + n):
        rng = np.random.default_rng(seed)
        clen = int(rng.integers(5000, 12000))
        nwin = int(rng.integers(0, 4))
        som = sorted(int(x) for x in rng.choice(np.arange(1500, clen - 1500, 50), size=nwin, replace=False)) if nwin else []
        if any(b - a < 2100 for a, b in zip(som, som[1:])):
            continue                                                 # variants closer than a window: refused, as pysam refuses them
        kw = dict(contig_len=clen, n_pairs=(int(rng.integers(8, 160)), int(rng.integers(8, 160))), read_len=int(rng.choice([50, 80, 100, 150])),
                  somatic_positions=som, snp_rate=float(rng.choice([1e-3, 4e-3])), indel_rate=float(rng.choice([0, 8e-4, 3e-3, 6e-3])),
                  clip_frac=float(rng.choice([0, 0.1, 0.4])))
        case = synth.make_case(seed, name=f"fuzz-{seed}", **kw)
        reads = [dict(r) for r in case["reads"]]
        drop = int(rng.choice([0, 0, 5, 9]))
        if drop:
            reads = [r for k, r in enumerate(reads) if k % drop != 3]
        unmap = int(rng.choice([0, 0, 3, 6]))
        if unmap:
            reads = unmap_some(reads, unmap)
        case["reads"] = reads
        try:
            flagged += device_sample_equals_plan_and_oracle(eng, case)
        except AssertionError as exc:
            bad += 1
            print("MISMATCH", seed, kw, "drop", drop, "unmap", unmap, str(exc)[:200], flush=True)
        except ValueError as exc:
            print("refused", seed, str(exc)[:120], flush=True)
        if (seed - seed0) % 20 == 19:
            print(f"{seed - seed0 + 1} cases, {bad} mismatches, {flagged} reads masked twice", flush=True)
    eng.close()
    print(f"done: {n} cases, {bad} mismatches, {flagged} reads masked twice")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
