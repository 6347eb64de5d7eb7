#!/usr/bin/env python
"""Randomised parity: engine (ga_run through the C ABI) against the CPU oracle on many small seeded samples whose shape
parameters are drawn at random (read length 30-300, indel / clip / SNP / error rates, indel lengths up to 60, coverage,
window count).  usage: tools/fuzz_parity.py [first seed] [cases] [--twist]     exit code 1 when anything differed."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeanonymizer_b200 import batch as B           # noqa: E402
from genomeanonymizer_b200 import synth                # noqa: E402
from genomeanonymizer_b200.engine import Engine        # noqa: E402
from oracle import oracle                              # noqa: E402
from tests import helpers as H                         # noqa: E402


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 100
    eng = Engine(0)
    bad = 0
    for seed in range(seed0, seed0 + n):
        rng = np.random.default_rng(seed)
        read_len = int(rng.choice([30, 50, 75, 100, 101, 125, 150, 151, 156, 157, 180, 200, 249, 250, 300]))
        n_t, n_n = int(rng.integers(20, 500)), int(rng.integers(0, 500))
        kw = dict(seed=seed, contig_len=int(rng.integers(3000, 9000)), n_pairs=(n_t, n_n), read_len=read_len,
                  snp_rate=float(rng.choice([1e-3, 3e-3, 1e-2])), indel_rate=float(rng.choice([0, 1e-4, 1e-3, 4e-3, 1e-2])),
                  clip_frac=float(rng.choice([0, 0.1, 0.5])), max_indel=int(rng.choice([3, 10, 20, 40, 60])))
        case = synth.make_case(**kw)
        if "--twist" in sys.argv:                                            # trailing / leading insertions, hard clips, reference skips
            case["reads"] = H.twist_reads(case["reads"])
        reads = [r for r in case["reads"] if r["dataset"] == 0] + [r for r in case["reads"] if r["dataset"] == 1]
        for sparse in (False, True):
            batch = B.pack_reads(reads, sparse_qual=sparse)
            sessions = B.pack_sessions(case["windows"])
            exp, st = oracle.run(batch, sessions, case["reference"])
            assert st == 0, (seed, st)
            eng.upload_reference(0, case["reference"])
            got = eng.run(batch, sessions)
            ok = got.totals == exp.totals and sorted(got.records) == sorted(exp.records) and np.array_equal(got.sess_counts, exp.sess_counts)
            if ok:
                for k, v in got.records.items():
                    e = exp.records[k]
                    if not np.array_equal(v["seq"], e["seq"]) or (v["qual"] is None) != (e["qual"] is None) or (v["qual"] is not None and not np.array_equal(v["qual"], e["qual"])):
                        ok = False
                        print("record differs", seed, sparse, k, flush=True)
                        break
            if not ok:
                bad += 1
                print("MISMATCH", seed, sparse, kw, got.totals, exp.totals, flush=True)
        if (seed - seed0) % 20 == 19:
            print(f"{seed - seed0 + 1} cases, {bad} mismatches", flush=True)
    eng.close()
    print(f"done: {n} cases, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
