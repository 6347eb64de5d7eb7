#!/usr/bin/env python
"""Randomised check of the ORACLE against the REFERENCE ITSELF (build container only: needs /root/reference): many small
seeded samples with random shape parameters go through the reference's unmodified CompleteGermlineAnonymizer.anonymize
(under tests/ref_stub, exactly as tests/golden/make_golden.py runs it) and through oracle/ga_oracle.c; sequences, printed
qualities, per-session counters and the set of session reads must agree.  This is how the normal-column rule (an insertion
that ends its read, golden case K-trailing-ins) would have been found; it widens the six seeded random golden sessions.
usage: tools/fuzz_reference.py [first seed] [cases] [--twist]   (--twist: trailing / leading insertions, hard clips, reference skips)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, ROOT)
import make_golden as MG                               # noqa: E402  (imports the reference under the stubs)
from genomeanonymizer_b200 import batch as B           # noqa: E402
from genomeanonymizer_b200 import synth                # noqa: E402
from oracle import oracle                              # noqa: E402
from tests import helpers as H                         # noqa: E402


def main():
    seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
    n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 100
    bad = 0
    for seed in range(seed0, seed0 + n):
        rng = np.random.default_rng(seed)
        read_len = int(rng.choice([20, 36, 50, 75, 100]))
        kw = dict(seed=seed, contig_len=int(rng.integers(1500, 4000)), n_pairs=(int(rng.integers(5, 90)), int(rng.integers(0, 90))), read_len=read_len,
                  snp_rate=float(rng.choice([1e-3, 5e-3, 2e-2])), indel_rate=float(rng.choice([0, 1e-3, 5e-3, 2e-2])),
                  clip_frac=float(rng.choice([0, 0.2, 0.6])), max_indel=int(rng.choice([2, 6, 15, 30])))
        case = synth.make_case(**kw)
        case["name"] = f"fuzz-{seed}"
        if "--twist" in sys.argv:
            case["reads"] = H.twist_reads(case["reads"])
        batch = B.pack_reads(H.ordered_reads(case))
        for widx, w in enumerate(case["windows"]):
            try:
                exp = MG.run_session(case, w)
            except Exception as exc:                                     # the reference's own exceptions are inputs it cannot process
                print("reference raised", seed, widx, repr(exc)[:120], flush=True)
                continue
            res, st = oracle.run(batch, B.pack_sessions([w]), case["reference"])
            try:
                assert st == 0, st
                H.check_session_against_golden(case, widx, exp, batch, res)
            except AssertionError as e:
                bad += 1
                print("MISMATCH", seed, widx, kw, str(e)[:300], flush=True)
        if (seed - seed0) % 25 == 24:
            print(f"{seed - seed0 + 1} cases, {bad} mismatches", flush=True)
    print(f"done: {n} cases, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
