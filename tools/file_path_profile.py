#!/usr/bin/env python
"""cProfile of one warm run of the file-level entry point (BAM + VCF + FASTA -> FASTQ files).
usage: tools/file_path_profile.py [pairs per dataset]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeanonymizer_b200 import synth                            # noqa: E402
from genomeanonymizer_b200.engine import Engine                    # noqa: E402
from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import run_short_read_tumor_normal_anonymizer  # noqa: E402
from tests import helpers as H                                     # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    contig_len = 150 * 2 * n_pairs // 30
    case = synth.make_case(seed=9, contig_len=contig_len, n_pairs=(n_pairs, n_pairs), read_len=150,
                           somatic_positions=list(range(3000, contig_len - 3000, 4000)))
    tmp = tempfile.mkdtemp(prefix="ga_files_")
    vcf = [[case["contig"], w["keep"]["pos"] + 1, w["keep"]["pos"] + 1, 1, "N", w["keep"]["allele"], "SNV"] for w in case["windows"]]
    t, n, fa, vc = H.write_sample_files(tmp, case, vcf)
    eng = Engine(0)
    args = ([vc], [(t, n)], fa, eng, [(os.path.join(tmp, "T.out"), os.path.join(tmp, "N.out"))], True, 0, False)
    for _ in range(2):
        t0 = time.perf_counter()
        res = run_short_read_tumor_normal_anonymizer(*args)
        dt = time.perf_counter() - t0
        print(f"reads {res[0]['reads']} sessions {res[0]['sessions']}: {1e3 * dt:.1f} ms = {res[0]['reads'] / dt / 1e6:.2f} M reads/s")
    pr = cProfile.Profile()
    pr.enable()
    run_short_read_tumor_normal_anonymizer(*args)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)


if __name__ == "__main__":
    main()
