#!/usr/bin/env python
"""Level (iii) timing: tumor / normal BAM + VCF + FASTA -> FASTQ files through
run_short_read_tumor_normal_anonymizer, stage by stage.  usage: tools/file_path_timing.py [pairs per dataset]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeanonymizer_b200 import driver as D                      # noqa: E402
from genomeanonymizer_b200 import genome_files as GF               # noqa: E402
from genomeanonymizer_b200 import synth                            # noqa: E402
from genomeanonymizer_b200.engine import Engine                    # noqa: E402
from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import run_short_read_tumor_normal_anonymizer  # noqa: E402
from tests import helpers as H                                     # noqa: E402  (test-side BAM / FASTA / VCF writers)


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    contig_len = 150 * 2 * n_pairs // 30                             # ~30x per dataset
    somatic = list(range(3000, contig_len - 3000, 4000))
    case = synth.make_case(seed=9, contig_len=contig_len, n_pairs=(n_pairs, n_pairs), read_len=150, somatic_positions=somatic)
    tmp = tempfile.mkdtemp(prefix="ga_files_")
    vcf = [[case["contig"], w["keep"]["pos"] + 1, w["keep"]["pos"] + 1, 1, "N", w["keep"]["allele"], "SNV"] for w in case["windows"]]
    t, n, fa, vc = H.write_sample_files(tmp, case, vcf)
    eng = Engine(0)
    for rep in range(2):                                             # the second pass is the warm one
        t0 = time.perf_counter()
        with GF.BamFile(t) as T, GF.BamFile(n) as N:
            cb = GF.pack_tumor_normal(T, N, case["contig"])
        t1 = time.perf_counter()
        table = cb.read_table()
        plan_py = D.plan_sample(table, case["windows"], contig_len)
        t2 = time.perf_counter()
        plan = GF.plan_contig(cb, case["windows"], contig_len)
        t3 = time.perf_counter()
        run_short_read_tumor_normal_anonymizer([vc], [(t, n)], fa, eng, [(os.path.join(tmp, "T.out"), os.path.join(tmp, "N.out"))], True, 0, False)
        t4 = time.perf_counter()
    reads = cb.batch.n_reads
    print(f"reads {reads}  sessions {len(plan.sessions)}  BAM decode+pack {1e3 * (t1 - t0):.1f} ms  Python plan (rows + plan_sample, not on the path) {1e3 * (t2 - t1):.1f} ms  "
          f"native plan {1e3 * (t3 - t2):.1f} ms  whole entry point {1e3 * (t4 - t3):.1f} ms = {reads / (t4 - t3) / 1e3:.1f} K reads/s")


if __name__ == "__main__":
    main()
