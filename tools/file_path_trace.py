#!/usr/bin/env python
"""Stage timeline of the file-level entry point on a sample of several contigs (GA_FILE_TRACE prints a time stamp at every
stage of short_read_tumor_normal_anonymizer.anonymize_genome).  usage: tools/file_path_trace.py [contigs] [pairs per dataset]"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomeanonymizer_b200 import synth                            # noqa: E402
from genomeanonymizer_b200.engine import Engine                    # noqa: E402
from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import run_short_read_tumor_normal_anonymizer  # noqa: E402
from tests import helpers as H                                     # noqa: E402


def main():
    n_contigs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    per = n_pairs // n_contigs
    contig_len = 150 * 2 * per // 30
    cases = [synth.make_case(seed=11 + k, contig_len=contig_len, n_pairs=(per, per), read_len=150, somatic_positions=list(range(3000, contig_len - 3000, 4000)))
             for k in range(n_contigs)]
    names = [f"ctg{k}" for k in range(n_contigs)]
    tmp = tempfile.mkdtemp(prefix="ga_trace_")
    reads = [dict(r, contig=names[k], name=f"k{k}_{r['name']}") for k, c in enumerate(cases) for r in c["reads"]]
    t, n, fa, vc = (os.path.join(tmp, f) for f in ("T.bam", "N.bam", "ref.fa", "somatic.vcf"))
    contigs = [(names[k], len(c["reference"])) for k, c in enumerate(cases)]
    H.write_bam(t, contigs, [r for r in reads if r["dataset"] == 0])
    H.write_bam(n, contigs, [r for r in reads if r["dataset"] == 1])
    H.write_fasta(fa, [(names[k], c["reference"]) for k, c in enumerate(cases)])
    H.write_vcf(vc, [[names[k], w["keep"]["pos"] + 1, w["keep"]["pos"] + 1, 1, "N", w["keep"]["allele"], "SNV"] for k, c in enumerate(cases) for w in c["windows"]])
    eng = Engine(0)
    for it in range(4):
        for f in os.listdir(tmp):
            if f.endswith(".fastq") or f.endswith(".statistics.txt"):
                os.remove(os.path.join(tmp, f))
        if it == 3:
            os.environ["GA_FILE_TRACE"] = "1"
        t0 = time.perf_counter()
        res = run_short_read_tumor_normal_anonymizer([vc], [(t, n)], fa, eng, [(os.path.join(tmp, "T.out"), os.path.join(tmp, "N.out"))], True, 0, False)
        dt = time.perf_counter() - t0
        print(f"reads {res[0]['reads']}: {1e3 * dt:.1f} ms = {res[0]['reads'] / dt / 1e6:.2f} M reads/s", flush=True)


if __name__ == "__main__":
    main()
