"""File-level entry point of the B200 path: tumor / normal BAM + somatic VCF + reference FASTA in, the reference's
FASTQ files and statistics file out (SURVEY.md 8(b) "Python entry point", 8(f) N4).

Mirrors `run_short_read_tumor_normal_anonymizer` (short_read_tumor_normal_anonymizer.py:889-893: same name, same
positional arguments) and `name_output` (:55-58).  Per sample the reference runs `anonymize_genome` (:625-760), which
walks the genome section by section through pysam; here every contig of the sample is decoded by the C++ readers
(genome_files.py -> csrc/ga_genome_io.cpp), planned on the host in native code (include/ga_plan.h: sections, island
sessions, mate pairing, first write wins; driver.plan_sample is the same algorithm in Python and its checker),
masked by ONE engine pass over all of its sessions and printed by the device
FASTQ renderer.  Outputs, named as the reference names them:
    <tumor_output>.1.fastq / .2.fastq, <normal_output>.1.fastq / .2.fastq       (:652-655)
    <tumor_output>.single_end.fastq, <normal_output>.single_end.fastq             (:603-622; only when a read stayed unpaired)
    <normal_bam_file>.statistics.txt                                              (:641, when record_statistics)

Differences that are stated rather than hidden: `cpus` is the number of host threads of the file decoders (the
reference forks one process per sample); `enhance_parallelization` (the reference splits a sample's BAMs into
per-region files for its process pool, :795-885) has no counterpart - one GPU pass needs no such split - and is
accepted and ignored; breakend VCF records are rejected (genome_files.read_vcf).  Mates aligned to different contigs
are paired as the reference pairs them: unpaired reads are carried from contig to contig and a pair is written where
its second mate is processed.  There is no CPU fallback: without a CUDA device the engine
constructor raises.
"""
import logging
import re
from typing import List, Optional, Tuple

from . import genome_files as GF
from .driver import anonymize_packed, statistics_text

DATASET_IDX_TUMORAL = 0      # variation_classifier.py:13
DATASET_IDX_NORMAL = 1       # variation_classifier.py:14


def name_output(sample: str) -> str:
    """short_read_tumor_normal_anonymizer.py:55-58 (the reference's own, unescaped pattern)."""
    return re.sub('.bam|.sam|.cram', '.anonymized', sample)


UNSUPPORTED_FLAGS = 0x800            # supplementary alignments


def _refuse_unsupported_records(len_flag, contig):
    """The reference keeps separate bookkeeping for supplementary alignments (primary replaces supplementary, SA tags,
    left-over variants: anonymizer_methods.py:98-149, 245-288).  It does not exist here, and treating such records as
    ordinary alignments would print truncated or unmasked reads, so the entry point refuses them instead of guessing.
    Secondary records (0x100) are first-wins in the reference as well and are accepted; placed-unmapped mates (0x4) are
    handled as the reference handles them (driver.plan_sample / ga_plan_sample)."""
    import numpy as np
    flags = np.asarray(len_flag) >> 16
    bad = flags & UNSUPPORTED_FLAGS
    if bad.any():
        raise ValueError(f"contig {contig}: {int(np.count_nonzero(bad))} supplementary (0x800) records - this engine does not implement the "
                         f"reference's supplementary-alignment bookkeeping; filter them first (samtools view -F 0x800)")


def anonymize_genome(windows_by_contig, tumor_bam_file: str, normal_bam_file: str, ref_genome_file: str, engine,
                     tumor_output_fastq: str, normal_output_fastq: str, record_statistics: bool = True, cpus: int = 0):
    """One tumor-normal sample, file to file.  Returns {"reads": n, "sessions": n, "modified_pairs_or_reads": ...}."""
    fasta = GF.FastaFile(ref_genome_file)
    outs = {("T", "1"): tumor_output_fastq + ".1.fastq", ("T", "2"): tumor_output_fastq + ".2.fastq",
            ("N", "1"): normal_output_fastq + ".1.fastq", ("N", "2"): normal_output_fastq + ".2.fastq"}
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(8)                               # the four files are opened, written and closed side by side (the calls release the GIL); one thread prepares the next contig
    # truncated first, as the reference does (:652-655), while the BAM files inflate; unbuffered: whole slices are written
    opening = {k: pool.submit(open, p, "wb", 0) for k, p in outs.items()}
    handles = {}

    def _append(handle, data):
        for part in (data if isinstance(data, list) else [data]):
            view = memoryview(part)
            while len(view):                                   # a raw write may be short
                view = view[handle.write(view):]
    import os, time
    t_start = time.perf_counter()
    trace = (lambda what: print(f"[ga-file-trace] {1e3 * (time.perf_counter() - t_start):8.1f} ms  {what}", flush=True)) if os.environ.get("GA_FILE_TRACE") else (lambda what: None)
    carry = {}                                                 # to_pair_anonymized_reads (:646): unpaired reads, kept across contigs
    stats_parts = []
    n_reads = n_sessions = 0
    try:
        opened = list(pool.map(lambda f: GF.BamFile(f, cpus), (tumor_bam_file, normal_bam_file)))    # both files inflate side by side
        trace("BAM files open")
        with opened[0] as tumor, opened[1] as normal:
            def prepare(contig_id):
                """Decode, pack and plan one contig (native code, no GIL): runs one contig ahead of the masking."""
                contig = fasta.references[contig_id]
                windows = windows_by_contig.get(contig, [])
                trace(f"pack {contig} begins")
                cb = GF.pack_tumor_normal(tumor, normal, contig, contig_id=contig_id)
                trace(f"pack {contig} done")
                if cb.batch.n_reads == 0 and not windows:
                    return None
                _refuse_unsupported_records(cb.batch.len_flag, contig)
                reference = fasta.fetch_bytes(contig)
                plan = pool.submit(GF.plan_contig, cb, windows, len(reference))   # include/ga_plan.h: no per-read Python object; made while the reads go to the device
                return contig, windows, cb, reference, plan

            n_contigs = len(fasta.references)
            ahead = pool.submit(prepare, 0) if n_contigs else None
            for contig_id in range(n_contigs):
                ready = ahead.result()
                ahead = pool.submit(prepare, contig_id + 1) if contig_id + 1 < n_contigs else None
                if ready is None:
                    continue
                contig, windows, cb, reference, plan = ready
                if cb.batch.n_reads == 0:
                    plan = plan.result()
                    stats_parts.append((contig, plan, [[0, 0, 0, 0]] * len(plan.sessions)))
                    continue
                if not handles:
                    handles.update({k: f.result() for k, f in opening.items()})
                trace(f"mask {contig} begins")
                got = anonymize_packed(engine, cb.batch, (cb.name_blob, cb.name_off), None, windows, reference, contig, plan=plan, as_bytes=True, carry=carry)
                trace(f"text of {contig} ready")
                list(pool.map(lambda k: _append(handles[k], got[f"{k[0]}.{k[1]}"]), list(handles)))   # done before the next contig reuses the buffer
                trace(f"write {contig} done")
                stats_parts.append((contig, got["_plan"], got["_counts"]))
                n_reads += cb.batch.n_reads
                n_sessions += len(got["_plan"].sessions)
    finally:
        handles.update({k: f.result() for k, f in opening.items() if k not in handles and f.exception() is None})
        list(pool.map(lambda h: h.close(), handles.values()))
        pool.shutdown()
        fasta.close()
    if carry:                                                      # write_single_end_reads opens both files (:603-605)
        with open(tumor_output_fastq + ".single_end.fastq", "wb") as t, open(normal_output_fastq + ".single_end.fastq", "wb") as n:
            for _, d, rec in carry.values():                       # insertion order of the collection (:606-622)
                (t if d == 0 else n).write(rec)
    if record_statistics:
        with open(f"{normal_bam_file}.statistics.txt", "w") as fh:
            fh.write(statistics_text(stats_parts))
    logging.info(f"Anonymization complete for samples {tumor_output_fastq} and {normal_output_fastq}")
    return {"reads": n_reads, "sessions": n_sessions}


def run_short_read_tumor_normal_anonymizer(vcf_variants_per_sample: List[str],
                                           tumor_normal_samples: List[Tuple[str, str]],
                                           ref_genome_file: str, anonymizer,
                                           output_filenames: List[Tuple[str, str]], record_statistics: bool,
                                           cpus: int, enhance_parallelization: bool = False):
    """Same signature as the reference's (short_read_tumor_normal_anonymizer.py:889-893).  `anonymizer` is a
    B200GermlineAnonymizer (its engine is used), an engine.Engine, or None (an engine on cuda:0 is created)."""
    from .engine import Engine
    engine: Optional[Engine] = None
    if isinstance(anonymizer, Engine):
        engine = anonymizer
    elif anonymizer is not None and hasattr(anonymizer, "_get_engine"):
        engine = anonymizer._get_engine()
    if engine is None:
        engine = Engine(0)
    fasta = GF.FastaFile(ref_genome_file)
    order = {name: k for k, name in enumerate(fasta.references)}          # get_ref_idxs (:60-63)
    fasta.close()
    results = []
    for vcf, samples, outputs in zip(vcf_variants_per_sample, tumor_normal_samples, output_filenames):
        windows = GF.windows_by_contig(GF.read_vcf(vcf), order)
        results.append(anonymize_genome(windows, samples[DATASET_IDX_TUMORAL], samples[DATASET_IDX_NORMAL], ref_genome_file,
                                        engine, outputs[DATASET_IDX_TUMORAL], outputs[DATASET_IDX_NORMAL],
                                        record_statistics, cpus))
    return results
