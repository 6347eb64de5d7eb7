/*
 * ga_genome_io.h - C ABI of the host-side genome file readers that feed the masking engine
 * (SURVEY.md 8(f) N4): BGZF/BAM -> structure-of-arrays read batch (ga_reads of ga_b200.h), FASTA -> contig
 * bases.  zlib is the only dependency.
 *
 * Replaces, for this path, what the reference does through pysam / htslib:
 *   pysam.AlignmentFile(tumor_bam_file) / (normal_bam_file)      short_read_tumor_normal_anonymizer.py:661-664
 *   AlignmentFile.pileup(reference, start, end, stepper='nofilter', ...)   pileup_io.pyx:12-17
 *   AlignmentFile.fetch(seq_name, start, stop)                    pileup_io.pyx:138-139
 *   pysam.FastaFile(ref_genome_file), .references, .fetch         short_read_tumor_normal_anonymizer.py:915-916,
 *                                                                 variation_classifier.py:89,193
 * The reference walks pysam objects read by read; here a whole contig of a BAM file is decoded by all host
 * threads straight into the arrays ga_run() / ga_run_host() take (seq4 with the LOW nibble first, records padded to
 * 16 bytes, qualities in alignment order, BAM CIGAR words), so no per-read Python object is ever built.
 * Like the reference's 'nofilter' stepper, no record is dropped by flag unless the caller passes flag_exclude.
 *
 * Plain pointers and sizes; the caller owns every output buffer.  Functions return 0 on success, a negative
 * ga_io_status otherwise; ga_io_last_error() gives the text (thread local).
 */
#ifndef GA_GENOME_IO_H
#define GA_GENOME_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ga_io_status {
    GA_IO_OK = 0,
    GA_IO_ERR_OPEN = -1,        /* file missing / unreadable                                   */
    GA_IO_ERR_FORMAT = -2,      /* not BGZF / BAM / FASTA, truncated or corrupt (CRC mismatch) */
    GA_IO_ERR_ARGUMENT = -3,    /* bad reference id, NULL pointer, capacity too small          */
    GA_IO_ERR_UNSUPPORTED = -4  /* reads longer than 65535 bases, CIGAR in a CG tag            */
} ga_io_status;

const char* ga_io_last_error(void);
/* Records an error text for ga_io_last_error() and returns `code`: used by the other host-side sources of the library
 * (ga_plan.h) so that there is one error channel. */
int ga_io_set_error(int code, const char* msg);

/* ------------------------------------------------------------------ BAM */
typedef struct ga_bam ga_bam;

/* Opens a BAM file: inflates every BGZF block (n_threads host threads, 0 = all), checks the block CRCs, parses
 * the header and indexes the alignment records per reference (file order is kept; no .bai is needed).
 * A file whose uncompressed stream exceeds GA_BAM_EAGER_BYTES (environment, default 2 GiB) and whose records are
 * grouped by reference is walked once, window by window, and then kept mapped: the records of one reference at a time
 * are inflated when ga_bam_contig_sizes / ga_bam_pack_contig ask for them (the handle caches one reference, so it is
 * for one thread at a time).  A file that does not begin with the gzip magic is read as SAM text (SAM v1.6 section 1:
 * header dictionary from the @SQ lines, eleven mandatory fields per line; optional fields are not kept). */
int  ga_bam_open(const char* path, int n_threads, ga_bam** out);
void ga_bam_close(ga_bam* b);
int  ga_bam_n_references(const ga_bam* b);
const char* ga_bam_reference_name(const ga_bam* b, int ref_id);
int64_t ga_bam_reference_length(const ga_bam* b, int ref_id);
int64_t ga_bam_n_records(const ga_bam* b);                 /* all alignment records of the file       */
int64_t ga_bam_inflated_bytes(const ga_bam* b);            /* size of the uncompressed BAM stream     */

/* Array sizes needed to pack the records of reference ref_id whose flag has none of the bits of flag_exclude. */
typedef struct ga_bam_sizes {
    int64_t n_reads;
    int64_t seq16_units;     /* 16-byte units of seq4 (qual takes 32 bytes per unit)          */
    int64_t n_cigar;         /* CIGAR words                                                   */
    int64_t name_bytes;      /* sum of name lengths (no terminators)                          */
    int32_t max_ref_span;    /* max(reference_end - reference_start)                          */
    int32_t sorted;          /* 1 when the records are in coordinate order (what ga_run needs) */
} ga_bam_sizes;
int  ga_bam_contig_sizes(const ga_bam* b, int ref_id, uint32_t flag_exclude, ga_bam_sizes* out);

/* Destination arrays of one dataset inside a (possibly larger) batch: read k of the contig goes to index k of
 * pos / len_flag / seq_off16 / ref_end / name_off and cigar_off (which gets n_reads + 1 entries); offsets written
 * are absolute: seq_off16 = seq16_base + ..., cigar_off = cigar_base + ..., name_off = name_base + ...; the
 * record bytes go to seq4 + 16 * seq_off16, qual + 32 * seq_off16, cigar + cigar_off, names + name_off (so the
 * caller passes the batch-wide base pointers and places the normal reads behind the tumor reads by the bases). */
typedef struct ga_bam_dest {
    int32_t*  pos;           /* [n] 0-based reference_start                                   */
    uint32_t* len_flag;      /* [n] flag << 16 | query length                                 */
    uint32_t* seq_off16;     /* [n]                                                           */
    uint32_t* cigar_off;     /* [n + 1]                                                       */
    int32_t*  ref_end;       /* [n] reference_end (pos + reference span of the CIGAR)         */
    uint64_t* name_off;      /* [n + 1]                                                       */
    uint32_t* cigar;         /* batch-wide                                                    */
    uint8_t*  seq4;          /* batch-wide, 16-byte aligned                                   */
    uint8_t*  qual;          /* batch-wide, may be NULL (qualities not wanted)                */
    uint8_t*  names;         /* batch-wide, may be NULL                                       */
    uint32_t  seq16_base;
    uint32_t  cigar_base;
    uint64_t  name_base;
} ga_bam_dest;
int  ga_bam_pack_contig(const ga_bam* b, int ref_id, uint32_t flag_exclude, const ga_bam_dest* dst, int n_threads);

/* ------------------------------------------------------------------ FASTA */
typedef struct ga_fasta ga_fasta;
int  ga_fasta_open(const char* path, ga_fasta** out);      /* FASTA, plain text or gzip / bgzip; sequences in file order */
void ga_fasta_close(ga_fasta* f);
int  ga_fasta_n_references(const ga_fasta* f);
const char* ga_fasta_reference_name(const ga_fasta* f, int idx);
int64_t ga_fasta_reference_length(const ga_fasta* f, int idx);
/* Bases [start, end) of sequence idx as they are in the file (case kept; ga_upload_reference upper-cases),
 * clamped to the sequence like FastaFile.fetch; returns the number of bases written. */
int64_t ga_fasta_fetch(const ga_fasta* f, int idx, int64_t start, int64_t end, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
