/*
 * ga_b200.h - C ABI of the B200-native per-read germline-masking engine.
 *
 * Drop-in boundary for the hot path of Computational-Genomics-BSC/GenomeAnonymizer:
 * one call of CompleteGermlineAnonymizer.anonymize() per pileup region ("session",
 * reference src/GenomeAnonymizer/anonymizer_methods.py:431-535, called from
 * short_read_tumor_normal_anonymizer.py:289-293) becomes one row of a session table, and a
 * whole contig's worth of sessions is processed by one ga_run() launch sequence.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.  The caller owns
 * every buffer (SURVEY.md 8(b) "Ownership"); the engine owns its handle and scratch.
 * Every function returns a ga_status; no exception crosses the ABI.
 *
 * The same structs (with HOST pointers) are consumed by the CPU oracle in oracle/ga_oracle.c,
 * which is test infrastructure and not part of this library.
 */
#ifndef GA_B200_H
#define GA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GA_ABI_VERSION 1

/* ------------------------------------------------------------------ status codes
 * Reference error behaviour (SURVEY.md 8(b) "Errors"): ValueError (AM.py:144,201,248,256),
 * IndexError from np.put(mode='raise') (AM.py:174).  The Python wrapper maps
 * GA_ERR_BAD_ARGUMENT/GA_ERR_LENGTH_MISMATCH -> ValueError, GA_ERR_OFFSET_RANGE -> IndexError,
 * everything else -> RuntimeError. */
typedef enum ga_status {
    GA_OK = 0,
    GA_ERR_BAD_ARGUMENT = 1,
    GA_ERR_OFFSET_RANGE = 2,     /* a query offset / read span fell outside its record or table   */
    GA_ERR_LENGTH_MISMATCH = 3,  /* sequence / quality length inconsistency (AM.py:200-201)        */
    GA_ERR_CUDA = 4,
    GA_ERR_CAPACITY = 5,         /* caller-provided output capacity too small; totals say how much */
    GA_ERR_UNSUPPORTED = 6,      /* e.g. more than GA_MAX_EDITS germline indels in one read        */
    GA_ERR_NO_DEVICE = 7
} ga_status;

/* BAM 4-bit base codes (=ACMGRSVTWYHKDBN -> 0..15): A=1 C=2 G=4 T=8 N=15. */
/* BAM CIGAR op codes: M=0 I=1 D=2 N=3 S=4 H=5 P=6 '='=7 X=8 (B=9 ignored). */

#define GA_DATASET_TUMOR 0   /* variation_classifier.py:13 */
#define GA_DATASET_NORMAL 1  /* variation_classifier.py:14 */

/* VariantType values as in variant_extractor (statistics column = value-1, SR.py:200,218-219). */
#define GA_VT_NONE 0
#define GA_VT_SNV 1
#define GA_VT_DEL 2
#define GA_VT_INS 3

#define GA_MAX_EDITS 32      /* germline indels applied to one read (documented limit) */

/* ------------------------------------------------------------------ input: reads of one contig
 * Structure-of-arrays batch of the aligned reads of ONE contig (or contig chunk) that overlap at
 * least one session, tumor reads [0,n_tumor) then normal reads [n_tumor,n_reads), each dataset in
 * coordinate (BAM file) order.  Replaces the pysam objects crossing the reference boundary
 * (SURVEY.md 8(b) "Objects crossing it"): query_sequence, query_qualities, cigarstring,
 * reference_start, flag.  Read names stay on the host; a read is identified by its index.
 *
 * seq4: 4-bit BAM base codes, base k of a read in bits [4*(k&1), 4*(k&1)+4) of byte k/2
 *       (LOW nibble first - note BAM itself stores the high nibble first; the packer swaps).
 *       Record r starts at byte 16*seq_off16[r]; record capacity is a multiple of 16 bytes.
 *       The array itself (and qual, out_seq4, out_qual) must be 16-byte aligned: records are moved with
 *       128-bit accesses and TMA bulk copies.
 * qual: phred bytes in BAM (alignment) order; record r starts at byte 32*seq_off16[r].
 *       May be NULL together with qual_reads==NULL only if no read has an I or D op.
 * qual_reads: optional sparse quality upload.  NULL = every read has a quality record at
 *       32*seq_off16[r].  Otherwise a sorted list of n_qual read indices that have quality records,
 *       record k at byte 32*qual_off16[k] (reads without I/D ops can never change qualities).
 */
typedef struct ga_reads {
    int64_t n_reads;
    int64_t n_tumor;
    const int32_t*  pos;        /* [n_reads]   0-based reference_start                               */
    const uint32_t* len_flag;   /* [n_reads]   (BAM flag << 16) | query length                       */
    const uint32_t* seq_off16;  /* [n_reads]                                                         */
    const uint32_t* cigar_off;  /* [n_reads+1] index of the first CIGAR op                           */
    const uint32_t* cigar;      /* [cigar_off[n_reads]] BAM encoding len<<4|op                       */
    const uint8_t*  seq4;
    const uint8_t*  qual;
    int64_t seq4_bytes;         /* total size of seq4 (multiple of 16)                               */
    int64_t n_qual;             /* sparse quality upload: number of entries, 0 when qual_reads==NULL */
    const int32_t*  qual_reads;
    const uint32_t* qual_off16;
    int32_t max_ref_span;       /* >= max over reads of (reference_end - reference_start); 0 = let the engine compute */
    int32_t contig_id;          /* which uploaded reference these reads align to                     */
} ga_reads;

/* ------------------------------------------------------------------ input: session table
 * One row per pileup region in processing (genome) order, sorted by (first, last):
 * variant windows from get_windows() (SR.py:71-131) and, later, island regions (PIO.pyx:124-298).
 * A read belongs to session s iff pos < last[s] and reference_end > first[s] (pileup fetch,
 * PIO.pyx:12-17, truncate=False).  Evidence AND output are per session: a read that lies in two
 * sessions (adjacent or overlapping windows) is masked independently in each, exactly as two
 * anonymize() calls would, and yields one output record per session that modified it.  Which of
 * those versions reaches the FASTQ is the writer's pairing / first-write-wins policy
 * (SR.py:134-165, 304-360), which depends on where the mates lie - see DESIGN.md "Writer policy".
 * keep_*: the validated somatic variant that must not be masked (AM.py:546-547), compared with
 * CalledGenomicVariant.__eq__ (variants.py:83-96) on (type,pos,end,length,allele); contig equality
 * is implied by the batch.  keep_type==GA_VT_NONE: nothing kept.
 */
typedef struct ga_sessions {
    int32_t n_sessions;
    const int32_t* first;            /* [S] region start, may be negative                             */
    const int32_t* last;             /* [S] region stop (exclusive)                                   */
    const int32_t* keep_type;        /* [S]                                                           */
    const int32_t* keep_pos;         /* [S] 0-based                                                   */
    const int32_t* keep_end;         /* [S] 0-based                                                   */
    const int32_t* keep_len;         /* [S]                                                           */
    const uint32_t* keep_allele_off; /* [S+1] into keep_alleles                                       */
    const uint8_t* keep_alleles;     /* ASCII, as in the VCF ALT column (not upper-cased)             */
} ga_sessions;

/* ------------------------------------------------------------------ output
 * Compacted records of the MODIFIED (session, read) pairs only (stream compaction for the writer,
 * north_star job 4).  Record k (0 <= k < totals->n_modified; records of one session are contiguous
 * and ascending in read index, the order of sessions is unspecified):
 *   mod_session[k]   session index in the table
 *   mod_read[k]      read index in the batch
 *   mod_len[k]       new query length (changes only when indels were masked)
 *   mod_seq_off16[k] new seq4 record at byte 16*off in out_seq4 (same nibble layout as the input)
 *   mod_qual_off16[k] 0xFFFFFFFF if qualities are unchanged (SNV-only masking, AM.py:170-176),
 *                    else new quality record at byte 32*off in out_qual, already in FASTQ (printed)
 *                    order, i.e. after the reference's reversal for reverse reads (AM.py:213,233).
 * sess_counts[4*s + {0,1,2}] = germline SNV / DEL / INS variants masked in session s
 *   (AnonymizedVariantsStatistics.window_var_counts, SR.py:198-204); [4*s+3] = session reads (T+N).
 */
typedef struct ga_totals {
    uint64_t n_modified;
    uint64_t seq16_used;      /* 16-byte units consumed in out_seq4   */
    uint64_t qual16_used;     /* 32-byte units consumed in out_qual   */
    uint64_t session_reads;   /* sum over sessions of reads processed */
    uint64_t session_bases;   /* sum of their query lengths           */
    uint64_t indel_records;   /* records that carry a quality record   */
    uint64_t masked[3];       /* SNV, DEL, INS totals (sum of sess_counts) */
    uint32_t error;           /* first ga_status raised on the device, 0 if none */
    uint32_t error_detail;    /* read or session index that raised it  */
} ga_totals;

typedef struct ga_result {
    int64_t cap_records;
    int64_t cap_seq16;
    int64_t cap_qual16;
    int32_t*  mod_session;
    int32_t*  mod_read;
    uint32_t* mod_len;
    uint32_t* mod_seq_off16;
    uint32_t* mod_qual_off16;
    uint8_t*  out_seq4;
    uint8_t*  out_qual;
    uint32_t* sess_counts;    /* [4*n_sessions] */
    ga_totals* totals;        /* one struct, same memory space as the other pointers */
} ga_result;

/* ------------------------------------------------------------------ engine (CUDA, sm_100a) */
typedef struct ga_engine ga_engine;

int  ga_abi_version(void);
const char* ga_status_string(int status);

/* Create an engine bound to CUDA device `device`.  Fails with GA_ERR_NO_DEVICE when no usable
 * sm_100 device exists - there is no CPU fallback. */
int  ga_engine_create(int device, ga_engine** out);
void ga_engine_destroy(ga_engine* e);
const char* ga_last_error(const ga_engine* e);

/* Upload one reference contig (ASCII, any case; replaces FastaFile.fetch at
 * variation_classifier.py:89,193).  `bases` may be a host or a device pointer. The engine keeps a
 * 4-bit upper-cased copy resident in HBM until destroy or re-upload of the same contig_id. */
int  ga_upload_reference(ga_engine* e, int contig_id, const uint8_t* bases, int64_t n_bases, void* stream);

/* Run discovery + masking + compaction for every session of the table over a DEVICE-resident batch.
 * All pointers inside reads/sessions/result are device pointers (result->totals too).
 * Asynchronous on `stream` (a cudaStream_t, may be NULL = default stream). */
int  ga_run(ga_engine* e, const ga_reads* reads, const ga_sessions* sessions, ga_result* result, void* stream);

/* Number of engine kernels launched since creation (bench.py's gpu_launches claim). */
int64_t ga_launch_count(const ga_engine* e);

/* Duration in milliseconds of the session kernel in the most recent ga_run() on this engine, measured
 * with CUDA events on the launching stream (synchronises that stream). */
float ga_last_kernel_ms(ga_engine* e);
/* Durations (ms) of the whole masking pass (scan + resolve + fallback + emission kernels) of the most recent
 * ga_run() calls, out[0] = latest; returns how many were written (at most min(n, 32)).  Synchronises on the
 * recorded events. */
int   ga_kernel_ms_history(ga_engine* e, float* out, int n);
/* Same per stage: 0 = scan kernel (allele discovery), 1 = resolve kernels (germline set, record list, headers),
 * 2 = emission kernel (record bodies), 3 = what remains of the fallback kernel for oversize sessions, which runs
 * beside the emission kernel, once that has finished. */
int   ga_stage_ms_history(ga_engine* e, int stage, float* out, int n);
/* Sessions of the most recent ga_run() that took the global-scratch fallback kernel (-1 on error), and why:
 * reasons[0] oversize session, [1] scan-kernel table overflow, [2] IUPAC read base, [3] more modified reads or
 * germline alleles than the shared-memory tables hold, [4] a read with more than two germline indels;
 * reasons[8] = sessions resolved by the one-CTA resolve kernel, reasons[9] = sessions the lean one-warp resolve kernel
 * handed to its larger instantiation (neither is a fallback).
 * Synchronises the device. */
int   ga_last_fallback_sessions(ga_engine* e, int32_t* reasons, int n_reasons);

/* Edit descriptions of indel-masked records, for a host that has to reproduce quirk Q12 of the reference (a read
 * that was masked, parked unpaired and met again by a later session has its left-over indels applied a second time,
 * anonymizer_methods.py:254-287).  ga_engine_keep_edits(e, 1) makes the following runs keep, per modified record, what was
 * applied to it; ga_record_edits then copies 8 words per requested record index (indices into the mod_* arrays of the last
 * ga_run on this engine): { in_read_pos 0, reference position 0, length 0 | INS << 31, in_read_pos 1, reference position 1,
 * length 1 | INS << 31, edits | DELs << 8, - } in application order (all DELs, then all INSs).  All bits set: nothing kept
 * (a record without indel edits, or one with more than two). */
int   ga_engine_keep_edits(ga_engine* e, int on);
int   ga_record_edits(ga_engine* e, const int64_t* rec_idx, int64_t n, uint32_t* out);

/* ------------------------------------------------------------------ end-to-end host entry
 * ga_run_host: all pointers are HOST pointers (pinned for full speed).  Splits the session table into
 * chunks of about `chunk_sessions` sessions (<=0: engine default), slices the reads each chunk needs,
 * overlaps H2D / kernels / D2H on double-buffered streams and appends the chunk results to the
 * caller's host result (mod_session / mod_read are table / batch indices, offsets are rebased).
 * Synchronous.  result->totals->error carries the first device-side status.
 * Precondition (ga_run has none): inside each dataset the seq4 records - and the sparse quality records - lie in read
 * order without overlap (seq_off16 / qual_off16 ascending with the read index), so that the records of a chunk are one
 * contiguous slice; a batch laid out otherwise is refused with GA_ERR_BAD_ARGUMENT. */
int  ga_run_host(ga_engine* e, const ga_reads* reads, const ga_sessions* sessions, ga_result* result,
                 int64_t chunk_sessions);

/* Bytes moved host->device / device->host by the most recent ga_run_host() (bench.py's e2e line). */
void ga_last_host_traffic(const ga_engine* e, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* ------------------------------------------------------------------ FASTQ rendering
 * The step right after the masking path: AnonymizedRead.get_anonymized_fastq_record (anonymizer_methods.py:205-243,
 * 57-58) plus the writer's newline (short_read_tumor_normal_anonymizer.py:157-158), rendered on the device:
 *     "@" query_name "/" (1 if flag & 0x40 else 2) "\n" SEQ "\n+\n" QUAL "\n"
 * Reverse reads (flag & 0x10) are reverse-complemented; qualities are printed in BAM order.
 * Item k renders read read[k] of `reads`; when record[k] >= 0 its sequence and length (and its qualities, when the
 * record has a quality slot) come from record record[k] of `result` (a modified record of a ga_run), otherwise the
 * read is rendered as it came in.  Every rendered read needs a quality record in `reads` (dense upload, or listed
 * in qual_reads) unless its result record carries one.  All pointers are DEVICE pointers.
 * Implemented in genomeanonymizer_b200/csrc/ga_fastq.cu. */
typedef struct ga_fastq_items {
    int64_t n_items;
    const int32_t* read;        /* [n_items] read index                                              */
    const int32_t* record;      /* [n_items] index of the modified record that replaces it, or -1    */
    const uint8_t* names;       /* concatenated query names, no terminators                          */
    const int64_t* name_off;    /* [n_reads + 1] name of read r = names[name_off[r] .. name_off[r+1]) */
} ga_fastq_items;

/* text_off[n_items + 1] (device, OUTPUT): byte offset of every record in the text, text_off[n_items] = total size.
 * result may be NULL when every record[k] is -1; n_records = number of valid records in result. */
int  ga_fastq_layout(ga_engine* e, const ga_reads* reads, const ga_result* result, int64_t n_records,
                     const ga_fastq_items* items, int64_t* text_off, void* stream);
/* Writes the text (device buffer of text_cap bytes) at the offsets ga_fastq_layout produced.  status->error /
 * error_detail (device) receive the first device-side failure (bad index, read without qualities, capacity). */
int  ga_fastq_render(ga_engine* e, const ga_reads* reads, const ga_result* result, int64_t n_records,
                     const ga_fastq_items* items, const int64_t* text_off, uint8_t* text, int64_t text_cap,
                     ga_totals* status, void* stream);

/* ------------------------------------------------------------------ result digest (parity checks at scale)
 * Key and 128-bit hash of every record of a DEVICE-resident result, and their order-independent sum; the
 * arithmetic is specified in include/ga_digest.h.  `ids` places the result in the unsharded input: which global
 * session its session 0 is, and the ordinals of its first tumor / first normal read inside their datasets.
 *   rec_keys[2k] = global session, rec_keys[2k+1] = read gid      (device, may be NULL)
 *   rec_hash[2k], rec_hash[2k+1] = hash lo / hi of record k       (device, may be NULL)
 *   digest[0..3] += {sum lo, sum hi, records, sum of new lengths} (device, NOT cleared: shards accumulate)
 * Asynchronous on `stream`.  Implemented in genomeanonymizer_b200/csrc/ga_wire.cu. */
typedef struct ga_digest_ids {
    int64_t session_base;
    int64_t tumor_base;
    int64_t normal_base;
    int64_t n_tumor;        /* reads [0, n_tumor) of the batch the result refers to are tumor reads */
    int64_t contig;
} ga_digest_ids;

int  ga_result_digest(ga_engine* e, const ga_result* result, int64_t n_records, const ga_digest_ids* ids,
                      uint64_t* rec_keys, uint64_t* rec_hash, uint64_t* digest, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GA_B200_H */
