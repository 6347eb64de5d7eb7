/*
 * ga_plan.h - C ABI of the host-side plan of one contig of a tumor-normal sample (SURVEY.md 8(f) N2 + N3).
 *
 * The reference walks a sample section by section (anonymize_genome, short_read_tumor_normal_anonymizer.py:625-760):
 * variant windows are sessions, the regions between them are fetched and passed through unless a tumor and a normal
 * read island overlap (one more session: iter_fetch_pair, pileup_io.pyx:124-298 - compiled Cython in the reference),
 * mates meet in to_pair_anonymized_reads, and a pair is written once, by the first section that completes it
 * (written_read_ids, :134-165).  None of it needs the bases: ga_plan_sample works on (name, flag, start, end) of the
 * packed batch (the arrays ga_bam_pack_contig writes) and returns the session table (windows and island sessions in
 * processing order) and the write plan: which read, which session's version (-1 = as it came in), in file order,
 * including the yield order inside a session (anonymizer_methods.py:472-532).
 * genomeanonymizer_b200/driver.py: plan_sample is the same algorithm in Python (kept as the checker of this one).
 *
 * Plain pointers and sizes; results are copied into caller-owned arrays.  Errors: negative ga_io_status
 * (ga_genome_io.h) with the text in ga_io_last_error().
 */
#ifndef GA_PLAN_H
#define GA_PLAN_H

#include <stdint.h>

#define GA_PLAN_REAPPLY (1 << 30)

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ga_plan ga_plan;

/* reads: tumor reads [0, n_tumor) then normal reads, each dataset in coordinate order.  pos / ref_end: 0-based start
 * and exclusive end; len_flag: BAM flag << 16 | anything; name_off[n_reads + 1] into names.  windows sorted by
 * (first, last) as the reference sorts them.  GA_IO_ERR_ARGUMENT when a dataset is not sorted or when two windows
 * are closer than a window (the reference's fetch coordinates would be rejected by pysam). */
int  ga_plan_sample(int64_t n_reads, int64_t n_tumor, const int32_t* pos, const int32_t* ref_end, const uint32_t* len_flag,
                    const int64_t* name_off, const uint8_t* names, int32_t n_windows, const int32_t* win_first,
                    const int32_t* win_last, int64_t contig_len, ga_plan** out);
void ga_plan_free(ga_plan* p);

int64_t ga_plan_n_sessions(const ga_plan* p);
int64_t ga_plan_n_pairs(const ga_plan* p);
int64_t ga_plan_n_singles(const ga_plan* p);
/* sessions: first / last / window index (-1: island session between windows), each [n_sessions]. */
void ga_plan_sessions(const ga_plan* p, int32_t* first, int32_t* last, int32_t* window);
/* A version is the session whose masking the read prints, or -1 (as it came in); bit 30 (GA_PLAN_REAPPLY) on a version says
 * that the reference applies the read's left-over indels a second time (quirk Q12 of DESIGN.md: the read waited unpaired with
 * that masking and was met again by a later session).
 * pairs: rows of 5 - dataset, read of mate 1, its session version, read of mate 2, its session version - in write order;
 * singles: rows of 4 - dataset, read, version, and the number of pairs of this contig that precede the place where a pair
 * completed by this read is written (the reference keeps unpaired reads across contigs, to_pair_anonymized_reads, and
 * writes such a pair the moment the second mate is processed) - in the order the reference spills them to the
 * single-end files. */
void ga_plan_pairs(const ga_plan* p, int32_t* rows);
void ga_plan_singles(const ga_plan* p, int32_t* rows);

#ifdef __cplusplus
}
#endif
#endif
