/*
 * ga_synth.h - C ABI of the synthetic tumor / normal generator (benchmark and test INPUT only).
 *
 * Lives in its own library, libga_synth.so (genomeanonymizer_b200/csrc/ga_synth.cu), so that a process that only
 * needs input data - bench.py --impl reference, the CPU tests - never maps the masking engine libga_b200.so.
 * Nothing here is on the masking path.
 */
#ifndef GA_SYNTH_H
#define GA_SYNTH_H

#include <stdint.h>
#include "ga_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ device-side synthetic data
 * Benchmark input generator (SURVEY.md 8(d) "Synthetic generator"): counter-based hashing, so any
 * shard of windows can be generated independently on its own GPU and regenerated bit-identically.
 * One session per somatic SNV; reads_per_window[d] = round(cov_d * (2*window_half + read_len) / read_len)
 * reads per dataset and window with stratified sorted start positions; shared germline SNPs / indels on
 * two haplotypes, tumor-only somatic SNV at the window centre, substitution errors, N bases, soft clips.
 * Inputs the reference cannot process are never generated.
 * All functions run on the CURRENT CUDA device, asynchronously on `stream`.
 * Implemented in genomeanonymizer_b200/csrc/ga_synth.cu. */
typedef struct ga_synth_params {
    int64_t contig_len;
    uint64_t seed;
    int32_t read_len;
    int32_t total_windows;      /* windows of the whole contig (fixes the window positions)            */
    int32_t window_begin;       /* this shard generates windows [window_begin, window_begin+n_windows) */
    int32_t n_windows;
    int32_t window_half;        /* 1000 (get_windows window_size/2, SR.py:71)                          */
    int32_t max_indel;          /* germline indel length 1..max_indel                                  */
    int32_t max_clip;           /* soft clip length 1..max_clip                                        */
    int32_t depth_var_pct;      /* 0 = every window at full depth; v > 0: window w of dataset d keeps a fraction drawn uniformly from
                                 * [1 - v/100, 1] of its reads, the others are flagged unmapped (0x4: placed reads, in no session) */
    float   cov_tumor, cov_normal;
    float   snp_rate, indel_rate, err_rate, n_rate, somatic_vaf;
    float   clip_frac;          /* fraction of reads with a soft clip                                  */
} ga_synth_params;

typedef struct ga_synth_plan {
    int64_t n_reads, n_tumor;
    int64_t seq4_bytes;         /* n_reads * 16 * units_per_read                                       */
    int32_t reads_per_window[2];
    int32_t units_per_read;     /* 16-byte seq4 units (32 bases) per record                            */
    int32_t window_stride;
} ga_synth_plan;

int ga_synth_plan_sizes(const ga_synth_params* p, ga_synth_plan* out);
/* ASCII bases of reference positions [begin, begin+n) into d_ascii[0..n). */
int ga_synth_reference(const ga_synth_params* p, uint8_t* d_ascii, int64_t begin, int64_t n, void* stream);
/* Session table of the shard (device arrays sized n_windows, keep_allele_off n_windows+1, keep_alleles n_windows+1). */
int ga_synth_sessions(const ga_synth_params* p, int32_t* first, int32_t* last, int32_t* keep_type, int32_t* keep_pos,
                      int32_t* keep_end, int32_t* keep_len, uint32_t* keep_allele_off, uint8_t* keep_alleles, void* stream);
/* Pass 1: per read the number of CIGAR ops and whether it has an I/D op; *max_ref_span (device int32, zeroed by the caller). */
int ga_synth_reads_count(const ga_synth_params* p, uint32_t* n_ops, uint8_t* has_indel, int32_t* max_ref_span, void* stream);
/* Pass 2: fill the batch.  dst->cigar_off (exclusive scan of n_ops, n_reads+1 entries) is an INPUT; every
 * other array of dst is written.  qual_slot[r] = index of read r's quality record (32*units_per_read bytes
 * each, record k at byte 32*units_per_read*k) or -1 for reads without one; dst->qual_reads / qual_off16
 * are not touched (the caller derives them from qual_slot). */
int ga_synth_reads_fill(const ga_synth_params* p, const ga_reads* dst, const int32_t* qual_slot, void* stream);
/* Host twins: identical arithmetic over HOST pointers (single-threaded; for CPU tests of small shapes). */
int ga_synth_reference_host(const ga_synth_params* p, uint8_t* ascii, int64_t begin, int64_t n);
int ga_synth_sessions_host(const ga_synth_params* p, int32_t* first, int32_t* last, int32_t* keep_type, int32_t* keep_pos,
                           int32_t* keep_end, int32_t* keep_len, uint32_t* keep_allele_off, uint8_t* keep_alleles);
int ga_synth_reads_count_host(const ga_synth_params* p, uint32_t* n_ops, uint8_t* has_indel, int32_t* max_ref_span);
int ga_synth_reads_fill_host(const ga_synth_params* p, const ga_reads* dst, const int32_t* qual_slot);

#ifdef __cplusplus
}
#endif
#endif /* GA_SYNTH_H */
