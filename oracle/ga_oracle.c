/*
 * ga_oracle.c - CPU restatement of GenomeAnonymizer's per-session germline masking.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this file.  The product (libga_b200.so) never
 * links or calls it and has no CPU fallback.
 *
 * Parity status: PINNED against outputs of the reference itself.  The reference has no tests or
 * golden vectors of its own, so tests/golden/make_golden.py executes the reference's
 * unmodified Python modules under stub pysam/variant_extractor and commits the results
 * (tests/golden/ JSON files); tests/test_oracle_golden.py checks this file against every one of them.
 * What stays unpinned is htslib's pileup arithmetic (not installed): the rules used are the
 * documented ones restated in tests/ref_stub/pysam.py.
 *
 * Formulation: the reference streams pileup columns (anonymizer_methods.py:440-534).  This file
 * uses the equivalent two-phase per-read form (SURVEY.md 8(a) "Equivalence"):
 *   1. discover  every aligned base that differs from an A/C/G/T reference base and is not N is an
 *                SNV allele (variation_classifier.py:144-150); every I/D CIGAR op is an indel allele
 *                keyed by (type,pos,len,read bases) (variation_classifier.py:52-107,
 *                variants.py:83-96); each key remembers which datasets showed it (the T/N state
 *                machine variation_classifier.py:163-182 collapses to "seen in T" | "seen in N").
 *   2. germline  keys seen in both datasets, minus the window's variant_to_keep
 *                (anonymizer_methods.py:546-547), at positions the normal pileup has a column for
 *                (the reference masks a position's variants when it reaches that column,
 *                anonymizer_methods.py:474-481: matters for an insertion that ends its read).
 *   3. mask      SNV: base <- reference base, quality untouched (anonymizer_methods.py:170-176);
 *                indels: all DELs then all INSs at original, unadjusted offsets
 *                (anonymizer_methods.py:254-270, 178-203), edits literally simulated on arrays.
 * Output is per (session, read): a read lying in two sessions is masked independently in each, as two
 * anonymize() calls would; the writer's pairing policy (SR.py:134-165, 304-360) picks the version.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>
#include "../include/ga_b200.h"
#include "../include/ga_digest.h"

/* ASCII -> BAM 4-bit code ("=ACMGRSVTWYHKDBN"), case-insensitive; anything else -> 15 (N). */
static uint8_t g_asc2code[256];
static const char g_code2asc[17] = "=ACMGRSVTWYHKDBN";
static void init_tables(void) {
    static int done = 0;
    if (done) return;
    memset(g_asc2code, 15, sizeof g_asc2code);
    for (int c = 0; c < 16; ++c) {
        char ch = g_code2asc[c];
        g_asc2code[(unsigned char)ch] = (uint8_t)c;
        if (ch >= 'A' && ch <= 'Z') g_asc2code[(unsigned char)(ch + 32)] = (uint8_t)c;
    }
    done = 1;
}

static inline int is_acgt_code(int c) { return c == 1 || c == 2 || c == 4 || c == 8; }

static inline int nib_at(const uint8_t* rec, int k) {
    uint8_t b = rec[k >> 1];
    return (k & 1) ? (b >> 4) : (b & 15);
}

typedef struct {
    int type, pos, len;   /* GA_VT_DEL / GA_VT_INS, 0-based reference pos, op length          */
    int irp;              /* in_read_pos with the reference's H/N quirk (VC.py:82, Q7)         */
    int alen;             /* allele length after Python-slice truncation                        */
    int read;             /* batch read index                                                   */
    int ds;
    int key;              /* index of the unique key this observation maps to                   */
} obs_t;

typedef struct {
    int type, pos, len, alen;
    int rep_obs;          /* observation holding the allele bytes                               */
    int mask;             /* bit0 tumor, bit1 normal                                            */
    int next;             /* chain of keys at the same column                                   */
    int germline;
} ikey_t;

typedef struct {
    int read;
    int newlen;
    uint8_t* seq;         /* codes, one per byte, newlen                                        */
    uint8_t* qual;        /* printed order, newlen; NULL when qualities unchanged               */
    uint32_t aux[8];     /* the record's edits in the form of ga_record_edits (include/ga_b200.h), all ones: none kept */
} rres_t;

typedef struct { rres_t* v; int n, cap; } sres_t;   /* modified reads of one session, ascending read index */

typedef struct {
    obs_t* obs; int n_obs, cap_obs;
    ikey_t* keys; int n_keys, cap_keys;
    uint32_t* snv; int* khead; uint8_t* ncov; int cap_cols;   /* ncov[col]: a normal read of the session covers the column */
    uint8_t* sbuf; uint8_t* qbuf; int cap_buf;
} scratch_t;

static int ref_end_of(const ga_reads* R, int64_t r) {
    int e = R->pos[r];
    for (uint32_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; ++c) {
        uint32_t op = R->cigar[c] & 15, ln = R->cigar[c] >> 4;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) e += (int)ln;
    }
    return e;
}

/* htslib's pileup drops BAM_FUNMAP records: a placed-unmapped mate is in no session */
static int is_unmapped(const ga_reads* R, int64_t r) { return (int)((R->len_flag[r] >> 16) & 0x4u); }

static int64_t lower_bound_pos(const int32_t* pos, int64_t b, int64_t e, int64_t v) {
    while (b < e) { int64_t m = (b + e) >> 1; if ((int64_t)pos[m] < v) b = m + 1; else e = m; }
    return b;
}

static const uint8_t* qual_record(const ga_reads* R, int64_t r) {
    if (!R->qual) return NULL;
    if (!R->qual_reads) return R->qual + 32ull * R->seq_off16[r];
    int64_t b = 0, e = R->n_qual;
    while (b < e) { int64_t m = (b + e) >> 1; if (R->qual_reads[m] < r) b = m + 1; else e = m; }
    if (b < R->n_qual && R->qual_reads[b] == r) return R->qual + 32ull * R->qual_off16[b];
    return NULL;
}

static int obs_allele_equal(const ga_reads* R, const obs_t* a, const obs_t* b) {
    if (a->alen != b->alen) return 0;
    const uint8_t* ra = R->seq4 + 16ull * R->seq_off16[a->read];
    const uint8_t* rb = R->seq4 + 16ull * R->seq_off16[b->read];
    for (int j = 0; j < a->alen; ++j)
        if (nib_at(ra, a->irp + j) != nib_at(rb, b->irp + j)) return 0;
    return 1;
}

static int key_equals_keep(const ga_reads* R, const ga_sessions* S, int s, const ikey_t* k, const obs_t* obs) {
    if (S->keep_type[s] != k->type || S->keep_pos[s] != k->pos || S->keep_len[s] != k->len) return 0;
    int end = (k->type == GA_VT_INS) ? k->pos + 1 : k->pos + k->len - 1;   /* VC.py:86 */
    if (S->keep_end[s] != end) return 0;
    uint32_t a0 = S->keep_allele_off[s], a1 = S->keep_allele_off[s + 1];
    if ((int)(a1 - a0) != k->alen) return 0;
    const obs_t* o = &obs[k->rep_obs];
    const uint8_t* rec = R->seq4 + 16ull * R->seq_off16[o->read];
    for (int j = 0; j < k->alen; ++j)
        if (S->keep_alleles[a0 + j] != (uint8_t)g_code2asc[nib_at(rec, o->irp + j)]) return 0;
    return 1;
}

#define GROW(ptr, cap, need, T) do { if ((need) > (cap)) { (cap) = (need) * 2 + 16; (ptr) = (T*)realloc((ptr), sizeof(T) * (size_t)(cap)); } } while (0)

/* Quirk Q12 of DESIGN.md (per-sample orchestration): (session << 32 | read) keys, ascending, of the reads whose left-over
 * indels the reference applies a SECOND time (anonymizer_methods.py:254-287: the read was masked, parked unpaired and met again
 * by a later session, which switches has_left_overs_to_mask on again over the list that was never cleared).  Set by the
 * whole-sample checkers before ga_oracle_run; empty for everything else. */
static const int64_t* g_reapply = NULL;
static int64_t g_n_reapply = 0;
void ga_oracle_set_reapply(const int64_t* keys, int64_t n) { g_reapply = keys; g_n_reapply = n; }
/* Checker twin of ga_record_edits (include/ga_b200.h): 8 words per record of the next runs, in record order. */
static uint32_t* g_edit_sink = NULL;
static int64_t g_edit_cap = 0;
void ga_oracle_set_edit_sink(uint32_t* out8, int64_t cap_records) { g_edit_sink = out8; g_edit_cap = cap_records; }
static int reapply_twice(int s, int64_t r) {
    const int64_t key = ((int64_t)s << 32) | (int64_t)r;
    int64_t lo = 0, hi = g_n_reapply;
    while (lo < hi) { const int64_t m = (lo + hi) >> 1; if (g_reapply[m] < key) lo = m + 1; else hi = m; }
    return lo < g_n_reapply && g_reapply[lo] == key;
}

static int process_session(const ga_reads* R, const ga_sessions* S, int s, const uint8_t* refc, int64_t ref_len,
                           int maxspan, scratch_t* W, sres_t* res, uint32_t* counts,
                           uint64_t* tot_reads, uint64_t* tot_bases, uint32_t* err_detail) {
    const int first = S->first[s], last = S->last[s];
    int64_t rng[2][2];
    rng[0][0] = lower_bound_pos(R->pos, 0, R->n_tumor, (int64_t)first - maxspan + 1);
    rng[0][1] = lower_bound_pos(R->pos, 0, R->n_tumor, last);
    rng[1][0] = lower_bound_pos(R->pos, R->n_tumor, R->n_reads, (int64_t)first - maxspan + 1);
    rng[1][1] = lower_bound_pos(R->pos, R->n_tumor, R->n_reads, last);
    /* column span of the session = [min start, max end) over its reads */
    int col_lo = 0x7fffffff, col_hi = -0x7fffffff;
    for (int d = 0; d < 2; ++d)
        for (int64_t r = rng[d][0]; r < rng[d][1]; ++r) {
            int e = ref_end_of(R, r);
            if (e <= first || is_unmapped(R, r)) continue;
            if (R->pos[r] < col_lo) col_lo = R->pos[r];
            if (e > col_hi) col_hi = e;
        }
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
    if (col_hi <= col_lo) return GA_OK;
    int n_cols = col_hi - col_lo + 1;          /* +1: an insertion at the very end of a read sits at col == end */
    if (n_cols > W->cap_cols) {
        W->cap_cols = n_cols * 2;
        W->snv = (uint32_t*)realloc(W->snv, sizeof(uint32_t) * (size_t)W->cap_cols);
        W->khead = (int*)realloc(W->khead, sizeof(int) * (size_t)W->cap_cols);
        W->ncov = (uint8_t*)realloc(W->ncov, (size_t)W->cap_cols);
    }
    memset(W->snv, 0, sizeof(uint32_t) * (size_t)n_cols);
    memset(W->ncov, 0, (size_t)n_cols);
    for (int i = 0; i < n_cols; ++i) W->khead[i] = -1;
    W->n_obs = 0; W->n_keys = 0;

    /* ---------------- phase 1: discover ---------------- */
    uint32_t n_sess_reads = 0;
    for (int d = 0; d < 2; ++d)
        for (int64_t r = rng[d][0]; r < rng[d][1]; ++r) {
            if (ref_end_of(R, r) <= first || is_unmapped(R, r)) continue;
            ++n_sess_reads;
            const int L = (int)(R->len_flag[r] & 0xffff);
            *tot_bases += (uint64_t)L;
            const uint8_t* rec = R->seq4 + 16ull * R->seq_off16[r];
            int rc = R->pos[r], q = 0, ccl = 0, rcb = 0;
            if (d == 1)                                           /* the columns of the normal pileup: [reference_start, reference_end) */
                for (int p = R->pos[r], e = ref_end_of(R, r); p < e; ++p) W->ncov[p - col_lo] = 1;
            for (uint32_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; ++c) {
                int op = (int)(R->cigar[c] & 15), ln = (int)(R->cigar[c] >> 4);
                if (op == 0 || op == 7 || op == 8) {              /* M = X */
                    for (int k = 0; k < ln; ++k) {
                        if (q + k >= L) { *err_detail = (uint32_t)r; return GA_ERR_OFFSET_RANGE; }  /* IndexError in VC.py:148 */
                        int64_t p = (int64_t)rc + k;
                        int rf = (p >= 0 && p < ref_len) ? refc[p] : 15;
                        int b = nib_at(rec, q + k);
                        if (b != 15 && b != rf && is_acgt_code(rf))
                            W->snv[p - col_lo] |= 1u << (b + 16 * d);
                    }
                    q += ln; rc += ln; ccl += ln;
                } else if (op == 1 || op == 2) {                   /* I / D : VC.py:72-107 */
                    GROW(W->obs, W->cap_obs, W->n_obs + 1, obs_t);
                    obs_t* o = &W->obs[W->n_obs];
                    o->type = (op == 1) ? GA_VT_INS : GA_VT_DEL;
                    o->pos = rc; o->len = ln; o->irp = ccl + rcb; o->read = (int)r; o->ds = d;
                    int want = (op == 1) ? ln : 2;                 /* VC.py:87-88 */
                    int avail = L - o->irp; if (avail < 0) avail = 0;
                    o->alen = want < avail ? want : avail;
                    /* find or create the key */
                    int col = rc - col_lo, kfound = -1;
                    for (int k = W->khead[col]; k >= 0; k = W->keys[k].next) {
                        ikey_t* kk = &W->keys[k];
                        if (kk->type == o->type && kk->len == o->len && obs_allele_equal(R, &W->obs[kk->rep_obs], o)) { kfound = k; break; }
                    }
                    if (kfound < 0) {
                        GROW(W->keys, W->cap_keys, W->n_keys + 1, ikey_t);
                        ikey_t* kk = &W->keys[W->n_keys];
                        kk->type = o->type; kk->pos = o->pos; kk->len = o->len; kk->alen = o->alen;
                        kk->rep_obs = W->n_obs; kk->mask = 0; kk->germline = 0;
                        kk->next = W->khead[col]; W->khead[col] = W->n_keys; kfound = W->n_keys++;
                    }
                    W->keys[kfound].mask |= 1 << d;
                    o->key = kfound;
                    W->n_obs++;
                    if (op == 1) { q += ln; rcb += ln; } else { rc += ln; ccl += ln; rcb -= ln; }
                } else if (op == 3) { rc += ln; ccl += ln; }       /* N: ref-consuming, not subtracted (Q7) */
                else if (op == 4) { q += ln; rcb += ln; }          /* S */
                else if (op == 5) { rcb += ln; }                   /* H counted as read-consuming (Q7) */
            }
        }
    counts[3] = n_sess_reads;
    *tot_reads += n_sess_reads;

    /* ---------------- phase 2: germline set and counters ---------------- */
    uint32_t keepbit_col = 0xffffffffu, keepbit = 0;
    if (S->keep_type[s] == GA_VT_SNV && S->keep_end[s] == S->keep_pos[s] && S->keep_len[s] == 1 &&
        S->keep_allele_off[s + 1] - S->keep_allele_off[s] == 1) {
        uint8_t ch = S->keep_alleles[S->keep_allele_off[s]];
        for (int c = 0; c < 16; ++c)
            if ((uint8_t)g_code2asc[c] == ch) { keepbit = 1u << c; keepbit_col = (uint32_t)(S->keep_pos[s] - col_lo); }
    }
    for (int i = 0; i < n_cols; ++i) {
        uint32_t w = W->snv[i];
        uint32_t g = (w & (w >> 16)) & 0xffffu;
        if ((uint32_t)i == keepbit_col) g &= ~keepbit;
        W->snv[i] = g;                                           /* now: germline alleles at this column */
        counts[0] += (uint32_t)__builtin_popcount(g);
    }
    for (int k = 0; k < W->n_keys; ++k) {
        ikey_t* kk = &W->keys[k];
        /* seen in both datasets, and masked when the reference reaches the NORMAL pileup column of the key's position
         * (AM.py:474-481): only an insertion that ends its read's alignment can sit at a column no normal read covers */
        if (kk->mask == 3 && W->ncov[kk->pos - col_lo] && !key_equals_keep(R, S, s, kk, W->obs)) {
            kk->germline = 1;
            counts[kk->type == GA_VT_DEL ? 1 : 2]++;
        }
    }

    /* ---------------- phase 3: mask every read of the session ---------------- */
    int oi = 0;   /* observations were appended in the same read/op order we now replay */
    for (int d = 0; d < 2; ++d)
        for (int64_t r = rng[d][0]; r < rng[d][1]; ++r) {
            if (ref_end_of(R, r) <= first || is_unmapped(R, r)) continue;
            int n_ops_indel = 0;
            for (uint32_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; ++c) { int op = R->cigar[c] & 15; n_ops_indel += (op == 1 || op == 2); }
            const int my_obs0 = oi; oi += n_ops_indel;
            const int L = (int)(R->len_flag[r] & 0xffff);
            const int reverse = (int)((R->len_flag[r] >> 16) & 0x10);
            const uint8_t* rec = R->seq4 + 16ull * R->seq_off16[r];
            int extra = 0, any_indel = 0, any_snv = 0;
            for (int j = 0; j < n_ops_indel; ++j) {
                const obs_t* o = &W->obs[my_obs0 + j];
                if (W->keys[o->key].germline) { any_indel = 1; if (o->type == GA_VT_DEL) extra += o->len; }
            }
            const int rounds = (any_indel && g_n_reapply && reapply_twice(s, r)) ? 2 : 1;
            int need = L + rounds * extra + 1;
            if (need > W->cap_buf) { W->cap_buf = need * 2; W->sbuf = (uint8_t*)realloc(W->sbuf, (size_t)W->cap_buf); W->qbuf = (uint8_t*)realloc(W->qbuf, (size_t)W->cap_buf); }
            uint8_t* sq = W->sbuf;
            for (int k = 0; k < L; ++k) sq[k] = (uint8_t)nib_at(rec, k);
            /* SNVs (immediate substitution, AM.py:554 -> 170-176) */
            {
                int rc = R->pos[r], q = 0;
                for (uint32_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; ++c) {
                    int op = (int)(R->cigar[c] & 15), ln = (int)(R->cigar[c] >> 4);
                    if (op == 0 || op == 7 || op == 8) {
                        for (int k = 0; k < ln; ++k) {
                            int b = sq[q + k];
                            if (b != 15 && (W->snv[rc + k - col_lo] >> b) & 1u) { sq[q + k] = refc[rc + k]; any_snv = 1; }
                        }
                        q += ln; rc += ln;
                    } else if (op == 1 || op == 4) q += ln;
                    else if (op == 2 || op == 3) rc += ln;
                }
            }
            if (!any_snv && !any_indel) continue;
            int cur = L;
            uint8_t* ql = NULL;
            if (any_indel) {
                const uint8_t* qrec = qual_record(R, r);
                if (!qrec) { *err_detail = (uint32_t)r; return GA_ERR_BAD_ARGUMENT; }
                ql = W->qbuf;
                for (int k = 0; k < L; ++k) ql[k] = reverse ? qrec[L - 1 - k] : qrec[k];   /* get_forward_qualities(), AM.py:95 */
                /* left-overs sorted by VariantType value: DEL (2) before INS (3), stable (AM.py:264) */
                for (int round = 0; round < rounds; ++round)
                for (int pass = 0; pass < 2; ++pass)
                    for (int j = 0; j < n_ops_indel; ++j) {
                        const obs_t* o = &W->obs[my_obs0 + j];
                        if (!W->keys[o->key].germline) continue;
                        if (pass == 0 && o->type == GA_VT_DEL) {       /* AM.py:188-195 */
                            if ((int64_t)o->pos + o->len > ref_len) { *err_detail = (uint32_t)r; return GA_ERR_LENGTH_MISMATCH; }
                            int p = o->irp < cur ? o->irp : cur;
                            uint32_t sum = 0; for (int k = 0; k < cur; ++k) sum += ql[k];
                            uint8_t m = (uint8_t)(cur ? sum / (uint32_t)cur : 0);
                            memmove(sq + p + o->len, sq + p, (size_t)(cur - p));
                            memmove(ql + p + o->len, ql + p, (size_t)(cur - p));
                            for (int k = 0; k < o->len; ++k) { sq[p + k] = refc[o->pos + k]; ql[p + k] = m; }
                            cur += o->len;
                        } else if (pass == 1 && o->type == GA_VT_INS) { /* AM.py:183-187 */
                            int p = o->irp < cur ? o->irp : cur;
                            int e = o->irp + o->len < cur ? o->irp + o->len : cur;
                            if (e > p) {
                                memmove(sq + p, sq + e, (size_t)(cur - e));
                                memmove(ql + p, ql + e, (size_t)(cur - e));
                                cur -= (e - p);
                            }
                        }
                    }
            }
            GROW(res->v, res->cap, res->n + 1, rres_t);
            rres_t* o = &res->v[res->n++];
            o->read = (int)r; o->newlen = cur;
            memset(o->aux, 0xff, sizeof o->aux);
            if (any_indel) {                                         /* application order: DELs, then INSs */
                int ne = 0, nd = 0;
                for (int pass = 0; pass < 2; ++pass)
                    for (int j = 0; j < n_ops_indel; ++j) {
                        const obs_t* e = &W->obs[my_obs0 + j];
                        if (!W->keys[e->key].germline || (e->type == GA_VT_DEL) != (pass == 0)) continue;
                        if (ne < 2) { o->aux[3 * ne] = (uint32_t)e->irp; o->aux[3 * ne + 1] = (uint32_t)e->pos; o->aux[3 * ne + 2] = (uint32_t)e->len | (pass ? 0x80000000u : 0u); }
                        ++ne; nd += (pass == 0);
                    }
                if (ne <= 2) { o->aux[6] = (uint32_t)ne | ((uint32_t)nd << 8); o->aux[7] = 0; for (int k = 3 * ne; k < 6; ++k) o->aux[k] = 0; }
                else memset(o->aux, 0xff, sizeof o->aux);
            }
            o->seq = (uint8_t*)malloc((size_t)cur + 1);
            memcpy(o->seq, sq, (size_t)cur);
            o->qual = NULL;
            if (any_indel) {
                o->qual = (uint8_t*)malloc((size_t)cur + 1);
                for (int k = 0; k < cur; ++k) o->qual[k] = reverse ? ql[cur - 1 - k] : ql[k];      /* AM.py:213 */
            }
        }
    return GA_OK;
}

static void free_scratch(scratch_t* W) { free(W->obs); free(W->keys); free(W->snv); free(W->khead); free(W->ncov); free(W->sbuf); free(W->qbuf); }

int ga_oracle_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct {
    const ga_reads* R; const ga_sessions* S; const uint8_t* refc; int64_t ref_len; int maxspan;
    sres_t* res; uint32_t* counts;
    int next;                       /* next session block to claim (atomic) */
    pthread_mutex_t mu;
    uint64_t tr, tb; int status; uint32_t detail;
} work_t;

/* Sessions are independent units (fresh dicts per anonymize() call, AM.py:433-439), so threads
 * simply claim blocks of sessions; only the owner session writes a read's result slot. */
static void* worker(void* arg) {
    work_t* wk = (work_t*)arg;
    scratch_t W; memset(&W, 0, sizeof W);
    uint64_t tr = 0, tb = 0;
    const int ns = wk->S->n_sessions;
    for (;;) {
        int s0 = __atomic_fetch_add(&wk->next, 8, __ATOMIC_RELAXED);
        if (s0 >= ns) break;
        int s1 = s0 + 8 < ns ? s0 + 8 : ns;
        for (int s = s0; s < s1; ++s) {
            uint32_t d = 0;
            int st = process_session(wk->R, wk->S, s, wk->refc, wk->ref_len, wk->maxspan, &W, &wk->res[s],
                                     wk->counts + 4 * (size_t)s, &tr, &tb, &d);
            if (st != GA_OK) {
                pthread_mutex_lock(&wk->mu);
                if (wk->status == GA_OK) { wk->status = st; wk->detail = d; }
                pthread_mutex_unlock(&wk->mu);
            }
        }
    }
    free_scratch(&W);
    pthread_mutex_lock(&wk->mu);
    wk->tr += tr; wk->tb += tb;
    pthread_mutex_unlock(&wk->mu);
    return NULL;
}

/* Host-pointer twin of ga_run().  ref_ascii: the contig as ASCII (any case).  n_threads<=0: all cores. */
int ga_oracle_run(const ga_reads* R, const ga_sessions* S, const uint8_t* ref_ascii, int64_t ref_len,
                  ga_result* out, int n_threads) {
    init_tables();
    if (!R || !S || !out || !out->totals) return GA_ERR_BAD_ARGUMENT;
    uint8_t* refc = (uint8_t*)malloc((size_t)ref_len + 1);
    for (int64_t i = 0; i < ref_len; ++i) refc[i] = g_asc2code[ref_ascii[i]];
    int maxspan = R->max_ref_span;
    if (maxspan <= 0) { maxspan = 1; for (int64_t r = 0; r < R->n_reads; ++r) { int sp = ref_end_of(R, r) - R->pos[r]; if (sp > maxspan) maxspan = sp; } }
    const int ns = S->n_sessions;
    sres_t* res = (sres_t*)calloc((size_t)ns + 1, sizeof(sres_t));
    work_t wk;
    memset(&wk, 0, sizeof wk);
    wk.R = R; wk.S = S; wk.refc = refc; wk.ref_len = ref_len; wk.maxspan = maxspan; wk.res = res;
    wk.counts = out->sess_counts; wk.status = GA_OK;
    pthread_mutex_init(&wk.mu, NULL);
    int nt = n_threads > 0 ? n_threads : ga_oracle_threads();
    if (nt > ns) nt = ns > 0 ? ns : 1;
    if (nt <= 1) worker(&wk);
    else {
        pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
        for (int i = 0; i < nt; ++i) pthread_create(&th[i], NULL, worker, &wk);
        for (int i = 0; i < nt; ++i) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&wk.mu);
    uint64_t tr = wk.tr, tb = wk.tb;
    int status = wk.status; uint32_t detail = wk.detail;
    ga_totals* T = out->totals;
    memset(T, 0, sizeof *T);
    T->session_reads = tr; T->session_bases = tb;
    for (int s = 0; s < ns; ++s) for (int k = 0; k < 3; ++k) T->masked[k] += out->sess_counts[4 * (size_t)s + k];
    T->error = (uint32_t)status; T->error_detail = detail;
    /* emit compacted records in (session, read) order */
    uint64_t nrec = 0, s16 = 0, q16 = 0, nq = 0;
    int overflow = 0;
    for (int s = 0; s < ns; ++s)
    for (int i = 0; i < res[s].n; ++i) {
        rres_t* o = &res[s].v[i];
        const int64_t r = o->read;
        uint64_t units = ((uint64_t)o->newlen + 31) / 32;       /* 16-byte units of seq4 = 32 bases */
        if (units == 0) units = 1;
        if ((int64_t)nrec < out->cap_records && (int64_t)(s16 + units) <= out->cap_seq16 &&
            (!o->qual || (int64_t)(q16 + units) <= out->cap_qual16)) {
            out->mod_session[nrec] = (int32_t)s;
            out->mod_read[nrec] = (int32_t)r;
            out->mod_len[nrec] = (uint32_t)o->newlen;
            out->mod_seq_off16[nrec] = (uint32_t)s16;
            uint8_t* dst = out->out_seq4 + 16ull * s16;
            memset(dst, 0, 16ull * units);
            for (int k = 0; k < o->newlen; ++k) dst[k >> 1] |= (uint8_t)(o->seq[k] << (4 * (k & 1)));
            if (o->qual) {
                out->mod_qual_off16[nrec] = (uint32_t)q16;
                uint8_t* qd = out->out_qual + 32ull * q16;
                memset(qd, 0, 32ull * units);
                memcpy(qd, o->qual, (size_t)o->newlen);
            } else out->mod_qual_off16[nrec] = 0xffffffffu;
            if (g_edit_sink && (int64_t)nrec < g_edit_cap) memcpy(g_edit_sink + 8 * nrec, o->aux, sizeof o->aux);
        } else overflow = 1;
        ++nrec; s16 += units; if (o->qual) { q16 += units; ++nq; }
        free(o->seq); free(o->qual);
    }
    for (int s = 0; s < ns; ++s) free(res[s].v);
    T->n_modified = nrec; T->seq16_used = s16; T->qual16_used = q16; T->indel_records = nq;
    free(res); free(refc);
    if (status != GA_OK) return status;
    if (overflow) { T->error = GA_ERR_CAPACITY; return GA_ERR_CAPACITY; }
    return GA_OK;
}

/* Host twin of ga_result_digest (include/ga_digest.h) over a HOST result - written separately from the device kernel:
 * per record the key (global session, read gid) and the two 64-bit hashes; digest[0..3] += {sum lo, sum hi, records,
 * sum of new lengths}.  rec_keys / rec_hash may be NULL. */
int ga_oracle_digest(const ga_result* out, int64_t n_records, const ga_digest_ids* ids, uint64_t* rec_keys, uint64_t* rec_hash,
                     uint64_t* digest) {
    if (!out || !ids || !digest || n_records < 0 || n_records > out->cap_records) return GA_ERR_BAD_ARGUMENT;
    for (int64_t k = 0; k < n_records; ++k) {
        const uint64_t s = (uint64_t)((int64_t)out->mod_session[k] + ids->session_base);
        const int64_t r = out->mod_read[k];
        const uint64_t gid = r < ids->n_tumor ? (uint64_t)(ids->tumor_base + r) : ((1ull << 40) | (uint64_t)(ids->normal_base + r - ids->n_tumor));
        const uint32_t len = out->mod_len[k];
        const int has_q = out->mod_qual_off16[k] != 0xffffffffu;
        uint64_t h[2] = {GA_DIGEST_SEED_LO, GA_DIGEST_SEED_HI};
        const uint64_t mul[2] = {GA_DIGEST_MUL_LO, GA_DIGEST_MUL_HI};
        const uint8_t* sq = out->out_seq4 + 16ull * out->mod_seq_off16[k];
        const uint8_t* qq = has_q ? out->out_qual + 32ull * out->mod_qual_off16[k] : NULL;
        for (int j = 0; j < 2; ++j) {
            uint64_t v = h[j];
            v = ga_digest_mix(v, (uint64_t)ids->contig, mul[j]);
            v = ga_digest_mix(v, s, mul[j]);
            v = ga_digest_mix(v, gid, mul[j]);
            v = ga_digest_mix(v, (uint64_t)len | ((uint64_t)has_q << 32), mul[j]);
            for (uint32_t b = 0; b < len; b += 8) {                 /* eight bases per word, absent nibbles are zero */
                uint32_t w = 0;
                for (uint32_t i = b; i < b + 8 && i < len; ++i) w |= (uint32_t)nib_at(sq, (int)i) << (4 * (i - b));
                v = ga_digest_mix(v, w, mul[j]);
            }
            if (has_q)
                for (uint32_t b = 0; b < len; b += 4) {             /* four qualities per word, absent bytes are zero */
                    uint32_t w = 0;
                    for (uint32_t i = b; i < b + 4 && i < len; ++i) w |= (uint32_t)qq[i] << (8 * (i - b));
                    v = ga_digest_mix(v, w, mul[j]);
                }
            h[j] = ga_digest_fin(v);
        }
        if (rec_keys) { rec_keys[2 * k] = s; rec_keys[2 * k + 1] = gid; }
        if (rec_hash) { rec_hash[2 * k] = h[0]; rec_hash[2 * k + 1] = h[1]; }
        digest[0] += h[0]; digest[1] += h[1]; digest[2] += 1; digest[3] += len;
    }
    return GA_OK;
}
