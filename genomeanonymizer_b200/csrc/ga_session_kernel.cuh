// ga_session_kernel.cuh - one CTA runs one session (= one anonymize() call of the reference,
// anonymizer_methods.py:431-535) from discovery to compacted output.
//
// Phases (block-synchronous):
//   A  discover   every candidate read: CIGAR walk, 8-bases-per-word compare with the 4-bit reference,
//                 SNV alleles OR-ed into the per-column table, indel observations chained per column
//   R  resolve    germline = seen in tumor AND normal, minus variant_to_keep; per-session counters
//   L  list       ordered list of the reads that showed any candidate allele
//   B1 analyse    which of them are modified and how long they become
//   S  scan       output slots (one atomicAdd per session on the global cursors)
//   B2 emit       compacted modified records (sequence, and qualities when indels were masked)
#pragma once
#include "ga_device.cuh"

namespace ga {

// Per-session working set.  Small sessions keep it in shared memory, oversized ones in the CTA's
// slice of global scratch; the code below only sees pointers.
struct Tables {
    uint32_t* snv;      // [n_cols]   bit c: tumor saw base code c, bit 16+c: normal; after resolve: germline codes
    int32_t* ihead;     // [n_cols]   head of the indel-observation chain at this column
    uint32_t* cand;     // [ceil(n_range/32)] reads that showed any candidate allele
    int32_t* o_col;     // observation arrays [obs_cap]
    uint32_t* o_meta;
    uint32_t* o_read;   // session-relative read id
    int32_t* o_irp;     // in_read_pos with the reference's H/N quirk
    int32_t* o_next;
    uint32_t* clist;    // [reads_cap] ordered candidate list (session-relative ids)
    uint32_t* msize;    // [reads_cap] per candidate: kModFlag | kQualFlag | new length
    int32_t obs_cap, reads_cap, cols_cap;
};

struct Edit {           // one germline indel of one read, in application order
    int32_t irp, len, pos;
    int32_t p_eff, e_eff;   // clamped offsets in the array state the edit is applied to
    uint32_t is_ins;
    uint32_t mean;          // quality assigned to re-inserted bases (DEL only)
};

struct SessCtx {
    BatchView B;
    SessionDesc d;
    Tables T;
    int32_t s;              // session index
    int32_t nt;             // tumor candidates; session-relative id i < nt => tumor read t_begin+i
    int32_t n_range;
    int32_t first;
    int32_t keep_type, keep_pos, keep_end, keep_len;
    const uint8_t* keep_allele;
    int32_t keep_alen;
    ga_totals* totals;
    uint4* edit_keep;       // session_kernel only: null, or where the edit descriptions of indel-masked records are kept (ga_record_edits)
};

__device__ __forceinline__ int64_t read_of(const SessCtx& c, int i) {
    return i < c.nt ? (int64_t)c.d.t_begin + i : (int64_t)c.d.n_begin + (i - c.nt);
}
__device__ __forceinline__ const uint32_t* rec_of(const SessCtx& c, int64_t r) {
    return reinterpret_cast<const uint32_t*>(c.B.seq4 + 16ull * __ldg(c.B.seq_off16 + r));
}

// Python-slice-clamped allele of an observation: read bases [irp, irp+alen).
__device__ __forceinline__ int allele_len(uint32_t meta, int irp, int L) {
    const int want = (meta & kMetaIns) ? (int)(meta & kMetaLenMask) : 2;   // variation_classifier.py:87-88
    int avail = L - irp;
    if (avail < 0) avail = 0;
    return want < avail ? want : avail;
}

__device__ bool obs_equal(const SessCtx& c, int a, int b) {
    const uint32_t ma = c.T.o_meta[a], mb = c.T.o_meta[b];
    if (((ma ^ mb) & (kMetaIns | kMetaLenMask)) != 0u) return false;       // type and length
    const int64_t ra = read_of(c, (int)c.T.o_read[a]), rb = read_of(c, (int)c.T.o_read[b]);
    const int La = (int)(__ldg(c.B.len_flag + ra) & 0xffffu), Lb = (int)(__ldg(c.B.len_flag + rb) & 0xffffu);
    const int ia = c.T.o_irp[a], ib = c.T.o_irp[b];
    const int na = allele_len(ma, ia, La), nb = allele_len(mb, ib, Lb);
    if (na != nb) return false;
    const uint32_t* pa = rec_of(c, ra);
    const uint32_t* pb = rec_of(c, rb);
    for (int j = 0; j < na; ++j)
        if (read_code(pa, ia + j) != read_code(pb, ib + j)) return false;
    return true;
}

__device__ bool obs_equals_keep(const SessCtx& c, int a) {
    const uint32_t m = c.T.o_meta[a];
    const int type = (m & kMetaIns) ? GA_VT_INS : GA_VT_DEL;
    const int len = (int)(m & kMetaLenMask);
    const int pos = c.T.o_col[a] + c.d.col_begin;
    if (c.keep_type != type || c.keep_pos != pos || c.keep_len != len) return false;
    const int end = (type == GA_VT_INS) ? pos + 1 : pos + len - 1;         // variation_classifier.py:86
    if (c.keep_end != end) return false;
    const int64_t r = read_of(c, (int)c.T.o_read[a]);
    const int L = (int)(__ldg(c.B.len_flag + r) & 0xffffu);
    const int irp = c.T.o_irp[a];
    const int na = allele_len(m, irp, L);
    if (na != c.keep_alen) return false;
    const uint32_t* p = rec_of(c, r);
    const char* code2asc = "=ACMGRSVTWYHKDBN";
    for (int j = 0; j < na; ++j)
        if (c.keep_allele[j] != (uint8_t)code2asc[read_code(p, irp + j)]) return false;
    return true;
}

// The reference masks the variants of a position when it reaches the NORMAL pileup column of that position
// (anonymizer_methods.py:474-481): a key seen in both datasets is masked only if some normal read of the session covers
// its position.  A read that shows an SNV or a DEL covers it, and so does a read whose insertion is followed by aligned
// bases; an insertion that ends its read's alignment sits at the read's reference_end, which the read does not cover -
// then another normal read has to (kMetaTrail / kMetaCol).  This is that test: rare, one thread, the reads in position order.
__device__ __noinline__ bool normal_covers(const SessCtx& c, int first, int p) {
    for (int64_t r = c.d.n_begin; r < c.d.n_end; ++r) {
        const int pos = __ldg(c.B.pos + r);
        if (pos > p) break;
        if ((__ldg(c.B.len_flag + r) >> 16) & 0x4u) continue;             // placed-unmapped: in no pileup
        const int span = ref_span_of(c.B.cigar, __ldg(c.B.cigar_off + r), __ldg(c.B.cigar_off + r + 1));
        if (pos + span <= first) continue;                                // fetched by range, does not reach the region
        if (p < pos + span) return true;
    }
    return false;
}

// ---------------------------------------------------------------- phase A: one read
__device__ void discover_read(const SessCtx& c, int i, uint32_t* n_obs, uint32_t* sess_reads, uint32_t* sess_bases) {
    const int64_t r = read_of(c, i);
    const uint32_t c0 = __ldg(c.B.cigar_off + r), c1 = __ldg(c.B.cigar_off + r + 1);
    const int pos = __ldg(c.B.pos + r);
    const int span = ref_span_of(c.B.cigar, c0, c1);
    if (pos + span <= c.first) return;                                   // fetched by range, does not reach the region
    const uint32_t lf = __ldg(c.B.len_flag + r);
    if ((lf >> 16) & 0x4u) return;                                        // a placed-unmapped mate: htslib's pileup drops BAM_FUNMAP records
    const int L = (int)(lf & 0xffffu);
    atomicAdd(sess_reads, 1u);
    atomicAdd(sess_bases, (uint32_t)L);
    if ((int64_t)pos + span > c.B.ref_len || pos < 0 || pos < c.d.col_begin || pos + span - c.d.col_begin >= c.d.n_cols) {
        raise_error(c.totals, GA_ERR_OFFSET_RANGE, (uint32_t)r);
        return;
    }
    const uint32_t ds = i < c.nt ? 0u : 1u;
    const uint32_t* rec = rec_of(c, r);
    bool flagged = false;
    int rc = pos, q = 0, ccl = 0, rcb = 0;
    for (uint32_t ci = c0; ci < c1; ++ci) {
        const uint32_t w = __ldg(c.B.cigar + ci), op = w & 15u;
        const int ln = (int)(w >> 4);
        if (op == 0u || op == 7u || op == 8u) {
            if (q + ln > L) { raise_error(c.totals, GA_ERR_OFFSET_RANGE, (uint32_t)r); return; }
            scan_segment(rec, c.B.ref4, q, q + ln, rc, [&](int, int rp, uint32_t b, uint32_t) {
                atomicOr(&c.T.snv[rp - c.d.col_begin], 1u << (b + 16u * ds));
                flagged = true;
            });
            q += ln; rc += ln; ccl += ln;
        } else if (op == 1u || op == 2u) {
            const uint32_t slot = atomicAdd(n_obs, 1u);
            if ((int)slot >= c.T.obs_cap) { raise_error(c.totals, GA_ERR_CAPACITY, (uint32_t)c.s); return; }
            const int col = rc - c.d.col_begin;
            c.T.o_col[slot] = col;
            c.T.o_meta[slot] = (op == 1u ? kMetaIns | (rc - pos == span ? kMetaTrail : 0u) : 0u) | (ds ? kMetaDs : 0u) | ((uint32_t)ln & kMetaLenMask);
            c.T.o_read[slot] = (uint32_t)i;
            c.T.o_irp[slot] = ccl + rcb;                                  // variation_classifier.py:82
            __threadfence_block();
            c.T.o_next[slot] = atomicExch(&c.T.ihead[col], (int)slot);
            flagged = true;
            if (op == 1u) { q += ln; rcb += ln; } else { rc += ln; ccl += ln; rcb -= ln; }
        } else if (op == 3u) { rc += ln; ccl += ln; }
        else if (op == 4u) { q += ln; rcb += ln; }
        else if (op == 5u) { rcb += ln; }
    }
    if (flagged) atomicOr(&c.T.cand[i >> 5], 1u << (i & 31));
}

// ---------------------------------------------------------------- phases B1/B2: one read
// Finds this read's germline indel observation for the op at (col, type); -1 if none / not germline.
__device__ __forceinline__ int find_my_obs(const SessCtx& c, int col, uint32_t ins_bit, int i) {
    for (int o = c.T.ihead[col]; o >= 0; o = c.T.o_next[o]) {
        const uint32_t m = c.T.o_meta[o];
        if ((int)c.T.o_read[o] == i && ((m & kMetaIns) == ins_bit)) return (m & kMetaGerm) ? o : -1;
    }
    return -1;
}

// Walks the read once: detects germline SNV hits and collects the germline indel edits in application
// order (all DELs, then all INSs: stable sort by VariantType value, anonymizer_methods.py:264).
// Returns the new length; *any_snv, *n_edits report what was found.
template <bool SCAN_SNV = true>
__device__ int analyse_read(const SessCtx& c, int i, int64_t r, Edit* edits, int* n_edits, bool* any_snv, bool* too_many) {
    const uint32_t c0 = __ldg(c.B.cigar_off + r), c1 = __ldg(c.B.cigar_off + r + 1);
    const int pos = __ldg(c.B.pos + r);
    const int L = (int)(__ldg(c.B.len_flag + r) & 0xffffu);
    const uint32_t* rec = rec_of(c, r);
    bool hit = false;
    int ne = 0, n_del = 0;
    // pass 1: SNVs and DELs
    {
        int rc = pos, q = 0, ccl = 0, rcb = 0;
        for (uint32_t ci = c0; ci < c1; ++ci) {
            const uint32_t w = __ldg(c.B.cigar + ci), op = w & 15u;
            const int ln = (int)(w >> 4);
            if (op == 0u || op == 7u || op == 8u) {
                if (SCAN_SNV)
                    scan_segment(rec, c.B.ref4, q, q + ln, rc, [&](int, int rp, uint32_t b, uint32_t) {
                        if ((c.T.snv[rp - c.d.col_begin] >> b) & 1u) hit = true;
                    });
                q += ln; rc += ln; ccl += ln;
            } else if (op == 2u) {
                if (find_my_obs(c, rc - c.d.col_begin, 0u, i) >= 0) {
                    if (ne < GA_MAX_EDITS) { edits[ne].irp = ccl + rcb; edits[ne].len = ln; edits[ne].pos = rc; edits[ne].is_ins = 0u; ++ne; }
                    else *too_many = true;
                }
                rc += ln; ccl += ln; rcb -= ln;
            } else if (op == 1u) { q += ln; rcb += ln; }
            else if (op == 3u) { rc += ln; ccl += ln; }
            else if (op == 4u) { q += ln; rcb += ln; }
            else if (op == 5u) { rcb += ln; }
        }
    }
    n_del = ne;
    // pass 2: INSs
    {
        int rc = pos, ccl = 0, rcb = 0;
        for (uint32_t ci = c0; ci < c1; ++ci) {
            const uint32_t w = __ldg(c.B.cigar + ci), op = w & 15u;
            const int ln = (int)(w >> 4);
            if (op == 0u || op == 7u || op == 8u || op == 3u) { rc += ln; ccl += ln; }
            else if (op == 2u) { rc += ln; ccl += ln; rcb -= ln; }
            else if (op == 1u) {
                if (find_my_obs(c, rc - c.d.col_begin, kMetaIns, i) >= 0) {
                    if (ne < GA_MAX_EDITS) { edits[ne].irp = ccl + rcb; edits[ne].len = ln; edits[ne].pos = rc; edits[ne].is_ins = 1u; ++ne; }
                    else *too_many = true;
                }
                rcb += ln;
            } else if (op == 4u || op == 5u) { rcb += ln; }
        }
    }
    // clamped offsets exactly as Python slicing applies them (anonymizer_methods.py:186-195)
    int cur = L;
    for (int k = 0; k < n_del; ++k) {
        edits[k].p_eff = edits[k].irp < cur ? edits[k].irp : cur;
        edits[k].e_eff = edits[k].p_eff + edits[k].len;
        cur += edits[k].len;
    }
    for (int k = n_del; k < ne; ++k) {
        const int p = edits[k].irp < cur ? edits[k].irp : cur;
        const int e = edits[k].irp + edits[k].len < cur ? edits[k].irp + edits[k].len : cur;
        edits[k].p_eff = p; edits[k].e_eff = e > p ? e : p;
        cur -= (edits[k].e_eff - p);
    }
    *n_edits = ne; *any_snv = hit;
    return cur;
}

__device__ __forceinline__ const uint8_t* qual_record(const BatchView& B, int64_t r) {
    if (!B.qual) return nullptr;
    if (!B.qual_reads) return B.qual + 32ull * __ldg(B.seq_off16 + r);
    int64_t b = 0, e = B.n_qual;
    while (b < e) { const int64_t m = (b + e) >> 1; if (__ldg(B.qual_reads + m) < r) b = m + 1; else e = m; }
    if (b < B.n_qual && __ldg(B.qual_reads + b) == r) return B.qual + 32ull * __ldg(B.qual_off16 + b);
    return nullptr;
}

// Same lookup restricted to the session's slice [lo, hi) of the sparse quality index.
__device__ __forceinline__ const uint8_t* qual_record_in(const BatchView& B, int64_t r, int64_t lo, int64_t hi) {
    if (!B.qual) return nullptr;
    if (!B.qual_reads) return B.qual + 32ull * __ldg(B.seq_off16 + r);
    int64_t b = lo, e = hi;
    while (b < e) { const int64_t m = (b + e) >> 1; if (__ldg(B.qual_reads + m) < r) b = m + 1; else e = m; }
    if (b < hi && __ldg(B.qual_reads + b) == r) return B.qual + 32ull * __ldg(B.qual_off16 + b);
    return nullptr;
}

// Maps a final array index back through the edits (last applied first).  Returns the original index,
// or -1 - (edit index) when the element was inserted by that DEL edit (then *k_in is its offset).
__device__ __forceinline__ int map_back(const Edit* edits, int n_del, int ne, int j, int* k_in) {
#pragma unroll 1
    for (int k = ne - 1; k >= n_del; --k)
        if (j >= edits[k].p_eff) j += edits[k].e_eff - edits[k].p_eff;
#pragma unroll 1
    for (int k = n_del - 1; k >= 0; --k) {
        if (j >= edits[k].e_eff) j -= edits[k].len;
        else if (j >= edits[k].p_eff) { *k_in = j - edits[k].p_eff; return -1 - k; }
    }
    return j;
}

// Base code at original query index j after SNV masking (anonymizer_methods.py:170-176).
__device__ uint32_t masked_base(const SessCtx& c, int64_t r, const uint32_t* rec, int j) {
    const uint32_t b = read_code(rec, j);
    if (b == 15u) return b;
    const uint32_t c0 = __ldg(c.B.cigar_off + r), c1 = __ldg(c.B.cigar_off + r + 1);
    int rc = __ldg(c.B.pos + r), q = 0;
    for (uint32_t ci = c0; ci < c1; ++ci) {
        const uint32_t w = __ldg(c.B.cigar + ci), op = w & 15u;
        const int ln = (int)(w >> 4);
        if (op == 0u || op == 7u || op == 8u) {
            if (j < q + ln) {
                const int rp = rc + (j - q);
                if ((c.T.snv[rp - c.d.col_begin] >> b) & 1u) return ref_code(c.B.ref4, rp);
                return b;
            }
            q += ln; rc += ln;
        } else if (op == 1u || op == 4u) { if (j < q + ln) return b; q += ln; }
        else if (op == 2u || op == 3u) rc += ln;
    }
    return b;
}

__device__ void emit_read(const SessCtx& c, const ResultView& O, int i, int64_t r, uint64_t rec_idx, uint64_t seq16, uint64_t qual16,
                          int new_len, bool has_qual) {
    Edit edits[GA_MAX_EDITS];
    int ne = 0; bool any_snv = false, too_many = false;
    const int L = (int)(__ldg(c.B.len_flag + r) & 0xffffu);
    const uint32_t* rec = rec_of(c, r);
    O.mod_session[rec_idx] = c.s;
    O.mod_read[rec_idx] = (int32_t)r;
    O.mod_len[rec_idx] = (uint32_t)new_len;
    O.mod_seq_off16[rec_idx] = (uint32_t)seq16;
    O.mod_qual_off16[rec_idx] = has_qual ? (uint32_t)qual16 : 0xffffffffu;
    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
    int units = (new_len + 31) / 32; if (units < 1) units = 1;
    if (!has_qual) {
        // SNV-only: copy the record, then patch the masked nibbles in place
        const int nw = units * 4;
        for (int w = 0; w < nw; ++w) {
            uint32_t v = __ldg(rec + w);
            const int qb = w * 8;
            if (qb + 8 > L) v &= (qb >= L) ? 0u : (0xffffffffu >> ((qb + 8 - L) * 4));
            oseq[w] = v;
        }
        const uint32_t c0 = __ldg(c.B.cigar_off + r), c1 = __ldg(c.B.cigar_off + r + 1);
        int rc = __ldg(c.B.pos + r), q = 0;
        for (uint32_t ci = c0; ci < c1; ++ci) {
            const uint32_t w = __ldg(c.B.cigar + ci), op = w & 15u;
            const int ln = (int)(w >> 4);
            if (op == 0u || op == 7u || op == 8u) {
                scan_segment(rec, c.B.ref4, q, q + ln, rc, [&](int qq, int rp, uint32_t b, uint32_t rf) {
                    if ((c.T.snv[rp - c.d.col_begin] >> b) & 1u) {
                        uint32_t v = oseq[qq >> 3];
                        v = (v & ~(0xfu << ((qq & 7) * 4))) | (rf << ((qq & 7) * 4));
                        oseq[qq >> 3] = v;
                    }
                });
                q += ln; rc += ln;
            } else if (op == 1u || op == 4u) q += ln;
            else if (op == 2u || op == 3u) rc += ln;
        }
        return;
    }
    // indel-masked read: rebuild every element through the backward index map
    analyse_read(c, i, r, edits, &ne, &any_snv, &too_many);
    int n_del = 0;
    while (n_del < ne && !edits[n_del].is_ins) ++n_del;
    if (c.edit_keep && ne <= 2 && !too_many) {                           // same eight words as the resolve kernels keep (ga_record_edits)
        const uint32_t l0 = ne > 0 ? (uint32_t)edits[0].len | (edits[0].is_ins ? 0x80000000u : 0u) : 0u;
        const uint32_t l1 = ne > 1 ? (uint32_t)edits[1].len | (edits[1].is_ins ? 0x80000000u : 0u) : 0u;
        c.edit_keep[2 * rec_idx] = make_uint4(ne > 0 ? (uint32_t)edits[0].irp : 0u, ne > 0 ? (uint32_t)edits[0].pos : 0u, l0, ne > 1 ? (uint32_t)edits[1].irp : 0u);
        c.edit_keep[2 * rec_idx + 1] = make_uint4(ne > 1 ? (uint32_t)edits[1].pos : 0u, l1, (uint32_t)ne | ((uint32_t)n_del << 8), 0u);
    }
    const uint8_t* qrec = qual_record(c.B, r);
    if (!qrec) { raise_error(c.totals, GA_ERR_BAD_ARGUMENT, (uint32_t)r); return; }
    const bool reverse = ((__ldg(c.B.len_flag + r) >> 16) & 0x10u) != 0u;
    // quality assigned by each DEL: floor(mean(current qualities)) (anonymizer_methods.py:193)
    {
        uint32_t sum = 0;
        for (int k = 0; k < L; ++k) sum += qrec[k];
        uint32_t n = (uint32_t)L;
        for (int k = 0; k < n_del; ++k) {
            const uint32_t m = n ? sum / n : 0u;
            edits[k].mean = m;
            sum += m * (uint32_t)edits[k].len; n += (uint32_t)edits[k].len;
            if ((int64_t)edits[k].pos + edits[k].len > c.B.ref_len) raise_error(c.totals, GA_ERR_LENGTH_MISMATCH, (uint32_t)r);
        }
    }
    for (int w = 0; w < units * 4; ++w) {
        uint32_t v = 0;
        for (int k = 0; k < 8; ++k) {
            const int j = w * 8 + k;
            if (j >= new_len) break;
            int kin = 0;
            const int src = map_back(edits, n_del, ne, j, &kin);
            const uint32_t code = src >= 0 ? masked_base(c, r, rec, src) : ref_code(c.B.ref4, (int64_t)edits[-1 - src].pos + kin);
            v |= code << (k * 4);
        }
        oseq[w] = v;
    }
    uint32_t* oq = reinterpret_cast<uint32_t*>(O.out_qual + 32ull * qual16);
    for (int w = 0; w < units * 8; ++w) {
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) {
            const int jp = w * 4 + k;
            if (jp >= new_len) break;
            // printed order = reversed forward-orientation array for reverse reads (anonymizer_methods.py:95,213)
            const int jf = reverse ? new_len - 1 - jp : jp;
            int kin = 0;
            const int src = map_back(edits, n_del, ne, jf, &kin);
            uint32_t qv;
            if (src >= 0) qv = qrec[reverse ? L - 1 - src : src];
            else qv = edits[-1 - src].mean;
            v |= qv << (k * 8);
        }
        oq[w] = v;
    }
}

// ---------------------------------------------------------------- the session kernel
struct SmemLayout {
    uint32_t snv[kColsCap];
    int32_t ihead[kColsCap];
    uint32_t cand[kReadsCap / 32];
    int32_t o_col[kObsCap];
    uint32_t o_meta[kObsCap];
    uint32_t o_read[kObsCap];
    int32_t o_irp[kObsCap];
    int32_t o_next[kObsCap];
    uint32_t clist[kReadsCap];
    uint32_t msize[kReadsCap];
};

struct BigScratch {       // per-CTA slice of global scratch for oversized sessions
    uint8_t* base;
    int64_t bytes_per_cta;
    int32_t cols_cap, reads_cap, obs_cap;
};

// BIG: tables in the CTA's slice of global scratch, otherwise in shared memory (SmemLayout).  The engine launches both
// over big_list: the shared-memory variant takes the listed sessions whose tables fit it, the global-scratch one the rest.
template <bool BIG>
__global__ void __launch_bounds__(kThreads) session_kernel(BatchView B, SessView S, const SessionDesc* __restrict__ descs,
                                                           const int32_t* __restrict__ big_list, const int32_t* __restrict__ n_big,
                                                           ResultView O, BigScratch scr, unsigned int* __restrict__ ticket, uint4* edit_keep) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_scan[kThreads / 32 + 1];
    __shared__ uint32_t s_nobs, s_reads, s_bases, s_cnt[3], s_ncand;
    __shared__ int s_session;
    __shared__ unsigned long long s_base[3];

    SessCtx c;
    c.B = B;
    c.totals = O.totals;
    c.edit_keep = edit_keep;
    if (BIG) {
        uint8_t* p = scr.base + (size_t)blockIdx.x * (size_t)scr.bytes_per_cta;
        c.T.snv = reinterpret_cast<uint32_t*>(p); p += 4ull * scr.cols_cap;
        c.T.ihead = reinterpret_cast<int32_t*>(p); p += 4ull * scr.cols_cap;
        c.T.cand = reinterpret_cast<uint32_t*>(p); p += 4ull * ((scr.reads_cap + 31) / 32);
        c.T.o_col = reinterpret_cast<int32_t*>(p); p += 4ull * scr.obs_cap;
        c.T.o_meta = reinterpret_cast<uint32_t*>(p); p += 4ull * scr.obs_cap;
        c.T.o_read = reinterpret_cast<uint32_t*>(p); p += 4ull * scr.obs_cap;
        c.T.o_irp = reinterpret_cast<int32_t*>(p); p += 4ull * scr.obs_cap;
        c.T.o_next = reinterpret_cast<int32_t*>(p); p += 4ull * scr.obs_cap;
        c.T.clist = reinterpret_cast<uint32_t*>(p); p += 4ull * scr.reads_cap;
        c.T.msize = reinterpret_cast<uint32_t*>(p);
        c.T.obs_cap = scr.obs_cap; c.T.reads_cap = scr.reads_cap; c.T.cols_cap = scr.cols_cap;
    } else {
        SmemLayout* sm = reinterpret_cast<SmemLayout*>(smem_raw);
        c.T.snv = sm->snv; c.T.ihead = sm->ihead; c.T.cand = sm->cand;
        c.T.o_col = sm->o_col; c.T.o_meta = sm->o_meta; c.T.o_read = sm->o_read; c.T.o_irp = sm->o_irp; c.T.o_next = sm->o_next;
        c.T.clist = sm->clist; c.T.msize = sm->msize;
        c.T.obs_cap = kObsCap; c.T.reads_cap = kReadsCap; c.T.cols_cap = kColsCap;
    }
    const int tid = threadIdx.x;
    const int n_work = *n_big;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_session = (int)atomicAdd(ticket, 1u);
        __syncthreads();
        const int widx = s_session;
        if (widx >= n_work) break;
        const int s = big_list[widx];
        c.d = descs[s];
        c.s = s;
        c.nt = c.d.t_end - c.d.t_begin;
        c.n_range = c.nt + (c.d.n_end - c.d.n_begin);
        const bool fits_smem = c.d.n_cols <= kColsCap && c.n_range <= kReadsCap && c.d.obs_bound <= kObsCap;
        if (BIG == fits_smem) continue;                              // the other variant's session
        c.first = S.first[s];
        c.keep_type = S.keep_type[s]; c.keep_pos = S.keep_pos[s]; c.keep_end = S.keep_end[s]; c.keep_len = S.keep_len[s];
        c.keep_allele = S.keep_alleles + S.keep_allele_off[s];
        c.keep_alen = (int)(S.keep_allele_off[s + 1] - S.keep_allele_off[s]);
        if (BIG && (c.d.n_cols > c.T.cols_cap || c.n_range > c.T.reads_cap || c.d.obs_bound > c.T.obs_cap)) {
            if (tid == 0) raise_error(O.totals, GA_ERR_CAPACITY, (uint32_t)s);
            continue;
        }
        const int n_cols = c.d.n_cols;
        const int n_cw = (c.n_range + 31) >> 5;

        // ---- zero the working set
        for (int k = tid; k < n_cols; k += kThreads) { c.T.snv[k] = 0u; c.T.ihead[k] = -1; }
        for (int k = tid; k < n_cw; k += kThreads) c.T.cand[k] = 0u;
        if (tid == 0) { s_nobs = 0; s_reads = 0; s_bases = 0; s_cnt[0] = s_cnt[1] = s_cnt[2] = 0; s_ncand = 0; }
        __syncthreads();

        // ---- phase A: discover
        for (int i = tid; i < c.n_range; i += kThreads) discover_read(c, i, &s_nobs, &s_reads, &s_bases);
        __syncthreads();
        const int n_obs = min((int)s_nobs, c.T.obs_cap);

        // ---- phase R: resolve.  SNVs: germline codes = tumor & normal, minus variant_to_keep (AM.py:546-547)
        {
            uint32_t keep_bit = 0u; int keep_col = -1;
            if (c.keep_type == GA_VT_SNV && c.keep_end == c.keep_pos && c.keep_len == 1 && c.keep_alen == 1) {
                const char* code2asc = "=ACMGRSVTWYHKDBN";
                const uint8_t ch = c.keep_allele[0];
                for (int k = 0; k < 16; ++k) if ((uint8_t)code2asc[k] == ch) { keep_bit = 1u << k; keep_col = c.keep_pos - c.d.col_begin; }
            }
            uint32_t cnt = 0;
            for (int k = tid; k < n_cols; k += kThreads) {
                const uint32_t w = c.T.snv[k];
                uint32_t g = (w & (w >> 16)) & 0xffffu;
                if (k == keep_col) g &= ~keep_bit;
                c.T.snv[k] = g;
                cnt += __popc(g);
            }
            if (cnt) atomicAdd(&s_cnt[0], cnt);
        }
        // indels: exact key equality along each column chain (variants.py:83-96)
        for (int o = tid; o < n_obs; o += kThreads) {
            const uint32_t m = c.T.o_meta[o];
            bool germ = false, rep = true;
            bool col = (m & kMetaDs) && !(m & kMetaTrail);               // a normal read that shows the key covers its position
            for (int o2 = c.T.ihead[c.T.o_col[o]]; o2 >= 0; o2 = c.T.o_next[o2]) {
                if (o2 == o) continue;
                if (!obs_equal(c, o, o2)) continue;
                const uint32_t m2 = c.T.o_meta[o2];
                if ((m2 ^ m) & kMetaDs) germ = true;
                if ((m2 & kMetaDs) && !(m2 & kMetaTrail)) col = true;
                if (o2 < o) rep = false;
            }
            if (germ && !col) germ = normal_covers(c, c.first, c.T.o_col[o] + c.d.col_begin);   // the normal pileup must have a column there (anonymizer_methods.py:474-481)
            if (germ && obs_equals_keep(c, o)) germ = false;
            if (germ) {
                atomicOr(&c.T.o_meta[o], kMetaGerm | (rep ? kMetaRep : 0u));
                if (rep) atomicAdd(&s_cnt[(m & kMetaIns) ? 2 : 1], 1u);
            }
        }
        __syncthreads();

        // ---- phase L: ordered candidate list
        for (int base = 0; base < n_cw; base += kThreads) {
            const int w = base + tid;
            const uint32_t bits = w < n_cw ? c.T.cand[w] : 0u;
            uint32_t total = 0;
            uint32_t off = block_exclusive_scan(__popc(bits), s_scan, &total);
            uint32_t dst = s_ncand + off;
            uint32_t b = bits;
            while (b) { const int k = __ffs(b) - 1; b &= b - 1; c.T.clist[dst++] = (uint32_t)(w * 32 + k); }
            __syncthreads();
            if (tid == 0) s_ncand += total;
            __syncthreads();
        }
        const int n_cand = (int)s_ncand;

        // ---- phase B1: analyse candidates
        for (int k = tid; k < n_cand; k += kThreads) {
            const int i = (int)c.T.clist[k];
            const int64_t r = read_of(c, i);
            Edit edits[GA_MAX_EDITS];
            int ne = 0; bool any_snv = false, too_many = false;
            const int new_len = analyse_read(c, i, r, edits, &ne, &any_snv, &too_many);
            if (too_many) raise_error(O.totals, GA_ERR_UNSUPPORTED, (uint32_t)r);
            uint32_t m = 0u;
            if (any_snv || ne > 0) m = kModFlag | (ne > 0 ? kQualFlag : 0u) | ((uint32_t)new_len & kLenMask);
            c.T.msize[k] = m;
        }
        __syncthreads();

        // ---- phase S: output slots.  Each thread owns a contiguous slice of the candidate list.
        const int per = (n_cand + kThreads - 1) / kThreads;
        const int k0 = min(tid * per, n_cand), k1 = min(k0 + per, n_cand);
        uint32_t my_rec = 0, my_seq = 0, my_qual = 0;
        for (int k = k0; k < k1; ++k) {
            const uint32_t m = c.T.msize[k];
            if (!(m & kModFlag)) continue;
            uint32_t units = ((m & kLenMask) + 31u) / 32u; if (units < 1u) units = 1u;
            ++my_rec; my_seq += units; if (m & kQualFlag) my_qual += units;
        }
        uint32_t tot_rec, tot_seq, tot_qual;
        const uint32_t off_rec = block_exclusive_scan(my_rec, s_scan, &tot_rec);
        const uint32_t off_seq = block_exclusive_scan(my_seq, s_scan, &tot_seq);
        const uint32_t off_qual = block_exclusive_scan(my_qual, s_scan, &tot_qual);
        if (tid == 0) {
            s_base[0] = atomicAdd((unsigned long long*)&O.totals->n_modified, (unsigned long long)tot_rec);
            s_base[1] = atomicAdd((unsigned long long*)&O.totals->seq16_used, (unsigned long long)tot_seq);
            s_base[2] = atomicAdd((unsigned long long*)&O.totals->qual16_used, (unsigned long long)tot_qual);
            atomicAdd((unsigned long long*)&O.totals->session_reads, (unsigned long long)s_reads);
            atomicAdd((unsigned long long*)&O.totals->session_bases, (unsigned long long)s_bases);
            for (int k = 0; k < 3; ++k) {
                O.sess_counts[4 * (size_t)s + k] = s_cnt[k];
                if (s_cnt[k]) atomicAdd((unsigned long long*)&O.totals->masked[k], (unsigned long long)s_cnt[k]);
            }
            O.sess_counts[4 * (size_t)s + 3] = s_reads;
        }
        __syncthreads();
        const bool fits = (int64_t)(s_base[0] + tot_rec) <= O.cap_records && (int64_t)(s_base[1] + tot_seq) <= O.cap_seq16 &&
                          (int64_t)(s_base[2] + tot_qual) <= O.cap_qual16;
        if (!fits) { if (tid == 0) raise_error(O.totals, GA_ERR_CAPACITY, 0xffffffffu); continue; }

        // ---- phase B2: emit
        uint64_t rec_idx = s_base[0] + off_rec, seq16 = s_base[1] + off_seq, qual16 = s_base[2] + off_qual;
        uint32_t n_q = 0;
        for (int k = k0; k < k1; ++k) {
            const uint32_t m = c.T.msize[k];
            if (!(m & kModFlag)) continue;
            const int new_len = (int)(m & kLenMask);
            uint32_t units = ((uint32_t)new_len + 31u) / 32u; if (units < 1u) units = 1u;
            const int i = (int)c.T.clist[k];
            emit_read(c, O, i, read_of(c, i), rec_idx, seq16, qual16, new_len, (m & kQualFlag) != 0u);
            ++rec_idx; seq16 += units;
            if (m & kQualFlag) { qual16 += units; ++n_q; }
        }
        if (n_q) atomicAdd((unsigned long long*)&O.totals->indel_records, (unsigned long long)n_q);
    }
}

}  // namespace ga
