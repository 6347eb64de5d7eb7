// ga_record_ops.cuh - record-level building blocks shared by the resolve and emission kernels of the streaming
// pipeline (ga_scan_kernel.cuh -> ga_resolve_kernel.cuh -> ga_emit_kernel.cuh): capacities, key equality of indel
// observations (variants.py:83-96), the germline edits of one read in application order with Python-slice clamping
// (anonymizer_methods.py:178-203, 254-270) and the general emission of SNV- / indel-masked records by a group of
// 8 lanes.  The first-generation session kernel (ga_session_kernel.cuh) is included for its context structs and
// remains the global-scratch fallback for oversize sessions.
#pragma once
#include "ga_session_kernel.cuh"

namespace ga {

constexpr int kCols2 = 2688;           // allele-table columns per session (shared-memory paths)
constexpr int kReads2 = 4096;          // candidate reads per session
constexpr int kMod2 = 1024;            // modified reads per session
constexpr uint32_t kLen2 = (1u << 24) - 1;   // msize: length bits (flags above: kModFlag, kQualFlag)
constexpr int kGermCap = 252;          // germline SNV alleles per session handed to the emission kernels (below 256: per-read hit counters are bytes)
constexpr int kGroup = 8;              // lanes that cooperate on one non-trivial output record

struct EditAux { int32_t irp0, pos0; uint32_t len0; int32_t irp1, pos1; uint32_t len1; uint32_t ne_ndel, qidx; };   // len bit 31 = INS; ne | n_del << 8;
// qidx = the read's expected slot in ga_reads.qual_reads (its slice begin + the read's ordinal among the item's
// reads with an I/D op), 0xffffffff when unknown: the emission kernel verifies it and searches otherwise
static_assert(sizeof(EditAux) == 32, "EditAux fills one 32-byte quality unit");


__device__ __forceinline__ uint4 ldg128(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}


__device__ __forceinline__ uint32_t tail_mask(int L, int word) {           // valid nibbles of query word `word`
    const int nv = L - word * 8;
    return nv >= 8 ? 0xffffffffu : (nv <= 0 ? 0u : (0xffffffffu >> ((8 - nv) * 4)));
}

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t* total) {
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += n; }
    *total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}


// ------------------------------------------------------------------ indel observations (shared memory only)
template <class SM> __device__ __forceinline__ int obs_read(const SM* sm, int o) { return (int)(sm->o_read[o] & 0xffffu); }
template <class SM> __device__ __forceinline__ int obs_alen(const SM* sm, int o) { return (int)(sm->o_read[o] >> 16); }

// CalledGenomicVariant.__eq__ (variants.py:83-96) between two observations of the same column: type, length
// and allele bases.  Alleles up to 16 bases are decided by the stored signature, longer ones re-read the records.
template <class SM> __device__ __forceinline__ bool obs_equal2(const SessCtx& c, const SM* sm, int a, int b) {
    if (((sm->o_meta[a] ^ sm->o_meta[b]) & (kMetaIns | kMetaLenMask)) != 0u) return false;
    const int na = obs_alen(sm, a);
    if (na != obs_alen(sm, b) || sm->o_sig0[a] != sm->o_sig0[b] || sm->o_sig1[a] != sm->o_sig1[b]) return false;
    if (na <= 16) return true;
    const uint32_t* pa = rec_of(c, read_of(c, obs_read(sm, a)));
    const uint32_t* pb = rec_of(c, read_of(c, obs_read(sm, b)));
    const int ia = sm->o_irp[a], ib = sm->o_irp[b];
    for (int j = 16; j < na; ++j)
        if (read_code(pa, ia + j) != read_code(pb, ib + j)) return false;
    return true;
}

template <class SM> __device__ bool obs_equals_keep2(const SessCtx& c, const SM* sm, int a) {
    const uint32_t m = sm->o_meta[a];
    const int type = (m & kMetaIns) ? GA_VT_INS : GA_VT_DEL;
    const int len = (int)(m & kMetaLenMask);
    const int pos = sm->o_col[a] + c.d.col_begin;
    if (c.keep_type != type || c.keep_pos != pos || c.keep_len != len) return false;
    const int end = (type == GA_VT_INS) ? pos + 1 : pos + len - 1;         // variation_classifier.py:86
    if (c.keep_end != end) return false;
    const int na = obs_alen(sm, a);
    if (na != c.keep_alen) return false;
    const uint32_t* p = rec_of(c, read_of(c, obs_read(sm, a)));
    const int irp = sm->o_irp[a];
    const char* code2asc = "=ACMGRSVTWYHKDBN";
    for (int j = 0; j < na; ++j)
        if (c.keep_allele[j] != (uint8_t)code2asc[read_code(p, irp + j)]) return false;
    return true;
}

// The germline indel edits of modified read k in application order (all DELs, then all INSs, each in CIGAR
// order: stable sort by VariantType value, anonymizer_methods.py:264) with the offsets clamped exactly as
// Python slicing applies them (anonymizer_methods.py:186-195).  The read's germline observations hang on
// mhead[k]; their slots ascend in CIGAR order (one thread allocated them), so sorting by slot restores it.
// Returns the new length.
__device__ __forceinline__ void write_record_meta(const ResultView& O, uint64_t rec_idx, int s, int64_t r, int new_len, uint64_t seq16, uint32_t qual16) {
    O.mod_session[rec_idx] = s;
    O.mod_read[rec_idx] = (int32_t)r;
    O.mod_len[rec_idx] = (uint32_t)new_len;
    O.mod_seq_off16[rec_idx] = (uint32_t)seq16;
    O.mod_qual_off16[rec_idx] = qual16;
}


constexpr int kGroupStage = 32;        // words of SNV-masked input staged per lane group (reads up to 256 bases)

// ------------------------------------------------------------------ up to two edits, entirely in registers
// (local-memory arrays are expensive here: with ~216 KB of the SM's 256 KB configured as shared memory there is
// almost no L1 left to hold them)
struct Ed2 {
    int irp[2], len[2], pos[2], p[2], e[2];
    uint32_t mean[2];
    int ne, n_del;
};

// Offsets clamped exactly as Python slicing applies them (anonymizer_methods.py:186-195); returns the new length.
__device__ __forceinline__ int clamp_edits2(Ed2& E, int L) {
    int cur = L;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (q >= E.ne) { E.p[q] = 0x7fffffff; E.e[q] = 0x7fffffff; continue; }
        if (q < E.n_del) {
            E.p[q] = E.irp[q] < cur ? E.irp[q] : cur; E.e[q] = E.p[q] + E.len[q]; cur += E.len[q];
        } else {
            const int pp = E.irp[q] < cur ? E.irp[q] : cur;
            const int ee = E.irp[q] + E.len[q] < cur ? E.irp[q] + E.len[q] : cur;
            E.p[q] = pp; E.e[q] = ee > pp ? ee : pp; cur -= (E.e[q] - pp);
        }
    }
    return cur;
}

// Germline indel edits of modified read k when there are at most two; false otherwise.  Same ordering and
// clamping rules as collect_edits.
template <class SM> __device__ __forceinline__ bool collect2_at(int col_begin, const SM* sm, int k, int L, Ed2& E, int* new_len);
template <class SM> __device__ __forceinline__ bool collect2(const SessCtx& c, const SM* sm, int k, int L, Ed2& E, int* new_len) {
    return collect2_at(c.d.col_begin, sm, k, L, E, new_len);
}
// (the session context stays out of the signature: a caller that reaches this through a call does not have to keep it in memory)
template <class SM> __device__ __forceinline__ bool collect2_at(int col_begin, const SM* sm, int k, int L, Ed2& E, int* new_len) {
    const int oa = sm->mhead[k];
    const int ob = oa >= 0 ? (int)sm->o_rnext[oa] : -1;
    if (ob >= 0 && sm->o_rnext[ob] >= 0) return false;
    int x = oa, y = ob;
    if (y >= 0 && y < x) { const int t = x; x = y; y = t; }          // CIGAR order = slot order
    E.ne = (x >= 0) + (y >= 0);
    const uint32_t mx = x >= 0 ? sm->o_meta[x] : 0u, my = y >= 0 ? sm->o_meta[y] : 0u;
    if (y >= 0 && (mx & kMetaIns) && !(my & kMetaIns)) { const int t = x; x = y; y = t; }   // DELs before INSs (AM.py:264)
    const int o[2] = {x, y};
    E.n_del = 0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const bool has = o[q] >= 0;
        const uint32_t m = has ? sm->o_meta[o[q]] : kMetaIns;
        E.irp[q] = has ? sm->o_irp[o[q]] : 0; E.len[q] = has ? (int)(m & kMetaLenMask) : 0;
        E.pos[q] = has ? sm->o_col[o[q]] + col_begin : 0; E.mean[q] = 0u;
        if (has && !(m & kMetaIns)) ++E.n_del;
    }
    const int cur = clamp_edits2(E, L);
    *new_len = cur;
    return true;
}

// Final index -> original index (>= 0), or -1 - q when the element was inserted by DEL edit q (*kin = offset in it).
__device__ __forceinline__ int map_back2(const Ed2& E, int j, int* kin) {
#pragma unroll
    for (int q = 1; q >= 0; --q)
        if (q < E.ne && q >= E.n_del && j >= E.p[q]) j += E.e[q] - E.p[q];
#pragma unroll
    for (int q = 1; q >= 0; --q)
        if (q < E.n_del) {
            if (j >= E.e[q]) j -= E.len[q];
            else if (j >= E.p[q]) { *kin = j - E.p[q]; return -1 - q; }
        }
    return j;
}

// SNV-masked words of a read, context-free form shared by the session kernel and the emission kernel:
// germ(column relative to col_begin, base code) says whether the allele is germline.
template <class Germ, class Store>
__device__ __forceinline__ void masked_words_g(const BatchView& B, int64_t r, int pos, int L, uint32_t c0, uint32_t c1, int col_begin,
                                               int n_words, int lane, int stride, Germ&& germ, Store&& store) {
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(B.seq4 + 16ull * __ldg(B.seq_off16 + r));
    for (int w = lane; w < n_words; w += stride) {
        const int qb = w << 3;
        uint32_t v = qb < L ? (__ldg(rec + w) & tail_mask(L, w)) : 0u;
        if (qb < L) {
            int rc = pos, q = 0;
            for (uint32_t ci = c0; ci < c1; ++ci) {
                const uint32_t cw = __ldg(B.cigar + ci), op = cw & 15u;
                const int ln = (int)(cw >> 4);
                if (op == 0u || op == 7u || op == 8u) {
                    const int lo = max(q, qb), hi = min(min(q + ln, qb + 8), L);
                    if (lo < hi) {
                        const int p0 = rc - q + qb;                       // reference position of query base qb under this segment
                        const uint32_t fw = ref_word(B.ref4, (int64_t)p0);
                        uint32_t mask = 0xffffffffu;
                        if (lo > qb) mask &= 0xffffffffu << ((lo - qb) * 4);
                        if (hi < qb + 8) mask &= 0xffffffffu >> ((qb + 8 - hi) * 4);
                        uint32_t x = (v ^ fw) & mask;
                        while (x) {
                            const int n = (__ffs(x) - 1) >> 2;
                            x &= ~(0xfu << (n * 4));
                            const uint32_t b = (v >> (n * 4)) & 15u;
                            if (b != 15u && germ(p0 + n - col_begin, b))
                                v = (v & ~(0xfu << (n * 4))) | (((fw >> (n * 4)) & 15u) << (n * 4));
                        }
                    }
                    q += ln; rc += ln;
                } else if (op == 1u || op == 4u) q += ln;
                else if (op == 2u || op == 3u) rc += ln;
                if (q >= qb + 8) break;
            }
        }
        store(w, v);
    }
}

template <class Germ>
__device__ __forceinline__ uint32_t masked_base_g(const BatchView& B, const uint32_t* rec, uint32_t c0, uint32_t c1, int pos, int col_begin, int j, Germ&& germ) {
    const uint32_t b = read_code(rec, j);
    if (b == 15u) return b;
    int rc = pos, q = 0;
    for (uint32_t ci = c0; ci < c1; ++ci) {
        const uint32_t w = __ldg(B.cigar + ci), op = w & 15u;
        const int ln = (int)(w >> 4);
        if (op == 0u || op == 7u || op == 8u) {
            if (j < q + ln) {
                const int rp = rc + (j - q);
                return germ(rp - col_begin, b) ? ref_code(B.ref4, rp) : b;
            }
            q += ln; rc += ln;
        } else if (op == 1u || op == 4u) { if (j < q + ln) return b; q += ln; }
        else if (op == 2u || op == 3u) rc += ln;
    }
    return b;
}

// Indel-masked records with at most two edits, one per group of kGroup lanes (see emit_indel_group_slow for the
// general form and the references).  A word whose bases (qualities) come from consecutive source positions is one
// funnel shift of two staged words; words that straddle an edit are assembled element by element.
// Context-free: used by the session kernel (in-kernel emission) and by the emission kernel.
template <class Germ>
__device__ void emit_indel_group_t(const BatchView& B, ga_totals* totals, const ResultView& O, bool act, const Ed2& E, int64_t r, int col_begin,
                                   int64_t q_lo, int64_t q_hi, uint32_t* stage, uint64_t seq16, uint64_t qual16, int new_len, int glane, Germ&& germ) {
    uint32_t lf = 0u, c0 = 0u, c1 = 0u; int pos = 0;
    if (act) { lf = __ldg(B.len_flag + r); c0 = __ldg(B.cigar_off + r); c1 = __ldg(B.cigar_off + r + 1); pos = __ldg(B.pos + r); }
    const int L = (int)(lf & 0xffffu);
    const uint8_t* qrec = nullptr;
    if (act) {
        qrec = qual_record_in(B, r, q_lo, q_hi);
        if (!qrec) { if (glane == 0) raise_error(totals, GA_ERR_BAD_ARGUMENT, (uint32_t)r); act = false; }
    }
    const bool reverse = ((lf >> 16) & 0x10u) != 0u;
    const uint32_t* rec = act ? reinterpret_cast<const uint32_t*>(B.seq4 + 16ull * __ldg(B.seq_off16 + r)) : nullptr;
    const bool staged = act && ((L + 7) >> 3) <= kGroupStage - 1;
    if (staged) {
        masked_words_g(B, r, pos, L, c0, c1, col_begin, (L + 7) >> 3, glane, kGroup, germ, [&](int w, uint32_t v) { stage[w] = v; });
        if (glane == 0) stage[(L + 7) >> 3] = 0u;
    }
    uint32_t mean0 = 0u, mean1 = 0u;
    {   // quality of re-inserted bases: floor(mean(current qualities)), recomputed after each DEL (AM.py:193)
        uint32_t part = 0;
        if (act) {
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(qrec);
            for (int q = glane; q < ((L + 3) >> 2); q += kGroup) {
                uint32_t v = __ldg(qw + q);
                if (4 * q + 4 > L) v &= 0xffffffffu >> ((4 * q + 4 - L) * 8);
                part += (v & 0xffu) + ((v >> 8) & 0xffu) + ((v >> 16) & 0xffu) + (v >> 24);
            }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1); part += __shfl_xor_sync(0xffffffffu, part, 2); part += __shfl_xor_sync(0xffffffffu, part, 4);
        uint32_t sum = part, n = (uint32_t)L;
        if (E.n_del >= 1) { mean0 = n ? sum / n : 0u; sum += mean0 * (uint32_t)E.len[0]; n += (uint32_t)E.len[0]; }
        if (E.n_del >= 2) { mean1 = n ? sum / n : 0u; }
        if (act && glane == 0) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (q < E.n_del && (int64_t)E.pos[q] + E.len[q] > B.ref_len) raise_error(totals, GA_ERR_LENGTH_MISMATCH, (uint32_t)r);
        }
    }
    __syncwarp();                                                     // staged words visible to the group
    if (!act) return;
    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
    for (int w = glane; w < units * 4; w += kGroup) {
        const int j0 = w << 3;
        uint32_t v = 0u;
        if (j0 < new_len) {
            int kin = 0;
            const int jl = min(j0 + 7, new_len - 1);
            const int s0 = map_back2(E, j0, &kin), s7 = map_back2(E, jl, &kin);
            if (staged && E.ne == 1 && s0 >= 0 && s7 - s0 == jl - j0) {   // one contiguous run (exact for a single edit)
                v = __funnelshift_r(stage[s0 >> 3], stage[(s0 >> 3) + 1], (uint32_t)(s0 & 7) * 4u);
                if (jl - j0 < 7) v &= 0xffffffffu >> ((7 - (jl - j0)) * 4);
            } else {
                for (int n = 0; n <= jl - j0; ++n) {
                    const int src = map_back2(E, j0 + n, &kin);
                    uint32_t code;
                    if (src < 0) code = ref_code(B.ref4, (int64_t)(src == -1 ? E.pos[0] : E.pos[1]) + kin);
                    else if (staged) code = (stage[src >> 3] >> ((src & 7) * 4)) & 15u;
                    else code = masked_base_g(B, rec, c0, c1, pos, col_begin, src, germ);
                    v |= code << (n * 4);
                }
            }
        }
        oseq[w] = v;
    }
    uint32_t* oq = reinterpret_cast<uint32_t*>(O.out_qual + 32ull * qual16);
    const uint32_t* qw = reinterpret_cast<const uint32_t*>(qrec);
    for (int w = glane; w < units * 8; w += kGroup) {
        const int p0 = w << 2;
        uint32_t v = 0u;
        if (p0 < new_len) {
            const int pl = min(p0 + 3, new_len - 1);
            int kin = 0;
            const int f0 = reverse ? new_len - 1 - p0 : p0, f3 = reverse ? new_len - 1 - pl : pl;
            const int s0 = map_back2(E, f0, &kin), s3 = map_back2(E, f3, &kin);
            const int b0 = reverse ? L - 1 - s0 : s0;                   // byte of the BAM-order quality record
            if (E.ne == 1 && s0 >= 0 && s3 >= 0 && (reverse ? s0 - s3 : s3 - s0) == pl - p0) {
                const uint32_t lo = __ldg(qw + (b0 >> 2)), hi = (b0 & 3) ? __ldg(qw + (b0 >> 2) + 1) : 0u;
                v = __funnelshift_r(lo, hi, (uint32_t)(b0 & 3) * 8u);
                if (pl - p0 < 3) v &= 0xffffffffu >> ((3 - (pl - p0)) * 8);
            } else {
                for (int n = 0; n <= pl - p0; ++n) {
                    const int f = reverse ? new_len - 1 - (p0 + n) : p0 + n;
                    const int src = map_back2(E, f, &kin);
                    const uint32_t qv = src >= 0 ? (uint32_t)qrec[reverse ? L - 1 - src : src] : (src == -1 ? mean0 : mean1);
                    v |= qv << (n * 8);
                }
            }
        }
        oq[w] = v;
    }
}

// ------------------------------------------------------------------ more than two edits (rare): the edit list
// travels through a global side buffer, one uint4 {p_eff, e_eff, length, reference position} per edit in application
// order, and every output element is mapped back through it.  Code size over speed: these are out-of-line.
constexpr int kManyEdits = 16;         // germline indels of one read handled by the streaming pipeline (more: fallback kernel)

// The germline indel edits of modified read k (any number up to kManyEdits) in application order with the Python-slice
// clamping of anonymizer_methods.py:186-195 (same rules as clamp_edits2).  Returns the number of edits (0: too many),
// writes them to dst when dst is not null.
template <class SM> __device__ __noinline__ int collect_many(const SessCtx& c, const SM* sm, int k, int L, uint4* dst, int* new_len, int* n_del_out) {
    int slots[kManyEdits];
    int ne = 0;
    for (int o = sm->mhead[k]; o >= 0; o = sm->o_rnext[o]) {
        if (ne == kManyEdits) return 0;
        int a = ne++;
        while (a > 0 && slots[a - 1] > o) { slots[a] = slots[a - 1]; --a; }    // CIGAR order = slot order
        slots[a] = o;
    }
    int cur = L, q = 0, n_del = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass)                                      // all DELs, then all INSs (AM.py:264)
#pragma unroll 1
        for (int j = 0; j < ne; ++j) {
            const uint32_t m = sm->o_meta[slots[j]];
            if (((m & kMetaIns) != 0u) != (pass == 1)) continue;
            const int irp = sm->o_irp[slots[j]], len = (int)(m & kMetaLenMask);
            int p, e;
            if (pass == 0) { p = irp < cur ? irp : cur; e = p + len; cur += len; ++n_del; }
            else {
                p = irp < cur ? irp : cur;
                const int ee = irp + len < cur ? irp + len : cur;
                e = ee > p ? ee : p; cur -= e - p;
            }
            if (dst) dst[q] = make_uint4((uint32_t)p, (uint32_t)e, (uint32_t)len, (uint32_t)(sm->o_col[slots[j]] + c.d.col_begin));
            ++q;
        }
    *new_len = cur; *n_del_out = n_del;
    return ne;
}

// Final index -> original index (>= 0), or -1 - q when DEL edit q inserted the element (*kin = offset in it).
__device__ __forceinline__ int map_back_many(const uint4* ed, int n_del, int ne, int j, int* kin) {
#pragma unroll 1
    for (int q = ne - 1; q >= n_del; --q) { const uint4 e = __ldg(ed + q); if (j >= (int)e.x) j += (int)e.y - (int)e.x; }
#pragma unroll 1
    for (int q = n_del - 1; q >= 0; --q) {
        const uint4 e = __ldg(ed + q);
        if (j >= (int)e.y) j -= (int)e.z;
        else if (j >= (int)e.x) { *kin = j - (int)e.x; return -1 - q; }
    }
    return j;
}

// One indel-masked record with more than two edits per group of G lanes (G = 8, 16 or 32); every lane of the warp
// calls it (act = false for groups without such a record).  Same rules as emit_indel_group_t: SNV masking first
// (anonymizer_methods.py:170-176), deleted reference bases come back with quality floor(mean(current qualities))
// (:193), inserted bases and their qualities go (:186-187), qualities are edited in forward orientation and printed
// in BAM order (:95, 213).
template <int G, class Germ>
__device__ __noinline__ void emit_many_group(const BatchView& B, ga_totals* totals, const ResultView& O, bool act, const uint4* ed, int ne, int n_del,
                                             int64_t r, int pos, int L, uint32_t src_unit, uint32_t c0, uint32_t c1, bool reverse, int col_begin,
                                             const uint8_t* qrec, uint64_t seq16, uint64_t qual16, int new_len, int glane, Germ germ) {
    if (act && !qrec) { if (glane == 0) raise_error(totals, GA_ERR_BAD_ARGUMENT, (uint32_t)r); act = false; }
    uint32_t part = 0;
    if (act) {
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(qrec);
        for (int q = glane; q < ((L + 3) >> 2); q += G) {
            uint32_t v = __ldg(qw + q);
            if (4 * q + 4 > L) v &= 0xffffffffu >> ((4 * q + 4 - L) * 8);
            part += __vsadu4(v, 0u);
        }
    }
#pragma unroll
    for (int d = 1; d < G; d <<= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (!act) return;
    auto mean_of = [&](int q) -> uint32_t {                               // recomputed after each DEL (AM.py:193)
        uint32_t sum = part, n = (uint32_t)L, m = 0u;
        for (int t = 0; t <= q; ++t) { m = n ? sum / n : 0u; const uint32_t len = __ldg(ed + t).z; sum += m * len; n += len; }
        return m;
    };
    if (glane == 0)
        for (int q = 0; q < n_del; ++q) { const uint4 e = __ldg(ed + q); if ((int64_t)(int)e.w + (int)e.z > B.ref_len) raise_error(totals, GA_ERR_LENGTH_MISMATCH, (uint32_t)r); }
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(B.seq4 + 16ull * src_unit);
    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
    for (int w = glane; w < units * 4; w += G) {
        uint32_t v = 0u;
        for (int n = 0; n < 8 && 8 * w + n < new_len; ++n) {
            int kin = 0;
            const int src = map_back_many(ed, n_del, ne, 8 * w + n, &kin);
            const uint32_t code = src >= 0 ? masked_base_g(B, rec, c0, c1, pos, col_begin, src, germ)
                                           : ref_code(B.ref4, (int64_t)(int)__ldg(ed + (-1 - src)).w + kin);
            v |= code << (n * 4);
        }
        oseq[w] = v;
    }
    uint32_t* oq = reinterpret_cast<uint32_t*>(O.out_qual + 32ull * qual16);
    for (int w = glane; w < units * 8; w += G) {
        uint32_t v = 0u;
        for (int n = 0; n < 4 && 4 * w + n < new_len; ++n) {
            const int f = reverse ? new_len - 1 - (4 * w + n) : 4 * w + n;
            int kin = 0;
            const int src = map_back_many(ed, n_del, ne, f, &kin);
            const uint32_t qv = src >= 0 ? (uint32_t)qrec[reverse ? L - 1 - src : src] : mean_of(-1 - src);
            v |= qv << (n * 8);
        }
        oq[w] = v;
    }
}

// In-kernel form: the germline test is the session's shared-memory table.

}  // namespace ga
