// ga_resolve_kernel.cuh - stage 2 of the streaming pipeline: one warp (lean variant, below) or one small CTA (large
// variant, for sessions whose tables do not fit the lean capacities) per session turns the scan kernel's
// candidate entries and indel observations into the germline set (seen in tumor AND normal, minus
// variant_to_keep: anonymizer_methods.py:546-547, variants.py:83-96), the ordered list of modified reads, their
// new lengths and their slots in the compacted output (north_star jobs (1) + (4)); it writes the record headers,
// the per-session counters (= AnonymizedVariantsStatistics.window_var_counts) and the hand-over to the emission
// kernel (ga_emit_kernel.cuh), which writes the record bodies.
//
// Large variant: 128 threads, ~30 KB of shared memory -> 7 CTAs per SM; per session a few dependent memory round
// trips (descriptor + counts, entries, per-record meta) and a handful of 4-warp barriers.
// Anything the shared-memory tables cannot hold (or describe to the emission kernel) sends the whole session to
// the global-scratch fallback kernel (ga_session_kernel.cuh) through big_list; n_big[4 + reason] counts why
// (1 scan-kernel overflow, 2 IUPAC read base, 3 more modified reads / germline alleles than the tables hold,
// 4 a read with more than two germline indels; 0 = oversize session found by the assignment kernel).
#pragma once
#include "ga_scan_kernel.cuh"

namespace ga {

constexpr int kResThreads = 128;
constexpr int kEntR = 1536;              // entries staged in shared memory; sessions with more stream them from the scan kernel's regions
constexpr int kObsR = 2 * kObsHalf;
constexpr int kHashR = 256;
constexpr int kGermStride = kGermCap + 4;   // per session: [0] germline SNV alleles, [1] col_begin, [4..] (column << 4) | base code
static_assert(kReads2 / 32 <= kResThreads, "phase L gives every bitmap word its own thread");

// Hand-over to the emission kernel.  kind[k] of output record k: 0 = nothing to do (record written by the fallback
// kernel), 1 = clean read, SNV-only, at most two germline hits: copy + patch from the descriptor, 2 = other CIGAR,
// SNV-only: re-walk, 3 = indel-masked with at most two edits (their description sits in the first 32 bytes of the
// record's out_qual slot until the emission kernel overwrites it with the qualities), 4 = clean read, SNV-only:
// copy, looking every mismatch up in the session's germline list.
struct EmitScratch2 {
    uint32_t* kind1_list; // [cap_records] record indices of the long clean reads (kind 1), appended through n_kind1: what emit_kernel walks
    uint4* edesc;        // [cap_records] kind 1: {source record unit, pos - col_begin, length | hits << 16, two hits ((column << 4) | reference code)}
                         //               else:   {source record unit, pos, length, session}
    uint32_t* germ;      // [n_sessions][kGermStride]
    uint4* sdesc;        // [cap_records][4] everything emit_special_kernel needs about one record of kind >= 2, packed densely:
                         //   {source record unit, pos, length | kind << 16 | reverse << 20, session}
                         //   {read, new length, output sequence unit, output quality unit}
                         //   {first CIGAR word, one past the last, col_begin, -}   {quality-index slice begin, end, -, -}
                         // kind 0 marks a slot whose session did not fit the caller's capacities
    uint32_t* n_special;
    uint4* many;         // [cap_many] edit lists of the records with more than two germline indels (collect_many)
    uint32_t* n_many;
    uint32_t cap_many;
    uint32_t* many_recs; // [cap_many] special-record slots of those records (emit_many_kernel walks them)
    uint32_t* n_many_recs;
    uint32_t* n_kind1;   // records left to emit_kernel (long clean reads); zero lets that kernel return at once
    unsigned int* ticket_large;   // next entry of large_list for the one-CTA resolve kernel
    unsigned int* ticket_lean;    // next session of the one-warp resolve kernel
    uint32_t* n_rare;             // special records that emit_records_kernel leaves to emit_special_kernel (two edits, many hits, reads beyond 160 bases)
    uint32_t* rare_list;          // [cap_records] their slots in sdesc
    uint4* edit_keep;             // null, or [cap_records][2]: a lasting copy of the EditAux of every indel-masked record, by record index
                                  // (ga_record_edits; all bits set = not kept: more than two edits)
};

// Reserves and writes the edit list of modified read k, which has more than two germline indels; the offset is parked
// in mpatch[k] (unused by indel-masked reads).  False: more than kManyEdits edits or no room - the fallback kernel
// takes the session.
template <class SM> __device__ __noinline__ bool reserve_many(const EmitScratch2& E, const SessCtx& c, SM* sm, int k, int L, int* new_len) {
    int nd = 0;
    const int ne = collect_many(c, sm, k, L, (uint4*)nullptr, new_len, &nd);
    if (!ne) return false;
    const uint32_t off = atomicAdd(E.n_many, (uint32_t)ne);
    if (off + (uint32_t)ne > E.cap_many) return false;
    collect_many(c, sm, k, L, E.many + off, new_len, &nd);
    sm->mpatch[k] = off;
    return true;
}

// The hand-over of such a record (EditAux layout; ne > 2 says that irp0 is the offset of the edit list).
template <class SM> __device__ __noinline__ void write_many_aux(const SessCtx& c, const SM* sm, int k, int L, uint32_t qidx, uint4* dst) {
    int nl = 0, nd = 0;
    const int ne = collect_many(c, sm, k, L, (uint4*)nullptr, &nl, &nd);
    dst[0] = make_uint4(sm->mpatch[k], 0u, 0u, 0u);
    dst[1] = make_uint4(0u, 0u, (uint32_t)ne | ((uint32_t)nd << 8), qidx);
}

__device__ __forceinline__ void write_special(const EmitScratch2& E, const BatchView& B, const SessionDesc& d, uint32_t slot, uint32_t kind,
                                              uint32_t so, int pos, uint32_t lf, int s, int64_t r, uint32_t new_len, uint64_t seq16, uint32_t qual16) {
    uint4* dp = E.sdesc + 4ull * slot;
    const bool tumor = r < B.n_tumor;
    dp[0] = make_uint4(so, (uint32_t)pos, (lf & 0xffffu) | (kind << 16) | (((lf >> 20) & 1u) << 20), (uint32_t)s);
    dp[1] = make_uint4((uint32_t)r, new_len, (uint32_t)seq16, qual16);
    dp[2] = make_uint4(__ldg(B.cigar_off + r), __ldg(B.cigar_off + r + 1), (uint32_t)d.col_begin, 0u);
    dp[3] = make_uint4((uint32_t)(tumor ? d.qt_begin : d.qn_begin), (uint32_t)(tumor ? d.qt_end : d.qn_end), 0u, 0u);
}

// The body of a common record - clean read, at most two germline hits, at most 160 bases (five 16-byte units) - written
// by the resolve kernels themselves: copy, base <- reference base at the hits (anonymizer_methods.py:170-176).  hits =
// two (column << 4 | reference code) half-words, rel0 = pos - col_begin.  The loads of the records a warp handles in one
// step are in flight together.
__device__ __forceinline__ void write_common_body(const BatchView& B, const ResultView& O, uint32_t so, uint64_t dst_unit, uint32_t L0,
                                                  uint32_t hits, uint32_t n_hits, int rel0) {
    const uint4* src = reinterpret_cast<const uint4*>(B.seq4 + 16ull * so);
    uint4* dst = reinterpret_cast<uint4*>(O.out_seq4 + 16ull * dst_unit);
    const int nu = (int)((L0 + 31u) >> 5);
    uint4 v[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) v[u] = u < nu ? ldg128(src + u) : make_uint4(0u, 0u, 0u, 0u);
    const int q0 = (int)((hits >> 4) & 0xfffu) - rel0, q1 = n_hits > 1u ? (int)(hits >> 20) - rel0 : -1;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
        if (u >= nu) break;
        uint4 w = v[u];
        if (32 * u + 32 > (int)L0) { w.x &= tail_mask((int)L0, 4 * u); w.y &= tail_mask((int)L0, 4 * u + 1); w.z &= tail_mask((int)L0, 4 * u + 2); w.w &= tail_mask((int)L0, 4 * u + 3); }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = h ? q1 : q0;
            if (q >= 0 && (q >> 5) == u) {
                const uint32_t sh = (uint32_t)(q & 7) * 4u, keepm = ~(0xfu << sh), ins = ((h ? hits >> 16 : hits) & 15u) << sh;
                const int ww = (q >> 3) & 3;
                if (ww == 0) w.x = (w.x & keepm) | ins; else if (ww == 1) w.y = (w.y & keepm) | ins;
                else if (ww == 2) w.z = (w.z & keepm) | ins; else w.w = (w.w & keepm) | ins;
            }
        }
        dst[u] = w;
    }
}

struct SmemR {
    uint32_t tab[kCols2 / 4];            // one byte per column: bits 0-3 tumor saw A,C,G,T, bits 4-7 normal
    uint32_t ent[kEntR];                 // entries of the tumor item, then of the normal item
    uint32_t msize[kMod2], mseq[kMod2], mqual[kMod2];
    uint16_t clist[kMod2];
    int16_t mhead[kMod2];                // per modified read: chain of its germline indel observations
    uint32_t modbits[kReads2 / 32], indelbits[kReads2 / 32], genbits[kReads2 / 32], woff[kReads2 / 32];
    int32_t o_col[kObsR];
    uint32_t o_meta[kObsR];
    uint32_t o_read[kObsR];              // session-relative read (low 16) | allele length (high 16)
    int32_t o_irp[kObsR];
    uint32_t o_sig0[kObsR], o_sig1[kObsR];
    int16_t o_rep[kObsR], o_rnext[kObsR]; // o_rep: the first observation with the same key (its representative)
    uint32_t kt[kHashR];                 // per hash bucket of the key: the smallest observation in it
    uint32_t mpatch[kMod2];              // two germline hits of a clean read: (column << 4) | reference code, 16 bits each
    uint8_t mpc[kMod2];                  // germline SNV hits per modified read
};

template <int T>
__device__ __forceinline__ uint32_t block_scan32(uint32_t v, uint32_t* tmp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += n; }
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    uint32_t before = 0u, all = 0u;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) { const uint32_t x = tmp[w]; if (w < warp) before += x; all += x; }
    *total = all;
    return before + inc - v;
}
template <int T>
__device__ __forceinline__ unsigned long long block_scan64(unsigned long long v, unsigned long long* tmp, unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += n; }
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    unsigned long long before = 0ull, all = 0ull;
#pragma unroll
    for (int w = 0; w < T / 32; ++w) { const unsigned long long x = tmp[w]; if (w < warp) before += x; all += x; }
    *total = all;
    return before + inc - v;
}

__device__ __forceinline__ int acgt_index(uint32_t b) { return b == 1u ? 0 : b == 2u ? 1 : b == 4u ? 2 : b == 8u ? 3 : -1; }

__global__ void __launch_bounds__(kResThreads, 6) resolve_kernel(BatchView B, SessView S, const SessionDesc* __restrict__ descs,
                                                                 int32_t* __restrict__ big_list, int32_t* __restrict__ n_big,
                                                                 const int32_t* __restrict__ large_list, const int32_t* __restrict__ n_large,
                                                                 ResultView O, ScanScratch X, EmitScratch2 E) {
    constexpr int T = kResThreads;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SmemR* sm = reinterpret_cast<SmemR*>(smem_raw);
    __shared__ unsigned long long s_scan64[2][T / 32];
    __shared__ uint32_t s_scan[T / 32];
    __shared__ uint32_t s_cnt[3], s_ngerm, s_overflow;
    __shared__ unsigned long long s_base[3];

    SessCtx c;
    c.B = B;
    c.totals = O.totals;
    memset(&c.T, 0, sizeof c.T);
    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t round = 0;

    const int n_list = *n_large;                                      // sessions the lean kernel handed over
    __shared__ int s_li;
    for (;; ++round) {                                                // sessions differ in cost: a ticket, not a stride
        if (tid == 0) s_li = (int)atomicAdd(E.ticket_large, 1u);
        __syncthreads();
        const int li = s_li;
        __syncthreads();
        if (li >= n_list) break;
        const int s = large_list[li];
        c.d = descs[s];
        if (c.d.big) continue;                                        // listed by the assignment kernel
        const uint4 cnt0 = X.cnt[2 * (size_t)s], cnt1 = X.cnt[2 * (size_t)s + 1];
        c.s = s;
        c.nt = c.d.t_end - c.d.t_begin;
        c.n_range = c.nt + (c.d.n_end - c.d.n_begin);
        const int n_cols = c.d.n_cols;
        const int n_cw = (c.n_range + 31) >> 5;
        if (cnt0.x == kCntOverflow || cnt1.x == kCntOverflow) {       // the scan kernel could not hold the session
            if (tid == 0) { big_list[atomicAdd(n_big, 1)] = s; atomicAdd(n_big + 4 + 1, 1); }
            continue;
        }
        const int n_ent0 = (int)cnt0.x, n_ent = n_ent0 + (int)cnt1.x;
        // staged in shared memory when they fit, streamed from the scan kernel's regions otherwise
        const uint32_t* ent0 = X.ent + (size_t)(2 * (size_t)s) * kEntHalf;
        const bool ent_staged = n_ent <= kEntR;
        auto ent_at = [&](int k) -> uint32_t { return ent_staged ? sm->ent[k] : (k < n_ent0 ? __ldg(ent0 + k) : __ldg(ent0 + kEntHalf + (k - n_ent0))); };
        const int n_obs0 = (int)cnt0.y, n_obs = n_obs0 + (int)cnt1.y;
        const uint32_t sess_reads = cnt0.z + cnt1.z, sess_bases = cnt0.w + cnt1.w;
        c.first = S.first[s];
        c.keep_type = S.keep_type[s]; c.keep_pos = S.keep_pos[s]; c.keep_end = S.keep_end[s]; c.keep_len = S.keep_len[s];
        const uint32_t ka0 = S.keep_allele_off[s];
        c.keep_allele = S.keep_alleles + ka0;
        c.keep_alen = (int)(S.keep_allele_off[s + 1] - ka0);

        // ---- stage the session: zeroed tables, entries and observations from the scan kernel's regions
        for (int k = tid; k < ((n_cols + 3) >> 2); k += T) sm->tab[k] = 0u;
        for (int k = tid; k < n_cw; k += T) { sm->modbits[k] = 0u; sm->indelbits[k] = 0u; sm->genbits[k] = 0u; }
        if (n_obs > 0) for (int k = tid; k < kHashR; k += T) sm->kt[k] = 0xffffffffu;
        if (tid == 0) { s_cnt[0] = s_cnt[1] = s_cnt[2] = 0u; s_ngerm = 0u; s_overflow = 0u; }
        {
            const uint32_t* e0 = X.ent + (size_t)(2 * (size_t)s) * kEntHalf;
            const uint32_t* e1 = e0 + kEntHalf;
            if (n_ent <= kEntR) for (int k = tid; k < n_ent; k += T) sm->ent[k] = k < n_ent0 ? __ldg(e0 + k) : __ldg(e1 + (k - n_ent0));
            const ObsRec* o0 = X.obs + (size_t)(2 * (size_t)s) * kObsHalf;
            const ObsRec* o1 = o0 + kObsHalf;
            for (int o = tid; o < n_obs; o += T) {
                const uint4* src = reinterpret_cast<const uint4*>(o < n_obs0 ? o0 + o : o1 + (o - n_obs0));
                const uint4 a = __ldg(src), b = __ldg(src + 1);
                sm->o_col[o] = (int32_t)a.x; sm->o_meta[o] = a.y; sm->o_read[o] = a.z; sm->o_irp[o] = (int32_t)a.w;
                sm->o_sig0[o] = b.x; sm->o_sig1[o] = b.y;
            }
        }
        // variant_to_keep as an SNV entry key (anonymizer_methods.py:546-547 compares with CalledGenomicVariant.__eq__)
        uint32_t keep_key = 0xffffffffu;
        if (c.keep_type == GA_VT_SNV && c.keep_end == c.keep_pos && c.keep_len == 1 && c.keep_alen == 1) {
            const int kc = c.keep_pos - c.d.col_begin;
            if (kc >= 0 && kc < n_cols) {
                const uint8_t ch = c.keep_allele[0];
                const uint32_t code = ch == 'A' ? 1u : ch == 'C' ? 2u : ch == 'G' ? 4u : ch == 'T' ? 8u : 0u;
                if (code) keep_key = ((uint32_t)kc << 4) | code;
            }
        }
        __syncthreads();

        // ---- build: allele table and observation chains
        for (int k = tid; k < n_ent; k += T) {
            const uint32_t e = ent_at(k);
            const uint32_t col = (e >> 4) & 0xfffu;
            const int idx = acgt_index(e & 15u);
            if (idx < 0) { s_overflow = 1u; continue; }                // IUPAC read base: the fallback kernel keeps all 16 codes
            atomicOr(&sm->tab[col >> 2], 1u << (idx + (k >= n_ent0 ? 4 : 0) + 8 * (int)(col & 3u)));
        }
        auto obs_hash = [&](int o) -> uint32_t {                         // equal keys (variants.py:83-96) hash alike
            uint32_t kh = (uint32_t)sm->o_col[o] * 0x9E3779B1u ^ ((sm->o_meta[o] & (kMetaIns | kMetaLenMask)) * 0x85EBCA77u) ^ ((sm->o_read[o] >> 16) * 0xC2B2AE3Du) ^
                          (sm->o_sig0[o] * 0x27D4EB2Fu) ^ (sm->o_sig1[o] * 0x165667B1u);
            return (kh ^ (kh >> 15)) & (kHashR - 1);
        };
        for (int o = tid; o < n_obs; o += T) atomicMin(&sm->kt[obs_hash(o)], (uint32_t)o);
        __syncthreads();
        if (s_overflow) {
            __syncthreads();
            if (tid == 0) { big_list[atomicAdd(n_big, 1)] = s; atomicAdd(n_big + 4 + 2, 1); }
            continue;
        }

        // ---- resolve + mark: germline = seen in tumor AND normal, minus variant_to_keep
        for (int k = tid; k < n_ent; k += T) {
            const uint32_t e = ent_at(k);
            const uint32_t col = (e >> 4) & 0xfffu;
            const uint32_t byte = (sm->tab[col >> 2] >> (8 * (col & 3u))) & 0xffu;
            const int idx = acgt_index(e & 15u);
            if ((((byte & (byte >> 4)) >> idx) & 1u) && (e & 0xffffu) != keep_key) {
                const uint32_t i = (e >> 16) & 0xfffu;
                atomicOr(&sm->modbits[i >> 5], 1u << (i & 31));
                if (e & kEntGen) atomicOr(&sm->genbits[i >> 5], 1u << (i & 31));
            }
        }
        {   // distinct germline SNV alleles: the per-session counter and the emission kernel's list
            uint32_t cnt = 0;
            for (int k = tid; k < ((n_cols + 3) >> 2); k += T) {
                const uint32_t w = sm->tab[k];
                uint32_t g = w & (w >> 4) & 0x0f0f0f0fu;
                while (g) {
                    const int bit = __ffs(g) - 1; g &= g - 1;
                    const uint32_t key = ((uint32_t)(4 * k + (bit >> 3)) << 4) | (1u << (bit & 7));
                    if (key == keep_key) continue;
                    ++cnt;
                    const uint32_t slot = atomicAdd(&s_ngerm, 1u);
                    if (slot < (uint32_t)kGermCap) E.germ[(size_t)s * kGermStride + 4 + slot] = key;
                }
            }
            cnt = warp_sum(cnt);
            if (lane == 0 && cnt) atomicAdd(&s_cnt[0], cnt);
        }
        // indels: exact key equality (variants.py:83-96).  Equal keys share a hash bucket, so an observation's
        // representative (the first observation equal to it) is the bucket's smallest member unless two keys met in the
        // bucket; the representative collects which datasets showed the key.
        for (int o = tid; o < n_obs; o += T) {
            const int col = sm->o_col[o];
            const int cand = (int)sm->kt[obs_hash(o)];
            int rep = o;
            if (cand != o) {
                if (sm->o_col[cand] == col && obs_equal2(c, sm, o, cand)) rep = cand;
                else for (int j = 0; j < o; ++j) if (sm->o_col[j] == col && obs_equal2(c, sm, o, j)) { rep = j; break; }
            }
            sm->o_rep[o] = (int16_t)rep;
            {
                const uint32_t mo = sm->o_meta[o];                        // a normal read that shows the key and covers its position: the normal column exists
                atomicOr(&sm->o_meta[rep], (mo & kMetaDs) ? (kMetaSeenN | ((mo & kMetaTrail) ? 0u : kMetaCol)) : kMetaSeenT);
            }
        }
        __syncthreads();
        for (int o = tid; o < n_obs; o += T) {
            const uint32_t m = sm->o_meta[o];
            const int rep = sm->o_rep[o];
            bool germ = (sm->o_meta[rep] & (kMetaSeenT | kMetaSeenN)) == (kMetaSeenT | kMetaSeenN);
            if (germ && !(sm->o_meta[rep] & kMetaCol)) germ = normal_covers(c, c.first, sm->o_col[o] + c.d.col_begin);   // anonymizer_methods.py:474-481
            if (germ && obs_equals_keep2(c, sm, o)) germ = false;
            if (germ) {
                atomicOr(&sm->o_meta[o], kMetaGerm | (rep == o ? kMetaRep : 0u));
                if (rep == o) atomicAdd(&s_cnt[(m & kMetaIns) ? 2 : 1], 1u);
                const uint32_t i = (uint32_t)obs_read(sm, o);
                atomicOr(&sm->modbits[i >> 5], 1u << (i & 31));
                atomicOr(&sm->indelbits[i >> 5], 1u << (i & 31));
            }
        }
        __syncthreads();

        // ---- ordered list of the modified reads (one bitmap word per thread)
        uint32_t n_mod;
        {
            const uint32_t bits = tid < n_cw ? sm->modbits[tid] : 0u;
            uint32_t off = block_scan32<T>(__popc(bits), s_scan, &n_mod);
            if (tid < n_cw) sm->woff[tid] = off;
            uint32_t b = bits;
            while (b) {
                const int k = __ffs(b) - 1; b &= b - 1;
                if (off < (uint32_t)kMod2) { sm->clist[off] = (uint16_t)(tid * 32 + k); sm->mhead[off] = (int16_t)-1; sm->mpc[off] = 0; }
                ++off;
            }
        }
        if (n_mod > (uint32_t)kMod2 || s_ngerm > (uint32_t)kGermCap) {
            __syncthreads();
            if (tid == 0) { big_list[atomicAdd(n_big, 1)] = s; atomicAdd(n_big + 4 + 3, 1); }
            continue;
        }
        __syncthreads();
        if (n_obs > 0) {                                              // hang every germline observation on its modified read
            for (int o = tid; o < n_obs; o += T) {
                if (!(sm->o_meta[o] & kMetaGerm)) continue;
                const uint32_t i = (uint32_t)obs_read(sm, o);
                const uint32_t k = sm->woff[i >> 5] + __popc(sm->modbits[i >> 5] & ((1u << (i & 31)) - 1u));
                // 16-bit exchange on the containing word
                uint32_t* hw = reinterpret_cast<uint32_t*>(sm->mhead) + (k >> 1);
                const int shift = (int)(k & 1u) * 16;
                uint32_t old = *hw, assumed;
                do {
                    assumed = old;
                    sm->o_rnext[o] = (int16_t)((assumed >> shift) & 0xffffu);
                    __threadfence_block();
                    old = atomicCAS(hw, assumed, (assumed & ~(0xffffu << shift)) | (((uint32_t)o & 0xffffu) << shift));
                } while (old != assumed);
            }
            __syncthreads();
        }

        // ---- the germline hits of every clean modified read (they travel in the emission descriptor)
        for (int k = tid; k < n_ent; k += T) {
            const uint32_t e = ent_at(k);
            if (e & kEntGen) continue;
            const uint32_t col = (e >> 4) & 0xfffu;
            const uint32_t byte = (sm->tab[col >> 2] >> (8 * (col & 3u))) & 0xffu;
            if (!((((byte & (byte >> 4)) >> acgt_index(e & 15u)) & 1u) && (e & 0xffffu) != keep_key)) continue;
            const uint32_t i = (e >> 16) & 0xfffu;
            const uint32_t m = sm->woff[i >> 5] + __popc(sm->modbits[i >> 5] & ((1u << (i & 31)) - 1u));
            uint32_t* cw = reinterpret_cast<uint32_t*>(sm->mpc) + (m >> 2);
            const uint32_t have = (atomicAdd(cw, 1u << (8 * (m & 3u))) >> (8 * (m & 3u))) & 0xffu;   // at most kGermCap hits per read: no carry
            if (have < 2u) reinterpret_cast<uint16_t*>(sm->mpatch)[2 * m + have] = (uint16_t)((col << 4) | (1u << ((e >> 29) & 3u)));
        }
        __syncthreads();
        // ---- new length of every modified read; indel-masked reads need the edit analysis
        const int per = ((int)n_mod + T - 1) / T;
        const int k0 = min(tid * per, (int)n_mod), k1 = min(k0 + per, (int)n_mod);
        unsigned long long mine = 0ull;                               // [records:16 | seq units:24 | qual units:24]
        uint32_t n_q = 0;
        for (int k = k0; k < k1; ++k) {
            const int i = (int)sm->clist[k];
            const int64_t r = read_of(c, i);
            const int L0 = (int)(__ldg(B.len_flag + r) & 0xffffu);
            uint32_t m;
            if ((sm->indelbits[i >> 5] >> (i & 31)) & 1u) {
                Ed2 E2;
                int new_len = L0;
                if (!collect2(c, sm, k, L0, E2, &new_len) && !reserve_many(E, c, sm, k, L0, &new_len)) s_overflow = 1u;   // the fallback kernel takes the session
                m = kModFlag | kQualFlag | ((uint32_t)new_len & kLen2);
                ++n_q;
            } else {
                m = kModFlag | (uint32_t)L0;
            }
            sm->msize[k] = m;
            uint32_t units = ((m & kLen2) + 31u) / 32u; if (units < 1u) units = 1u;
            mine += (1ull << 48) | ((unsigned long long)units << 24) | ((m & kQualFlag) ? (unsigned long long)units : 0ull);
        }
        unsigned long long total;
        const unsigned long long off = block_scan64<T>(mine, s_scan64[round & 1u], &total);
        if (s_overflow) {
            __syncthreads();
            if (tid == 0) { big_list[atomicAdd(n_big, 1)] = s; atomicAdd(n_big + 4 + 4, 1); }
            continue;
        }
        const uint32_t tot_rec = (uint32_t)(total >> 48), tot_seq = (uint32_t)((total >> 24) & 0xffffffu), tot_qual = (uint32_t)(total & 0xffffffu);
        if (tid == 0) {
            s_base[0] = atomicAdd((unsigned long long*)&O.totals->n_modified, (unsigned long long)tot_rec);
            s_base[1] = atomicAdd((unsigned long long*)&O.totals->seq16_used, (unsigned long long)tot_seq);
            s_base[2] = atomicAdd((unsigned long long*)&O.totals->qual16_used, (unsigned long long)tot_qual);
            atomicAdd((unsigned long long*)&O.totals->session_reads, (unsigned long long)sess_reads);
            atomicAdd((unsigned long long*)&O.totals->session_bases, (unsigned long long)sess_bases);
            for (int k = 0; k < 3; ++k) {
                O.sess_counts[4 * (size_t)s + k] = s_cnt[k];
                if (s_cnt[k]) atomicAdd((unsigned long long*)&O.totals->masked[k], (unsigned long long)s_cnt[k]);
            }
            O.sess_counts[4 * (size_t)s + 3] = sess_reads;
            E.germ[(size_t)s * kGermStride] = s_ngerm;
            E.germ[(size_t)s * kGermStride + 1] = (uint32_t)c.d.col_begin;
        }
        {   // per-record output offsets (session-relative)
            uint32_t so = (uint32_t)((off >> 24) & 0xffffffu), qo = (uint32_t)(off & 0xffffffu);
            for (int k = k0; k < k1; ++k) {
                const uint32_t m = sm->msize[k];
                uint32_t units = ((m & kLen2) + 31u) / 32u; if (units < 1u) units = 1u;
                sm->mseq[k] = so; sm->mqual[k] = qo;
                so += units; if (m & kQualFlag) qo += units;
            }
        }
        n_q = warp_sum(n_q);
        if (lane == 0 && n_q) atomicAdd((unsigned long long*)&O.totals->indel_records, (unsigned long long)n_q);
        __syncthreads();
        const bool fits = (int64_t)(s_base[0] + tot_rec) <= O.cap_records && (int64_t)(s_base[1] + tot_seq) <= O.cap_seq16 &&
                          (int64_t)(s_base[2] + tot_qual) <= O.cap_qual16;
        if (!fits) { if (tid == 0) raise_error(O.totals, GA_ERR_CAPACITY, 0xffffffffu); __syncthreads(); continue; }

        // ---- record headers and the hand-over to the emission kernel
        for (int k = tid; k < (int)n_mod; k += T) {
            const uint32_t m = sm->msize[k];
            const bool q = (m & kQualFlag) != 0u;
            const int i = (int)sm->clist[k];
            const int64_t r = read_of(c, i);
            const uint64_t rec_idx = s_base[0] + k;
            const uint32_t qual16 = q ? (uint32_t)(s_base[2] + sm->mqual[k]) : 0xffffffffu;
            write_record_meta(O, rec_idx, s, r, (int)(m & kLen2), s_base[1] + sm->mseq[k], qual16);
            const uint32_t lf = __ldg(B.len_flag + r);
            const uint8_t kind = q ? 3 : (((sm->genbits[i >> 5] >> (i & 31)) & 1u) ? 2 : (sm->mpc[k] <= 2 ? 1 : 4));
            const int pos = __ldg(B.pos + r);
            if (kind == 1 && (lf & 0xffffu) <= 160u) {              // the common record: written here, nothing left for the emission kernels
                write_common_body(B, O, __ldg(B.seq_off16 + r), s_base[1] + sm->mseq[k], lf & 0xffffu, sm->mpatch[k], sm->mpc[k], pos - c.d.col_begin);
            } else if (kind == 1) {                                   // a long clean read: the copy kernel takes it
                E.edesc[rec_idx] = make_uint4(__ldg(B.seq_off16 + r), (uint32_t)(pos - c.d.col_begin), (lf & 0xffffu) | ((uint32_t)sm->mpc[k] << 16), sm->mpatch[k]);
                E.kind1_list[atomicAdd(E.n_kind1, 1u)] = (uint32_t)rec_idx;
            } else {
                const uint32_t slot = atomicAdd(E.n_special, 1u);
                uint32_t skind = kind;
                if (kind == 3) {                                      // more than two edits: emit_many_kernel writes the body (kind 5)
                    const int oa = sm->mhead[k];
                    const int ob = oa >= 0 ? (int)sm->o_rnext[oa] : -1;
                    if (ob >= 0 && sm->o_rnext[ob] >= 0) skind = 5;
                }
                if ((int64_t)slot < O.cap_records) {
                    write_special(E, B, c.d, slot, skind, __ldg(B.seq_off16 + r), pos, lf, s, r, m & kLen2, s_base[1] + sm->mseq[k], qual16);
                    if (skind == 5) { const uint32_t at = atomicAdd(E.n_many_recs, 1u); if (at < E.cap_many) E.many_recs[at] = slot; }
                }
            }
            if (kind == 3) {                                          // the edits travel in the record's (still unused) quality slot
                Ed2 E2; int nl = 0;
                const bool two = collect2(c, sm, k, (int)(lf & 0xffffu), E2, &nl);
                EditAux a;
                a.irp0 = E2.irp[0]; a.pos0 = E2.pos[0]; a.len0 = (uint32_t)E2.len[0] | ((E2.ne >= 1 && E2.n_del < 1) ? 0x80000000u : 0u);
                a.irp1 = E2.irp[1]; a.pos1 = E2.pos[1]; a.len1 = (uint32_t)E2.len[1] | ((E2.ne >= 2 && E2.n_del < 2) ? 0x80000000u : 0u);
                a.ne_ndel = (uint32_t)E2.ne | ((uint32_t)E2.n_del << 8);
                {   // where the read's quality record should sit in the sparse index
                    const int o = (int)sm->mhead[k];
                    const ObsRec* ob = X.obs + (size_t)(2 * (size_t)s) * kObsHalf;
                    const uint32_t qord = __ldg(&(o < n_obs0 ? ob + o : ob + kObsHalf + (o - n_obs0))->qord);
                    a.qidx = (uint32_t)(i < c.nt ? c.d.qt_begin : c.d.qn_begin) + qord;
                }
                uint4* dst = reinterpret_cast<uint4*>(O.out_qual + 32ull * qual16);
                const uint4* src = reinterpret_cast<const uint4*>(&a);
                if (two) { dst[0] = src[0]; dst[1] = src[1]; if (E.edit_keep) { E.edit_keep[2 * rec_idx] = src[0]; E.edit_keep[2 * rec_idx + 1] = src[1]; } }
                else write_many_aux(c, sm, k, (int)(lf & 0xffffu), a.qidx, dst);
            }
        }
        __syncthreads();                                              // tables and flags are reused by the next session
    }
}

}  // namespace ga

// ====================================================================================================================
// Lean variant: ONE WARP per session, no block barrier, 10.4 KB of shared memory per warp (20 warps per SM), and
// deliberately compact code (rolled loops): a warp runs the whole kernel body once per session, so the body has to
// fit the instruction cache (an unrolled version of 123 KB spent 70 % of its time waiting for instructions).
// Takes every session whose tables fit the small capacities below (the normal case); the others are listed in
// large_list for the CTA-per-session kernel above.
// Clean SNV-only records get their (at most two) germline hits written into the emission descriptor (kind 1), so
// the emission kernel copies and patches without touching the reference or the germline list.
namespace ga {

constexpr int kLeanWarps = 2;            // warps per CTA
constexpr int kReadsL = 1536;            // candidate reads per session
constexpr int kModL = 256;               // modified reads per session
constexpr int kObsL = 64;                // indel observations per session
constexpr int kEntL = 512;               // candidate entries per session
// "Mid" instantiation of the same kernel: the sessions of an indel-dense workload (hundreds of indel observations and
// modified reads per session) stay with the one-warp formulation instead of the barrier-bound one-CTA kernel; it walks
// the list the lean instantiation handed over and lists what it cannot hold for the one-CTA kernel.
constexpr int kReadsM = 4096;             // candidate reads per session in the mid instantiation (12 bits of an entry)
constexpr int kModM = 512;
constexpr int kObsM = 512;
constexpr int kEntM = 768;

template <int kReadsL, int kModL, int kObsL, int kEntL, int kChunks>
struct SmemLT {
    uint32_t tab[kCols2 / 4];            // one byte per column: bits 0-3 tumor saw A,C,G,T, bits 4-7 normal
    uint32_t modbits[kReadsL / 32], indelbits[kReadsL / 32], genbits[kReadsL / 32], woff[kReadsL / 32];
    uint32_t mpatch[kModL];              // two germline hits of a clean read: (column << 4) | reference code, 16 bits each
    int32_t mhead[kModL];                // per modified read: chain of its germline indel observations
    uint8_t mpc[kModL];                  // germline SNV hits per modified read
    uint32_t o_meta[kObsL], o_ra[kObsL]; int32_t o_irp[kObsL], o_col[kObsL], o_rnext[kObsL];
    uint32_t o_key[kObsL];               // hash of (column, type, length, allele): one compare rejects almost every pair
    uint32_t o_cls[kObsL];               // low 16 bits: the first observation with the same key; bits 16 / 17 (on that one): tumor / normal saw the key;
                                         // bit 18: a normal read that shows it covers its position
    union {                              // the allele signatures are dead once the observations are compared,
        struct { uint32_t o_s0[kObsL], o_s1[kObsL]; };
        uint16_t clist[kModL];           // ... which is before the list of modified reads is built
    };
    union {                              // the entries are dead after their third pass,
        uint32_t ent[kEntL];             // candidate entries (tumor item, then normal item); bit 31: germline hit
        uint32_t rnew[kModL];            // ... which is before the new lengths are known: new length | kind << 24
    };
    // what the warps of a team tell each other (a one-warp team keeps most of it in registers)
    unsigned long long base[3];
    uint32_t ngerm, flag, cnt_del, cnt_ins, tot_seq, tot_qual, n_q, n_spec, n_mod, ticket, spec_base, pad[3];
    uint32_t chunk_s[kChunks], chunk_q[kChunks], chunk_p[kChunks];   // per chunk of 32 modified reads: sequence units, quality units, special records
    static_assert(sizeof(uint16_t) * kModL <= 2 * sizeof(uint32_t) * kObsL && kModL <= kEntL, "aliases fit");
};
constexpr int kMidTeam = 4;              // warps that share one session in the mid instantiation
using SmemL = SmemLT<kReadsL, kModL, kObsL, kEntL, 4>;
using SmemM = SmemLT<kReadsM, kModM, kObsM, kEntM, kModM / 32>;
static_assert(sizeof(SmemL) % 16 == 0 && sizeof(SmemM) % 16 == 0, "per-warp slices stay 16-byte aligned");

// Alleles longer than the 16-base signature (rare): the bases behind it are compared from the records.
__device__ __noinline__ bool long_allele_tail_equal(const SessCtx& c, uint32_t ra_a, int irp_a, uint32_t ra_b, int irp_b) {
    const uint32_t* pa = rec_of(c, read_of(c, (int)(ra_a & 0xffffu)));
    const uint32_t* pb = rec_of(c, read_of(c, (int)(ra_b & 0xffffu)));
    const int na = (int)(ra_a >> 16);
    for (int j = 16; j < na; ++j)
        if (read_code(pa, irp_a + j) != read_code(pb, irp_b + j)) return false;
    return true;
}

// collect2 behind a call: the lean kernel needs it twice and must stay small enough for the instruction cache
template <class SM> __device__ __noinline__ bool lean_collect(int col_begin, const SM* sm, int k, int L, Ed2& E, int* new_len) { return collect2_at(col_begin, sm, k, L, E, new_len); }

// kFromList: the sessions are the entries of in_list (what the lean instantiation handed over) instead of 0 .. n_sessions - 1.
// kTeam: warps that work on one session together.  1 = every warp of the CTA has its own session and its own tables
// (no block barrier); kTeam = kWarps = the CTA is one team: loops are strided over the team, the phases are separated by
// block barriers and sums travel through shared memory - for sessions with hundreds of observations and modified reads,
// whose tables leave room for few sessions per SM: a team keeps four times as many warps (and memory requests) in
// flight per table.
template <int kReadsL, int kModL, int kObsL, int kEntL, int kWarps, int kMinBlocks, bool kFromList, int kTeam>
__global__ void __launch_bounds__(32 * kWarps, kMinBlocks) resolve_warp_kernel(BatchView B, SessView S, const SessionDesc* __restrict__ descs,
                                                                            int32_t* __restrict__ big_list, int32_t* __restrict__ n_big,
                                                                            const int32_t* __restrict__ in_list, const int32_t* __restrict__ n_in, unsigned int* __restrict__ ticket_p,
                                                                            int32_t* __restrict__ large_list, int32_t* __restrict__ n_large,
                                                                            ResultView O, ScanScratch X, EmitScratch2 E) {
    static_assert(kTeam == 1 || kTeam == kWarps, "a team is one warp or the whole CTA");
    constexpr bool kSolo = kTeam == 1;
    constexpr int TS = 32 * kTeam;                                     // threads of a team
    using SmemT = SmemLT<kReadsL, kModL, kObsL, kEntL, kSolo ? 4 : kModL / 32>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tw = kSolo ? 0 : warp;                                   // warp within the team
    const int tl = kSolo ? lane : (int)threadIdx.x;                    // lane within the team
    SmemT* sm = reinterpret_cast<SmemT*>(smem_raw) + (kSolo ? warp : 0);
    auto tsync = [] { if constexpr (kSolo) __syncwarp(); else __syncthreads(); };
    SessCtx c;
    c.B = B;
    c.totals = O.totals;
    memset(&c.T, 0, sizeof c.T);
    // statistics counters are summed per team and added to ga_totals once, at the end: every session already sends
    // three atomics to that one cache line for its output slots, and same-address atomics serialise in L2
    unsigned long long acc_reads = 0ull, acc_bases = 0ull;
    uint32_t acc_snv = 0u, acc_del = 0u, acc_ins = 0u, acc_q = 0u;

    // sessions differ in cost: a team takes the next one from a ticket counter (the ticket after that travels while the
    // session is processed), so that no team is left with a long tail of expensive sessions
    const uint32_t n_work = kFromList ? (uint32_t)*n_in : (uint32_t)S.n_sessions;
    uint32_t next_ticket = 0u;
    auto take = [&](uint32_t mine) -> uint32_t {                       // the team's thread 0 holds the ticket
        if constexpr (kSolo) return __shfl_sync(0xffffffffu, mine, 0);
        else {
            __syncthreads();                                          // everybody is done with the previous session's tables
            if (tl == 0) sm->ticket = mine;
            __syncthreads();
            return sm->ticket;
        }
    };
    uint32_t ticket = take(tl == 0 ? atomicAdd(ticket_p, 1u) : 0u);
#pragma unroll 1
    for (; ticket < n_work; ticket = take(next_ticket)) {
        const int s = kFromList ? in_list[ticket] : (int)ticket;
        next_ticket = tl == 0 ? atomicAdd(ticket_p, 1u) : 0u;
        // ---- round trip 1: descriptor, scan counts, variant_to_keep (every warp of a team for itself: same lines)
        uint32_t w1 = 0u, w2 = 0u;
        if (lane < 20) w1 = __ldg(reinterpret_cast<const uint32_t*>(descs + s) + lane);
        else if (lane < 28) w1 = __ldg(reinterpret_cast<const uint32_t*>(X.cnt + 2 * (size_t)s) + (lane - 20));
        if (lane == 0) w2 = (uint32_t)__ldg(S.keep_type + s); else if (lane == 1) w2 = (uint32_t)__ldg(S.keep_pos + s);
        else if (lane == 2) w2 = (uint32_t)__ldg(S.keep_end + s); else if (lane == 3) w2 = (uint32_t)__ldg(S.keep_len + s);
        else if (lane == 4) w2 = __ldg(S.keep_allele_off + s); else if (lane == 5) w2 = __ldg(S.keep_allele_off + s + 1);
        {
            uint32_t* dd = reinterpret_cast<uint32_t*>(&c.d);
#pragma unroll
            for (int k = 0; k < 20; ++k) dd[k] = __shfl_sync(0xffffffffu, w1, k);
        }
        if (c.d.big) continue;                                        // listed by the assignment kernel
        const uint32_t n_ent0 = __shfl_sync(0xffffffffu, w1, 20), n_obs0 = __shfl_sync(0xffffffffu, w1, 21);
        const uint32_t n_ent1 = __shfl_sync(0xffffffffu, w1, 24), n_obs1 = __shfl_sync(0xffffffffu, w1, 25);
        const uint32_t sess_reads = __shfl_sync(0xffffffffu, w1, 22) + __shfl_sync(0xffffffffu, w1, 26);
        const uint32_t sess_bases = __shfl_sync(0xffffffffu, w1, 23) + __shfl_sync(0xffffffffu, w1, 27);
        c.keep_type = (int)__shfl_sync(0xffffffffu, w2, 0); c.keep_pos = (int)__shfl_sync(0xffffffffu, w2, 1);
        c.keep_end = (int)__shfl_sync(0xffffffffu, w2, 2); c.keep_len = (int)__shfl_sync(0xffffffffu, w2, 3);
        const uint32_t ka0 = __shfl_sync(0xffffffffu, w2, 4), ka1 = __shfl_sync(0xffffffffu, w2, 5);
        c.keep_allele = S.keep_alleles + ka0;
        c.keep_alen = (int)(ka1 - ka0);
        c.s = s;
        c.nt = c.d.t_end - c.d.t_begin;
        c.n_range = c.nt + (c.d.n_end - c.d.n_begin);
        const int n_cols = c.d.n_cols;
        const int n_cw = (c.n_range + 31) >> 5;
        if (n_ent0 == kCntOverflow || n_ent1 == kCntOverflow) {       // the scan kernel could not hold the session
            if (tl == 0) { big_list[atomicAdd(n_big, 1)] = s; atomicAdd(n_big + 4 + 1, 1); }
            continue;
        }
        const int n_obs = (int)(n_obs0 + n_obs1);
        const uint32_t n_ent = n_ent0 + n_ent1;
        if (c.n_range > kReadsL || n_obs > kObsL) {
            if (tl == 0) large_list[atomicAdd(n_large, 1)] = s;
            continue;
        }
        // ---- round trip 2: entries, observations, keep allele; tables zeroed meanwhile
        // the entries are staged in shared memory when they fit (the usual session) and streamed from the scan kernel's
        // regions - three passes over L2 - when they do not (noisy reads: one candidate per mismatch)
        const uint32_t* e0 = X.ent + (size_t)(2 * (size_t)s) * kEntHalf;
        const uint32_t* e1 = e0 + kEntHalf;
        const bool ent_staged = n_ent <= (uint32_t)kEntL;
        auto ent_at = [&](uint32_t k) -> uint32_t { return ent_staged ? sm->ent[k] : (k < n_ent0 ? __ldg(e0 + k) : __ldg(e1 + (k - n_ent0))); };
        if (ent_staged) {
#pragma unroll 2
            for (uint32_t k = tl; k < n_ent; k += TS) sm->ent[k] = k < n_ent0 ? __ldg(e0 + k) : __ldg(e1 + (k - n_ent0));
        }
#pragma unroll 1
        for (int o = tl; o < n_obs; o += TS) {
            const ObsRec* op = X.obs + (size_t)(2 * (size_t)s) * kObsHalf;
            const uint4* src = reinterpret_cast<const uint4*>(o < (int)n_obs0 ? op + o : op + kObsHalf + (o - (int)n_obs0));
            const uint4 a = __ldg(src), b = __ldg(src + 1);
            sm->o_col[o] = (int)a.x; sm->o_meta[o] = a.y; sm->o_ra[o] = a.z; sm->o_irp[o] = (int)a.w; sm->o_s0[o] = b.x; sm->o_s1[o] = b.y;
            uint32_t kh = a.x * 0x9E3779B1u ^ ((a.y & (kMetaIns | kMetaLenMask)) * 0x85EBCA77u) ^ ((a.z >> 16) * 0xC2B2AE3Du) ^ (b.x * 0x27D4EB2Fu) ^ (b.y * 0x165667B1u);
            sm->o_key[o] = kh ^ (kh >> 15);
            sm->o_cls[o] = 0u;
        }
        {
            uint4* t4 = reinterpret_cast<uint4*>(sm->tab);
#pragma unroll 1
            for (int k = tl; k < ((n_cols + 15) >> 4); k += TS) t4[k] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
            for (int k = tl; k < n_cw; k += TS) { sm->modbits[k] = 0u; sm->indelbits[k] = 0u; sm->genbits[k] = 0u; }
            if (tl == 0) { sm->ngerm = 0u; sm->flag = 0u; sm->cnt_del = 0u; sm->cnt_ins = 0u; sm->tot_seq = 0u; sm->tot_qual = 0u; sm->n_q = 0u; sm->n_spec = 0u; }
        }
        uint32_t keep_key = 0xffffffffu;                              // variant_to_keep as an SNV entry key
        if (c.keep_type == GA_VT_SNV && c.keep_end == c.keep_pos && c.keep_len == 1 && c.keep_alen == 1) {
            const int kc = c.keep_pos - c.d.col_begin;
            if (kc >= 0 && kc < n_cols) {
                const uint8_t ch = c.keep_allele[0];
                const uint32_t code = ch == 'A' ? 1u : ch == 'C' ? 2u : ch == 'G' ? 4u : ch == 'T' ? 8u : 0u;
                if (code) keep_key = ((uint32_t)kc << 4) | code;
            }
        }
        tsync();
        // ---- pass 1: allele table
        bool bad = false;
#pragma unroll 1
        for (uint32_t k = tl; k < n_ent; k += TS) {
            const uint32_t e = ent_at(k), b = e & 15u;
            if (b == 0u || (b & (b - 1u))) { bad = true; continue; }   // IUPAC read base: the fallback kernel keeps all 16 codes
            const uint32_t col = (e >> 4) & 0xfffu;
            atomicOr(&sm->tab[col >> 2], 1u << ((__ffs(b) - 1) + (k >= n_ent0 ? 4 : 0) + 8 * (int)(col & 3u)));
        }
        if constexpr (kSolo) bad = __any_sync(0xffffffffu, bad);
        else { if (bad) sm->flag = 1u; __syncthreads(); bad = sm->flag != 0u; }
        if (bad) {
            if (tl == 0) { big_list[atomicAdd(n_big, 1)] = s; atomicAdd(n_big + 4 + 2, 1); }
            continue;
        }
        if constexpr (kSolo) __syncwarp();
        // ---- pass 2: germline = seen in tumor AND normal, minus variant_to_keep; mark the reads that carry one.
        // Bit 31 of an entry remembers that it is a germline hit (pass 3 needs only those).
#pragma unroll 1
        for (uint32_t k = tl; k < n_ent; k += TS) {
            const uint32_t e = ent_at(k);
            const uint32_t col = (e >> 4) & 0xfffu;
            const uint32_t byte = (sm->tab[col >> 2] >> (8 * (col & 3u))) & 0xffu;
            if ((((byte & (byte >> 4)) >> (__ffs(e & 15u) - 1)) & 1u) && (e & 0xffffu) != keep_key) {
                const uint32_t i = (e >> 16) & 0xfffu;
                atomicOr(&sm->modbits[i >> 5], 1u << (i & 31));
                if (e & kEntGen) atomicOr(&sm->genbits[i >> 5], 1u << (i & 31));
                if (ent_staged) sm->ent[k] = e | 0x80000000u;
            }
        }
        {   // distinct germline SNV alleles: the per-session counter (= the list's length) and the emission kernel's list
#pragma unroll 1
            for (int kw = tl; kw < ((n_cols + 3) >> 2); kw += TS) {
                const uint32_t w = sm->tab[kw];
                uint32_t g = w & (w >> 4) & 0x0f0f0f0fu;
#pragma unroll 1
                while (g) {
                    const int bit = __ffs(g) - 1; g &= g - 1;
                    const uint32_t key = ((uint32_t)(4 * kw + (bit >> 3)) << 4) | (1u << (bit & 7));
                    if (key == keep_key) continue;
                    const uint32_t slot = atomicAdd(&sm->ngerm, 1u);
                    if (slot < (uint32_t)kGermCap) E.germ[(size_t)s * kGermStride + 4 + slot] = key;
                }
            }
        }
        // indels: exact key equality (variants.py:83-96).  Every observation looks for the FIRST observation equal to it
        // (the representative of its key: equality is an equivalence, so that is the smallest member of the class) - one
        // hashed compare per candidate, a full compare only where the hash agrees - and tells the representative which
        // dataset saw the key; an observation is germline when its representative was told by both.
        uint32_t cnt_del = 0, cnt_ins = 0;
        // the smallest observation of every hash bucket (mpatch is not in use yet): equal keys share a bucket, so the
        // bucket's smallest member is the representative unless two different keys met in the bucket
        uint32_t* kt = sm->mpatch;
        static_assert((kModL & (kModL - 1)) == 0, "bucket count is a power of two");
        if (n_obs > 0) {
#pragma unroll 1
            for (int k = tl; k < kModL; k += TS) kt[k] = 0xffffffffu;
            tsync();
#pragma unroll 1
            for (int o = tl; o < n_obs; o += TS) atomicMin(&kt[sm->o_key[o] & (kModL - 1)], (uint32_t)o);
            tsync();
        }
#pragma unroll 1
        for (int o = tl; o < n_obs; o += TS) {
            const int o_col = sm->o_col[o];
            const uint32_t o_meta = sm->o_meta[o], o_ra = sm->o_ra[o], o_s0 = sm->o_s0[o], o_s1 = sm->o_s1[o], o_key = sm->o_key[o];
            int rep = o;
            const int cand = (int)kt[o_key & (kModL - 1)];
            auto same_key = [&](int j) -> bool {
                if (sm->o_key[j] != o_key || sm->o_col[j] != o_col) return false;
                const uint32_t j_meta = sm->o_meta[j];
                if (((j_meta ^ o_meta) & (kMetaIns | kMetaLenMask)) != 0u || (sm->o_ra[j] >> 16) != (o_ra >> 16) || sm->o_s0[j] != o_s0 || sm->o_s1[j] != o_s1) return false;
                return !((o_ra >> 16) > 16u && !long_allele_tail_equal(c, o_ra, sm->o_irp[o], sm->o_ra[j], sm->o_irp[j]));   // bases behind the 16-base signature
            };
            if (cand != o) {
                if (same_key(cand)) rep = cand;
                else {                                                    // two keys in one bucket (rare): look for the first equal one
#pragma unroll 1
                    for (int j = 0; j < o; ++j) if (same_key(j)) { rep = j; break; }
                }
            }
            atomicOr(&sm->o_cls[o], (uint32_t)rep);                   // the entry was zeroed; others may be adding their dataset bits to it
            atomicOr(&sm->o_cls[rep], (o_meta & kMetaDs) ? (0x20000u | ((o_meta & kMetaTrail) ? 0u : 0x40000u)) : 0x10000u);   // bit 18: the normal column of the key exists
        }
        tsync();
#pragma unroll 1
        for (int o = tl; o < n_obs; o += TS) {
            const uint32_t cls = sm->o_cls[o];
            const int rep = (int)(cls & 0xffffu);
            const uint32_t seen = sm->o_cls[rep] >> 16;
            bool germ = (seen & 3u) == 3u;
            const uint32_t o_meta = sm->o_meta[o], o_ra = sm->o_ra[o];
            if (germ && !(seen & 4u)) germ = normal_covers(c, __ldg(S.first + s), sm->o_col[o] + c.d.col_begin);   // only insertions that end their reads: anonymizer_methods.py:474-481
            if (germ) {                                               // variant_to_keep may be this indel
                const int o_col = sm->o_col[o];
                const int type = (o_meta & kMetaIns) ? GA_VT_INS : GA_VT_DEL;
                const int len = (int)(o_meta & kMetaLenMask), pos = o_col + c.d.col_begin;
                const int end = (type == GA_VT_INS) ? pos + 1 : pos + len - 1;       // variation_classifier.py:86
                const int na = (int)(o_ra >> 16);
                if (c.keep_type == type && c.keep_pos == pos && c.keep_len == len && c.keep_end == end && c.keep_alen == na) {
                    const char* code2asc = "=ACMGRSVTWYHKDBN";
                    const uint32_t o_s0 = sm->o_s0[o], o_s1 = sm->o_s1[o];
                    bool same = true;
#pragma unroll 1
                    for (int j = 0; j < na; ++j) {
                        const uint32_t code = j < 16 ? (((j < 8 ? o_s0 : o_s1) >> (4 * (j & 7))) & 15u)
                                                     : read_code(rec_of(c, read_of(c, (int)(o_ra & 0xffffu))), sm->o_irp[o] + j);
                        if (c.keep_allele[j] != (uint8_t)code2asc[code]) same = false;
                    }
                    if (same) germ = false;
                }
            }
            if (germ) {
                const uint32_t i = o_ra & 0xffffu;
                atomicOr(&sm->modbits[i >> 5], 1u << (i & 31));
                atomicOr(&sm->indelbits[i >> 5], 1u << (i & 31));
                sm->o_meta[o] = o_meta | kMetaGerm;
                if (rep == o) { if (o_meta & kMetaIns) ++cnt_ins; else ++cnt_del; }
            }
        }
        if (n_obs > 0) {
            cnt_del = warp_sum(cnt_del); cnt_ins = warp_sum(cnt_ins);
            if constexpr (!kSolo) { if (lane == 0) { if (cnt_del) atomicAdd(&sm->cnt_del, cnt_del); if (cnt_ins) atomicAdd(&sm->cnt_ins, cnt_ins); } }
        }
        tsync();
        if constexpr (!kSolo) { cnt_del = sm->cnt_del; cnt_ins = sm->cnt_ins; }
        const uint32_t ngerm = sm->ngerm, cnt_snv = ngerm;
        // ---- ordered list of the modified reads (the bitmap words are scanned by the team's first warp)
        uint32_t n_mod = 0u;
        if (tw == 0) {
            constexpr int kWPL = (kReadsL / 32 + 31) / 32;                // bitmap words per lane: word lane + 32 j
            uint32_t bw[kWPL], offw[kWPL];
#pragma unroll
            for (int j = 0; j < kWPL; ++j) {
                const int idx = lane + 32 * j;
                bw[j] = idx < n_cw ? sm->modbits[idx] : 0u;
                uint32_t t;
                offw[j] = n_mod + warp_excl_scan(__popc(bw[j]), lane, &t);
                n_mod += t;
                if (idx < n_cw) sm->woff[idx] = offw[j];
            }
            if (n_mod <= (uint32_t)kModL) {
#pragma unroll
                for (int j = 0; j < kWPL; ++j) {
                    uint32_t b = bw[j], o = offw[j];
#pragma unroll 1
                    while (b) { const int k = __ffs(b) - 1; b &= b - 1; sm->clist[o++] = (uint16_t)((lane + 32 * j) * 32 + k); }
                }
            }
            if constexpr (!kSolo) { if (lane == 0) sm->n_mod = n_mod; }
        }
        if constexpr (!kSolo) { __syncthreads(); n_mod = sm->n_mod; }
        if (n_mod > (uint32_t)kModL) {
            if (tl == 0) large_list[atomicAdd(n_large, 1)] = s;
            continue;
        }
        if (ngerm > (uint32_t)kGermCap) {
            if (tl == 0) { big_list[atomicAdd(n_big, 1)] = s; atomicAdd(n_big + 4 + 3, 1); }
            continue;
        }
        {
            uint32_t* pc4 = reinterpret_cast<uint32_t*>(sm->mpc);
#pragma unroll 1
            for (int k = tl; k < (int)((n_mod + 3) >> 2); k += TS) pc4[k] = 0u;
            if (n_obs > 0) for (int k = tl; k < (int)n_mod; k += TS) sm->mhead[k] = -1;
        }
        tsync();
#pragma unroll 1
        for (int o = tl; o < n_obs; o += TS) {                      // hang every germline observation on its modified read
            if (!(sm->o_meta[o] & kMetaGerm)) continue;
            const uint32_t i = sm->o_ra[o] & 0xffffu;
            const uint32_t k = sm->woff[i >> 5] + __popc(sm->modbits[i >> 5] & ((1u << (i & 31)) - 1u));
            sm->o_rnext[o] = atomicExch(&sm->mhead[k], o);
        }
        // ---- pass 3: the germline hits of every clean modified read
#pragma unroll 1
        for (uint32_t k = tl; k < n_ent; k += TS) {
            const uint32_t e = ent_at(k);
            if (ent_staged) { if ((e & (0x80000000u | kEntGen)) != 0x80000000u) continue; }
            else {                                                    // streamed: the germline test of pass 2 again
                if (e & kEntGen) continue;
                const uint32_t c2 = (e >> 4) & 0xfffu, byte = (sm->tab[c2 >> 2] >> (8 * (c2 & 3u))) & 0xffu;
                if (!((((byte & (byte >> 4)) >> (__ffs(e & 15u) - 1)) & 1u) && (e & 0xffffu) != keep_key)) continue;
            }
            const uint32_t col = (e >> 4) & 0xfffu, i = (e >> 16) & 0xfffu;
            const uint32_t m = sm->woff[i >> 5] + __popc(sm->modbits[i >> 5] & ((1u << (i & 31)) - 1u));
            uint32_t* cw = reinterpret_cast<uint32_t*>(sm->mpc) + (m >> 2);
            const uint32_t have = (atomicAdd(cw, 1u << (8 * (m & 3u))) >> (8 * (m & 3u))) & 0xffu;   // at most kGermCap hits per read: no carry
            if (have < 2u) reinterpret_cast<uint16_t*>(sm->mpatch)[2 * m + have] = (uint16_t)((col << 4) | (1u << ((e >> 29) & 3u)));
        }
        tsync();
        // ---- round trip 3: new length of every modified read (indel-masked reads need the edit analysis), output sizes.
        // A warp takes chunks of 32 consecutive modified reads; a team notes every chunk's sums for the header loop.
        uint32_t tot_seq = 0, tot_qual = 0, n_q = 0, n_spec = 0;
        bool slow = false;
#pragma unroll 1
        for (uint32_t kb = 32u * tw; kb < n_mod; kb += TS) {
            const uint32_t k = kb + lane;
            uint32_t units = 0u, qunits = 0u, spec = 0u;
            if (k < n_mod) {
                const int i = (int)sm->clist[k];
                const int64_t r3 = read_of(c, i);
                const uint32_t so3 = __ldg(B.seq_off16 + r3);         // travels with the length; only used to ask the record
                const int L0 = (int)(__ldg(B.len_flag + r3) & 0xffffu);
                // the header loop below reads pos and the record body of this read: ask them into L2 now
                asm volatile("prefetch.global.L2 [%0];" ::"l"(B.pos + r3));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(B.seq4 + 16ull * so3));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(B.seq4 + 16ull * so3 + 64));
                int new_len = L0;
                uint32_t kind;
                if ((sm->indelbits[i >> 5] >> (i & 31)) & 1u) {
                    Ed2 E2;
                    kind = 3u;
                    if (!lean_collect(c.d.col_begin, sm, (int)k, L0, E2, &new_len)) {  // more than two edits: the edit list travels in the side buffer (kind 5)
                        kind = 5u;
                        if (!reserve_many(E, c, sm, (int)k, L0, &new_len)) slow = true;   // no room, or more edits than the pipeline takes
                    }
                    ++n_q;
                } else {
                    kind = ((sm->genbits[i >> 5] >> (i & 31)) & 1u) ? 2u : (sm->mpc[k] <= 2 ? 1u : 4u);
                }
                if (kind != 1u) spec = 1u;
                units = ((uint32_t)new_len + 31u) / 32u; if (units < 1u) units = 1u;
                if (kind == 3u || kind == 5u) qunits = units;
                sm->rnew[k] = ((uint32_t)new_len & kLen2) | (kind << 24);
            }
            if constexpr (kSolo) { tot_seq += units; tot_qual += qunits; n_spec += spec; }
            else {
                const uint32_t cs = warp_sum(units), cq = warp_sum(qunits), cp = warp_sum(spec);
                if (lane == 0) { sm->chunk_s[kb >> 5] = cs; sm->chunk_q[kb >> 5] = cq; sm->chunk_p[kb >> 5] = cp; }
                tot_seq += cs; tot_qual += cq; n_spec += cp;
            }
        }
        if constexpr (kSolo) {
            slow = __any_sync(0xffffffffu, slow);
            tot_seq = warp_sum(tot_seq); tot_qual = warp_sum(tot_qual); n_q = warp_sum(n_q); n_spec = warp_sum(n_spec);
        } else {
            n_q = warp_sum(n_q);
            if (slow) sm->flag = 1u;
            if (lane == 0) { atomicAdd(&sm->tot_seq, tot_seq); atomicAdd(&sm->tot_qual, tot_qual); atomicAdd(&sm->n_q, n_q); atomicAdd(&sm->n_spec, n_spec); }
            __syncthreads();
            slow = sm->flag != 0u; tot_seq = sm->tot_seq; tot_qual = sm->tot_qual; n_q = sm->n_q; n_spec = sm->n_spec;
        }
        if (slow) {                                                   // nothing is reserved yet: hand the session over
            if (tl == 0) large_list[atomicAdd(n_large, 1)] = s;
            continue;
        }
        // ---- output slots: one atomicAdd per cursor per session (north_star job (4): compaction)
        unsigned long long base = 0ull;
        uint32_t spec_base = 0u;
        if (tw == 0) {
            if (lane == 0) base = atomicAdd((unsigned long long*)&O.totals->n_modified, (unsigned long long)n_mod);
            else if (lane == 1) base = atomicAdd((unsigned long long*)&O.totals->seq16_used, (unsigned long long)tot_seq);
            else if (lane == 2) base = atomicAdd((unsigned long long*)&O.totals->qual16_used, (unsigned long long)tot_qual);
            else if (lane >= 5 && lane < 9) O.sess_counts[4 * (size_t)s + (lane - 5)] = lane == 5 ? cnt_snv : lane == 6 ? cnt_del : lane == 7 ? cnt_ins : sess_reads;
            else if (lane == 9) { E.germ[(size_t)s * kGermStride] = ngerm; E.germ[(size_t)s * kGermStride + 1] = (uint32_t)c.d.col_begin; }
            else if (lane == 11) { if (n_spec) spec_base = atomicAdd(E.n_special, n_spec); }
            acc_reads += sess_reads; acc_bases += sess_bases; acc_snv += cnt_snv; acc_del += cnt_del; acc_ins += cnt_ins; acc_q += n_q;
            if constexpr (!kSolo) { if (lane < 3) sm->base[lane] = base; else if (lane == 11) sm->spec_base = spec_base; }
        }
        unsigned long long base_rec, base_seq, base_qual;
        if constexpr (kSolo) {
            spec_base = __shfl_sync(0xffffffffu, spec_base, 11);
            base_rec = __shfl_sync(0xffffffffu, base, 0); base_seq = __shfl_sync(0xffffffffu, base, 1); base_qual = __shfl_sync(0xffffffffu, base, 2);
        } else {
            __syncthreads();
            spec_base = sm->spec_base; base_rec = sm->base[0]; base_seq = sm->base[1]; base_qual = sm->base[2];
        }
        const bool fits = (int64_t)(base_rec + n_mod) <= O.cap_records && (int64_t)(base_seq + tot_seq) <= O.cap_seq16 &&
                          (int64_t)(base_qual + tot_qual) <= O.cap_qual16;
        if (!fits) {                                                  // the reserved special slots must not stay undefined
            for (uint32_t t = tl; t < n_spec; t += TS) if ((int64_t)(spec_base + t) < O.cap_records) E.sdesc[4ull * (spec_base + t)] = make_uint4(0u, 0u, 0u, 0u);
            if (tl == 0) raise_error(O.totals, GA_ERR_CAPACITY, 0xffffffffu);
            continue;
        }
        // ---- record headers and the hand-over to the emission kernels
        uint32_t run_seq = 0u, run_qual = 0u;
#pragma unroll 1
        for (uint32_t kb = 32u * tw; kb < n_mod; kb += TS) {
            if constexpr (!kSolo) {                                    // where the chunk starts: the sums of the chunks in front of it
                const uint32_t ch = kb >> 5;
                const bool in_front = (uint32_t)lane < ch;
                run_seq = warp_sum(in_front ? sm->chunk_s[lane] : 0u); run_qual = warp_sum(in_front ? sm->chunk_q[lane] : 0u);
                spec_base = sm->spec_base + warp_sum(in_front ? sm->chunk_p[lane] : 0u);
            }
            const uint32_t k = kb + lane;
            const bool have_k = k < n_mod;
            const uint32_t rn = have_k ? sm->rnew[k] : 0u;
            const uint32_t kind = rn >> 24, new_len = rn & kLen2;
            const bool has_qual = kind == 3u || kind == 5u;
            const int i = have_k ? (int)sm->clist[k] : 0;
            const int64_t r = read_of(c, i);
            uint32_t lf = 0u, so = 0u; int pos = 0;
            if (have_k) { lf = __ldg(B.len_flag + r); so = __ldg(B.seq_off16 + r); pos = __ldg(B.pos + r); }
            uint32_t units = have_k ? (new_len + 31u) / 32u : 0u; if (have_k && units < 1u) units = 1u;
            uint32_t ts, tq;
            const uint32_t so_rel = run_seq + warp_excl_scan(units, lane, &ts);
            const uint32_t qo_rel = run_qual + warp_excl_scan(has_qual ? units : 0u, lane, &tq);
            run_seq += ts; run_qual += tq;
            const bool is_spec = have_k && kind != 1u;
            const uint32_t sb = __ballot_sync(0xffffffffu, is_spec);
            const uint32_t my_slot = spec_base + __popc(sb & ((1u << lane) - 1u));
            spec_base += __popc(sb);
            if (!have_k) continue;
            const uint64_t rec_idx = base_rec + k;
            const uint32_t qual16 = has_qual ? (uint32_t)(base_qual + qo_rel) : 0xffffffffu;
            write_record_meta(O, rec_idx, s, r, (int)new_len, base_seq + so_rel, qual16);
            const uint32_t L0 = lf & 0xffffu;
            if (kind == 1u && L0 <= 160u) {
                // the common record: the lane writes the body itself; the emission kernel skips kind 0
                write_common_body(B, O, so, base_seq + so_rel, L0, sm->mpatch[k], sm->mpc[k], pos - c.d.col_begin);
                continue;                                             // nothing left for the emission kernels
            }
            if (kind == 1u) {                                         // a long clean read: the copy kernel takes it
                E.edesc[rec_idx] = make_uint4(so, (uint32_t)(pos - c.d.col_begin), L0 | ((uint32_t)sm->mpc[k] << 16), sm->mpatch[k]);
                E.kind1_list[atomicAdd(E.n_kind1, 1u)] = (uint32_t)rec_idx;
            } else if ((int64_t)my_slot < O.cap_records) {
                write_special(E, B, c.d, my_slot, kind, so, pos, lf, s, r, new_len, base_seq + so_rel, qual16);
                if (kind == 5u) { const uint32_t at = atomicAdd(E.n_many_recs, 1u); if (at < E.cap_many) E.many_recs[at] = my_slot; }
            }
            if (has_qual) {                                           // the edits travel in the record's (still unused) quality slot
                uint4* dst = reinterpret_cast<uint4*>(O.out_qual + 32ull * qual16);
                const int o = sm->mhead[k];                           // any germline observation of the read knows its ordinal
                const ObsRec* ob = X.obs + (size_t)(2 * (size_t)s) * kObsHalf;
                const uint32_t qord = __ldg(&(o < (int)n_obs0 ? ob + o : ob + kObsHalf + (o - (int)n_obs0))->qord);
                const uint32_t qidx = (uint32_t)(i < c.nt ? c.d.qt_begin : c.d.qn_begin) + qord;
                if (kind == 3u) {
                    Ed2 E2; int nl = 0;
                    lean_collect(c.d.col_begin, sm, (int)k, (int)L0, E2, &nl);
                    dst[0] = make_uint4((uint32_t)E2.irp[0], (uint32_t)E2.pos[0], (uint32_t)E2.len[0] | ((E2.ne >= 1 && E2.n_del < 1) ? 0x80000000u : 0u), (uint32_t)E2.irp[1]);
                    dst[1] = make_uint4((uint32_t)E2.pos[1], (uint32_t)E2.len[1] | ((E2.ne >= 2 && E2.n_del < 2) ? 0x80000000u : 0u),
                                        (uint32_t)E2.ne | ((uint32_t)E2.n_del << 8), qidx);
                    if (E.edit_keep) { E.edit_keep[2 * rec_idx] = dst[0]; E.edit_keep[2 * rec_idx + 1] = dst[1]; }
                } else write_many_aux(c, sm, (int)k, (int)L0, qidx, dst);
            }
        }
    }
    if (tw == 0) {
        if (lane == 3 && acc_reads) atomicAdd((unsigned long long*)&O.totals->session_reads, acc_reads);
        else if (lane == 4 && acc_bases) atomicAdd((unsigned long long*)&O.totals->session_bases, acc_bases);
        else if (lane == 5 && acc_snv) atomicAdd((unsigned long long*)&O.totals->masked[0], (unsigned long long)acc_snv);
        else if (lane == 6 && acc_del) atomicAdd((unsigned long long*)&O.totals->masked[1], (unsigned long long)acc_del);
        else if (lane == 7 && acc_ins) atomicAdd((unsigned long long*)&O.totals->masked[2], (unsigned long long)acc_ins);
        else if (lane == 10 && acc_q) atomicAdd((unsigned long long*)&O.totals->indel_records, (unsigned long long)acc_q);
    }
}

}  // namespace ga
