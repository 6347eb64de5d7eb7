// ga_session_v2.cuh - the production session kernel: persistent CTAs, one session (= one anonymize() call of
// the reference, anonymizer_methods.py:431-535) at a time per CTA, working set in shared memory.
//
// Pipeline of one session (block-synchronous phases, 256 threads):
//   A1 scan      thread per read.  Clean reads (one M/=/X op over the whole read - nearly all of them) are
//                fetched with 128-bit loads and compared with the 4-bit reference 32 bases per unit; only the
//                indices of mismatching 8-base words are queued.  Reads with any other CIGAR are queued too.
//   A2 discover  thread per queued word / queued read, lanes dense: SNV alleles are OR-ed into the
//                per-column table and appended to an entry list, indel observations chained per column
//                (variation_classifier.py:52-182).
//   R  resolve   germline = seen in tumor AND normal, minus variant_to_keep (anonymizer_methods.py:546-547)
//   M  mark      entries / observations whose allele is germline mark their read as modified
//   L  list      ordered list of the modified reads, B1 their new lengths, one scan for the output slots
//   B2 emit      one warp per record, coalesced word-per-lane copies; SNV masking of clean reads is applied
//                afterwards from the entry list (one 4-bit XOR per hit), other reads are re-walked by their
//                warp; indel-masked reads go through a shared-memory staging and the backward edit map.
// The first kernel (ga_session_kernel.cuh) remains as the global-scratch fallback for oversize sessions.
#pragma once
#include "ga_session_kernel.cuh"

namespace ga {

constexpr int kCols2 = 2688;           // allele-table columns per session
constexpr int kReads2 = 4096;          // candidate reads per session
constexpr int kObs2 = 512;             // indel observations per session
constexpr int kEnt2 = 2048;            // SNV candidate entries per session
constexpr int kMod2 = 512;             // modified reads per session
constexpr int kWords2 = 768;           // queued mismatching words per session
constexpr int kGen2 = 1024;            // queued non-clean reads per session
constexpr int kStageWords = 64;        // reads longer than 8*kStageWords bases send their session to the fallback kernel
constexpr uint32_t kLen2 = (1u << 24) - 1;   // msize: length bits (flags above: kModFlag, kQualFlag, kSlowFlag)
constexpr uint32_t kSlowFlag = 1u << 29;   // msize: indel-masked record with more than two edits (emitted in-kernel)
constexpr int kGermCap = 32;           // germline SNV alleles per session handed to the emission kernel (more: emitted in-kernel)
constexpr int kGroup = 8;              // lanes that cooperate on one non-trivial output record

// Hand-over from the session kernel to the emission kernel (engine scratch, one per lane).
// kind[k] of output record k: 0 = already written (or nothing to do), 1 = clean read, SNV-only: copy + patch,
// 2 = other CIGAR, SNV-only: re-walk, 3 = indel-masked with <= 2 edits (their description sits in the first
// 32 bytes of the record's out_qual slot until the emission kernel overwrites it with the qualities).
struct EmitScratch {
    uint8_t* kind;          // [cap_records]
    uint32_t* germ;         // [n_sessions][kGermCap]  (column << 4) | base code, column relative to the session's col_begin
    uint32_t* germ_n;       // [n_sessions]
};
struct EditAux { int32_t irp0, pos0; uint32_t len0; int32_t irp1, pos1; uint32_t len1; uint32_t ne, n_del; };   // len bit 31 = INS
static_assert(sizeof(EditAux) == 32, "EditAux fills one 32-byte quality unit");

struct Smem2 {
    uint32_t snv[kCols2];              // bit c: tumor saw base code c, bit 16+c: normal; after resolve: germline codes
    int16_t ihead[kCols2];             // head of the indel-observation chain of the column (-1: none)
    uint32_t ent[kEnt2];               // (session-relative read << 16) | (column << 4) | base code
    uint32_t wlist[kWords2];           // (session-relative read << 5) | word index; after phase A2 reused as clist
    uint32_t lists[3 * kMod2];         // msize | mseq | mqual
    uint32_t sref[kCols2 / 8 + 8];     // 4-bit reference of the session's columns (word 0 = ref4 word of column col_begin)
    uint16_t glist[kGen2];             // queued non-clean reads (phase A); after A2: indices of the records that need a lane group
    uint32_t modbits[kReads2 / 32];
    uint32_t indelbits[kReads2 / 32];
    uint32_t genbits[kReads2 / 32];
    uint32_t woff[kReads2 / 32];       // modified reads before bitmap word w
    int32_t o_col[kObs2];              // indel observations: column,
    uint32_t o_meta[kObs2];            //   type / dataset / germline flags / op length,
    uint32_t o_read[kObs2];            //   session-relative read (low 16) | allele length (high 16),
    int32_t o_irp[kObs2];              //   in_read_pos with the reference's H/N quirk,
    uint32_t o_sig0[kObs2];            //   first 16 allele bases, 4 bits each,
    uint32_t o_sig1[kObs2];
    int16_t o_next[kObs2];             //   next observation at the same column,
    int16_t o_rnext[kObs2];            //   next germline observation of the same modified read
    int32_t mhead[kMod2];              // per modified read: chain of its germline indel observations
};
static_assert(kWords2 >= kMod2, "clist aliases the word queue");

__device__ __forceinline__ uint4 ldg128(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Bulk prefetch of a byte range into L2 (TMA engine, no shared memory or registers tied up).
__device__ __forceinline__ void prefetch_l2(const void* p, uint64_t bytes) {
    uint64_t a = reinterpret_cast<uint64_t>(p) & ~15ull;
    bytes = (bytes + (reinterpret_cast<uint64_t>(p) - a) + 15ull) & ~15ull;
    while (bytes) {
        const uint32_t n = bytes > (1u << 20) ? (1u << 20) : (uint32_t)bytes;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(n) : "memory");
        a += n; bytes -= n;
    }
}

// Everything phase A1 of session `d` will read, requested while the previous session is still running.
__device__ __forceinline__ void prefetch_session(const BatchView& B, const SessionDesc& d, bool records) {
    if (d.big) return;
    if (records) {
        prefetch_l2(B.seq4 + 16ull * d.t_seq_lo, 16ull * d.t_seq_n);
        prefetch_l2(B.seq4 + 16ull * d.n_seq_lo, 16ull * d.n_seq_n);
    }
    const int64_t nt = d.t_end - d.t_begin, nn = d.n_end - d.n_begin;
    prefetch_l2(B.pos + d.t_begin, 4 * nt);       prefetch_l2(B.pos + d.n_begin, 4 * nn);
    prefetch_l2(B.len_flag + d.t_begin, 4 * nt);  prefetch_l2(B.len_flag + d.n_begin, 4 * nn);
    prefetch_l2(B.seq_off16 + d.t_begin, 4 * nt); prefetch_l2(B.seq_off16 + d.n_begin, 4 * nn);
    prefetch_l2(B.cigar_off + d.t_begin, 4 * nt + 4); prefetch_l2(B.cigar_off + d.n_begin, 4 * nn + 4);
    prefetch_l2(B.cigar + d.t_cig_lo, 4ull * d.t_cig_n); prefetch_l2(B.cigar + d.n_cig_lo, 4ull * d.n_cig_n);
}

__device__ __forceinline__ uint32_t tail_mask(int L, int word) {           // valid nibbles of query word `word`
    const int nv = L - word * 8;
    return nv >= 8 ? 0xffffffffu : (nv <= 0 ? 0u : (0xffffffffu >> ((8 - nv) * 4)));
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Bitmask of the 8-base words of a clean read that differ from the reference (U 16-byte units).  The record
// comes in with U independent 128-bit loads; the reference words come from the session's window in shared memory.
template <int U>
__device__ __forceinline__ uint32_t clean_word_mask(const uint4* __restrict__ rec, const uint32_t* __restrict__ sref, int rel_nibble, int L) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ldg128(rec + u);
    const uint32_t* rp = sref + (rel_nibble >> 3);
    const uint32_t sh = (uint32_t)(rel_nibble & 7) * 4u;
    uint32_t prev = rp[0];
    uint32_t wm = 0u;
#pragma unroll
    for (int k = 0; k < 4 * U; ++k) {
        const uint32_t next = rp[k + 1];
        const uint32_t fw = __funnelshift_r(prev, next, sh);             // sh == 0 returns prev
        prev = next;
        const uint32_t rw = (k & 3) == 0 ? v[k >> 2].x : (k & 3) == 1 ? v[k >> 2].y : (k & 3) == 2 ? v[k >> 2].z : v[k >> 2].w;
        if (k >= 4 * (U - 1)) {                                           // only the last unit can hold padding
            if ((rw ^ fw) & tail_mask(L, k)) wm |= 1u << k;
        } else if (rw != fw) wm |= 1u << k;
    }
    return wm;
}

struct Queues { uint32_t* n_words; uint32_t* n_gen; uint32_t* n_ent; uint32_t* n_obs; uint32_t* overflow; };

__device__ __forceinline__ void push_entry(Smem2* sm, const Queues& Q, int i, int col, uint32_t b, uint32_t ds) {
    atomicOr(&sm->snv[col], 1u << (b + 16u * ds));
    const uint32_t e = atomicAdd(Q.n_ent, 1u);
    if (e < (uint32_t)kEnt2) sm->ent[e] = ((uint32_t)i << 16) | ((uint32_t)col << 4) | b;
    else *Q.overflow = 1u;
}

// ------------------------------------------------------------------ phase A1: one read
struct ReadMeta { int pos; uint32_t lf, so, c0, c1; };

__device__ __forceinline__ ReadMeta load_meta(const SessCtx& c, int i) {
    const int64_t r = read_of(c, i);
    ReadMeta m;
    m.pos = __ldg(c.B.pos + r); m.lf = __ldg(c.B.len_flag + r); m.so = __ldg(c.B.seq_off16 + r);
    m.c0 = __ldg(c.B.cigar_off + r); m.c1 = __ldg(c.B.cigar_off + r + 1);
    return m;
}

// The record and reference loads are issued together with the load of the first CIGAR word (they only need
// the prefetched meta), so a read costs two dependent memory round trips instead of three.
__device__ __forceinline__ void scan_read(const SessCtx& c, Smem2* sm, const Queues& Q, int i, const ReadMeta& m, uint32_t& n_reads, uint32_t& n_bases) {
    const int pos = m.pos;
    const int L = (int)(m.lf & 0xffffu);
    const int units = (L + 31) >> 5;
    const bool one_op = (m.c1 - m.c0 == 1u);
    // speculative: a single-op read whose span L stays inside the reference and the session table
    const bool spec = one_op && pos >= 0 && (int64_t)pos + L <= c.B.ref_len && pos >= c.d.col_begin && pos + L - c.d.col_begin < c.d.n_cols;
    const uint32_t w0 = m.c1 > m.c0 ? __ldg(c.B.cigar + m.c0) : 0u;
    const uint4* rec = reinterpret_cast<const uint4*>(c.B.seq4 + 16ull * m.so);
    uint32_t wm = 0u;
    bool supported = spec;
    // nibble offset of base `pos` inside the staged window (ref4 nibble of base p is p + 8; the window starts at
    // the ref4 word holding column col_begin)
    const int rel = pos + 8 - ((c.d.col_begin + 8) & ~7);
    if (spec) {
        switch (units) {
            case 5: wm = clean_word_mask<5>(rec, sm->sref, rel, L); break;
            case 4: wm = clean_word_mask<4>(rec, sm->sref, rel, L); break;
            case 3: wm = clean_word_mask<3>(rec, sm->sref, rel, L); break;
            case 8: wm = clean_word_mask<8>(rec, sm->sref, rel, L); break;
            default: supported = false; break;
        }
    }
    const uint32_t op0 = w0 & 15u;
    const bool clean = supported && (op0 == 0u || op0 == 7u || op0 == 8u) && ((int)(w0 >> 4) == L);
    if (!clean) {                                                        // any other CIGAR (or an error case): queued for phase A2
        const uint32_t g = atomicAdd(Q.n_gen, 1u);
        if (g < (uint32_t)kGen2) sm->glist[g] = (uint16_t)i; else *Q.overflow = 1u;
        atomicOr(&sm->genbits[i >> 5], 1u << (i & 31));
        return;
    }
    if (pos + L <= c.first) return;                                      // fetched by range, does not reach the region
    n_reads += 1u; n_bases += (uint32_t)L;
    while (wm) {
        const int k = __ffs(wm) - 1; wm &= wm - 1;
        const uint32_t e = atomicAdd(Q.n_words, 1u);
        if (e < (uint32_t)kWords2) sm->wlist[e] = ((uint32_t)i << 5) | (uint32_t)k; else *Q.overflow = 1u;
    }
}

// One tile of 32 consecutive reads of one dataset whose records are contiguous and equally long (the normal case):
// the warp fetches the tile's records with U fully coalesced 128-bit loads per lane - every 32-byte sector is
// requested exactly once - and each lane compares the 16-byte units it happens to hold (32 bases of some read
// of the tile) with that read's reference bases, whose position comes from the owning lane by shuffle.
// The record loads are issued before anything that depends on the CIGAR word, so a tile costs one memory
// round trip (its meta was fetched while the previous tile was being compared).
template <int U>
__device__ __forceinline__ void scan_tile(const SessCtx& c, Smem2* sm, const Queues& Q, int i0, int n_valid, int lane, const ReadMeta& m,
                                          uint32_t so0, uint32_t& n_reads, uint32_t& n_bases) {
    const uint4* base = reinterpret_cast<const uint4*>(c.B.seq4 + 16ull * so0);
    const int n_chunks = n_valid * U;
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int g = j * 32 + lane;
        v[j] = g < n_chunks ? ldg128(base + g) : make_uint4(0u, 0u, 0u, 0u);
    }
    const int i = i0 + lane;
    const bool valid = lane < n_valid;
    const int pos = m.pos;
    const int L = (int)(m.lf & 0xffffu);
    const bool one_op = valid && (m.c1 - m.c0 == 1u);
    const bool spec = one_op && pos >= 0 && (int64_t)pos + L <= c.B.ref_len && pos >= c.d.col_begin && pos + L - c.d.col_begin < c.d.n_cols;
    const uint32_t w0 = one_op ? __ldg(c.B.cigar + m.c0) : 0u;
    const uint32_t op0 = w0 & 15u;
    const bool clean = spec && (op0 == 0u || op0 == 7u || op0 == 8u) && ((int)(w0 >> 4) == L);
    const int rel = pos + 8 - ((c.d.col_begin + 8) & ~7);              // nibble offset of base `pos` inside the staged window
    const bool in_session = clean && pos + L > c.first;                  // fetched by range but not reaching the region: skipped
    const uint32_t clean_mask = __ballot_sync(0xffffffffu, in_session);
    if (in_session) { n_reads += 1u; n_bases += (uint32_t)L; }
    if (valid && !clean) {                                               // any other CIGAR (or an error case): queued for phase A2
        const uint32_t g = atomicAdd(Q.n_gen, 1u);
        if (g < (uint32_t)kGen2) sm->glist[g] = (uint16_t)i; else *Q.overflow = 1u;
        atomicOr(&sm->genbits[i >> 5], 1u << (i & 31));
    }
    uint32_t all = 0u;                                                   // bit 4*j + k: word k of chunk j differs from the reference
#pragma unroll
    for (int j = 0; j < U; ++j) {
        const int g = j * 32 + lane;
        const int src = g / U, u = g - src * U;                          // owning read (lane of the tile) and unit inside it
        const int rel_s = __shfl_sync(0xffffffffu, rel, src & 31);
        const int L_s = __shfl_sync(0xffffffffu, L, src & 31);
        const bool ok = g < n_chunks && ((clean_mask >> src) & 1u);     // only clean reads inside the table have a valid offset
        const int nib = ok ? rel_s + 32 * u : 0;
        const uint32_t* rp = sm->sref + (nib >> 3);
        const uint32_t sh = (uint32_t)(nib & 7) * 4u;
        const uint32_t r0 = rp[0], r1 = rp[1], r2 = rp[2], r3 = rp[3], r4 = rp[4];
        uint32_t x0 = v[j].x ^ __funnelshift_r(r0, r1, sh), x1 = v[j].y ^ __funnelshift_r(r1, r2, sh);
        uint32_t x2 = v[j].z ^ __funnelshift_r(r2, r3, sh), x3 = v[j].w ^ __funnelshift_r(r3, r4, sh);
        // padding: words of this unit with any valid base, and the partial mask of the read's last word
        const int nvw = ((L_s + 7) >> 3) - 4 * u;
        const uint32_t pm = (L_s & 7) ? (0xffffffffu >> ((8 - (L_s & 7)) * 4)) : 0xffffffffu;
        x0 = nvw <= 0 ? 0u : (nvw == 1 ? x0 & pm : x0);
        x1 = nvw <= 1 ? 0u : (nvw == 2 ? x1 & pm : x1);
        x2 = nvw <= 2 ? 0u : (nvw == 3 ? x2 & pm : x2);
        x3 = nvw <= 3 ? 0u : (nvw == 4 ? x3 & pm : x3);
        uint32_t wm = (x0 ? 1u : 0u) | (x1 ? 2u : 0u) | (x2 ? 4u : 0u) | (x3 ? 8u : 0u);
        if (!ok) wm = 0u;
        all |= wm << (4 * j);
    }
    while (all) {                                                        // rare: queue the mismatching words for phase A2
        const int b = __ffs(all) - 1; all &= all - 1;
        const int g = (b >> 2) * 32 + lane;
        const int src = g / U, u = g - src * U;
        const uint32_t e = atomicAdd(Q.n_words, 1u);
        if (e < (uint32_t)kWords2) sm->wlist[e] = ((uint32_t)(i0 + src) << 5) | (uint32_t)(4 * u + (b & 3)); else *Q.overflow = 1u;
    }
}

// Phase A1: warp per tile of 32 reads.
__device__ __forceinline__ void phase_scan(const SessCtx& c, Smem2* sm, const Queues& Q, uint32_t& n_reads, uint32_t& n_bases) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nn = c.n_range - c.nt;
    const int tiles_t = (c.nt + 31) >> 5, tiles = tiles_t + ((nn + 31) >> 5);
    auto tile_i0 = [&](int t) { return t < tiles_t ? 32 * t : c.nt + 32 * (t - tiles_t); };
    auto tile_nv = [&](int t) { return min(32, (t < tiles_t ? c.nt : c.n_range) - tile_i0(t)); };
    for (int t = warp; t < tiles; t += kThreads / 32) {
        ReadMeta m = {};
        if (lane < tile_nv(t)) m = load_meta(c, tile_i0(t) + lane);
        const int i0 = tile_i0(t), n_valid = tile_nv(t);
        const bool valid = lane < n_valid;
        const int units = (int)(((m.lf & 0xffffu) + 31u) >> 5);
        const int U0 = __shfl_sync(0xffffffffu, units, 0);
        const uint32_t so0 = __shfl_sync(0xffffffffu, m.so, 0);
        const bool fits = !valid || (units == U0 && m.so == so0 + (uint32_t)(lane * U0));
        if (__all_sync(0xffffffffu, fits) && (U0 == 5 || U0 == 4 || U0 == 3 || U0 == 8)) {
            switch (U0) {
                case 5: scan_tile<5>(c, sm, Q, i0, n_valid, lane, m, so0, n_reads, n_bases); break;
                case 4: scan_tile<4>(c, sm, Q, i0, n_valid, lane, m, so0, n_reads, n_bases); break;
                case 3: scan_tile<3>(c, sm, Q, i0, n_valid, lane, m, so0, n_reads, n_bases); break;
                default: scan_tile<8>(c, sm, Q, i0, n_valid, lane, m, so0, n_reads, n_bases); break;
            }
        } else if (valid) {
            scan_read(c, sm, Q, i0 + lane, m, n_reads, n_bases);         // irregular tile: one read per lane
        }
    }
}

// ------------------------------------------------------------------ phase A2: one queued word of a clean read
__device__ __forceinline__ void discover_word(const SessCtx& c, Smem2* sm, const Queues& Q, uint32_t item) {
    const int i = (int)(item >> 5), k = (int)(item & 31u);
    const int64_t r = read_of(c, i);
    const int pos = __ldg(c.B.pos + r);
    const int L = (int)(__ldg(c.B.len_flag + r) & 0xffffu);
    const uint32_t rw = __ldg(rec_of(c, r) + k);
    const uint32_t fw = ref_word(c.B.ref4, (int64_t)pos + 8 * k);
    uint32_t x = (rw ^ fw) & tail_mask(L, k);
    const uint32_t ds = i < c.nt ? 0u : 1u;
    const int colb = pos - c.d.col_begin + 8 * k;
    while (x) {
        const int n = (__ffs(x) - 1) >> 2;
        x &= ~(0xfu << (n * 4));
        const uint32_t b = (rw >> (n * 4)) & 15u, rf = (fw >> (n * 4)) & 15u;
        if (b != 15u && is_acgt(rf)) push_entry(sm, Q, i, colb + n, b, ds);   // variation_classifier.py:147-150
    }
}

// ------------------------------------------------------------------ phase A2: one queued non-clean read per warp
// Lane w owns query words w, w+32, ...: it walks the CIGAR (a warp-uniform loop) and compares the part of every
// aligned segment that overlaps its 8 bases with the reference; lane 0 records the I/D observations
// (variation_classifier.py:52-107: pos, in_read_pos with the H/N quirk, Python-slice clamped allele).
__device__ void discover_generic_warp(const SessCtx& c, Smem2* sm, const Queues& Q, int i, int lane, uint32_t& n_reads, uint32_t& n_bases) {
    const int64_t r = read_of(c, i);
    const uint32_t c0 = __ldg(c.B.cigar_off + r), c1 = __ldg(c.B.cigar_off + r + 1);
    const int pos = __ldg(c.B.pos + r);
    const int L = (int)(__ldg(c.B.len_flag + r) & 0xffffu);
    const uint32_t ds = i < c.nt ? 0u : 1u;
    int span = 0;
    for (uint32_t ci = c0; ci < c1; ++ci) {
        const uint32_t w = __ldg(c.B.cigar + ci), op = w & 15u;
        if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) span += (int)(w >> 4);
    }
    if (pos + span <= c.first) return;
    if (lane == 0) { n_reads += 1u; n_bases += (uint32_t)L; }
    if ((int64_t)pos + span > c.B.ref_len || pos < 0 || pos < c.d.col_begin || pos + span - c.d.col_begin >= c.d.n_cols) {
        if (lane == 0) raise_error(c.totals, GA_ERR_OFFSET_RANGE, (uint32_t)r);
        return;
    }
    if (L > 8 * kStageWords) { *Q.overflow = 1u; return; }              // very long read: the fallback kernel takes the session
    const uint32_t* rec = rec_of(c, r);
    // ---- SNV candidates, one 8-base word per lane
    for (int w = lane; w < ((L + 7) >> 3); w += 32) {
        const int qb = w << 3;
        const uint32_t v = __ldg(rec + w);
        int rc = pos, q = 0;
        for (uint32_t ci = c0; ci < c1; ++ci) {
            const uint32_t cw = __ldg(c.B.cigar + ci), op = cw & 15u;
            const int ln = (int)(cw >> 4);
            if (op == 0u || op == 7u || op == 8u) {
                const int lo = max(q, qb), hi = min(min(q + ln, qb + 8), L);
                if (lo < hi) {
                    const int p0 = rc - q + qb;                           // reference position of query base qb under this segment
                    const uint32_t fw = ref_word(c.B.ref4, (int64_t)p0);
                    uint32_t mask = 0xffffffffu;
                    if (lo > qb) mask &= 0xffffffffu << ((lo - qb) * 4);
                    if (hi < qb + 8) mask &= 0xffffffffu >> ((qb + 8 - hi) * 4);
                    uint32_t x = (v ^ fw) & mask;
                    while (x) {
                        const int n = (__ffs(x) - 1) >> 2;
                        x &= ~(0xfu << (n * 4));
                        const uint32_t b = (v >> (n * 4)) & 15u, rf = (fw >> (n * 4)) & 15u;
                        if (b != 15u && is_acgt(rf)) push_entry(sm, Q, i, p0 + n - c.d.col_begin, b, ds);   // variation_classifier.py:147-150
                    }
                }
                q += ln; rc += ln;
            } else if (op == 1u || op == 4u) q += ln;
            else if (op == 2u || op == 3u) rc += ln;
            if (q >= qb + 8) break;
        }
    }
    // ---- indel observations
    if (lane != 0) return;
    int rc = pos, q = 0, ccl = 0, rcb = 0;
    for (uint32_t ci = c0; ci < c1; ++ci) {
        const uint32_t w = __ldg(c.B.cigar + ci), op = w & 15u;
        const int ln = (int)(w >> 4);
        if (op == 0u || op == 7u || op == 8u) {
            if (q + ln > L) { raise_error(c.totals, GA_ERR_OFFSET_RANGE, (uint32_t)r); return; }   // IndexError in variation_classifier.py:148
            q += ln; rc += ln; ccl += ln;
        } else if (op == 1u || op == 2u) {
            const uint32_t slot = atomicAdd(Q.n_obs, 1u);
            if (slot >= (uint32_t)kObs2) { *Q.overflow = 1u; return; }
            const int col = rc - c.d.col_begin;
            sm->o_col[slot] = col;
            const uint32_t meta = (op == 1u ? kMetaIns : 0u) | (ds ? kMetaDs : 0u) | ((uint32_t)ln & kMetaLenMask);
            const int irp = ccl + rcb;                                    // variation_classifier.py:82
            const int alen = allele_len(meta, irp, L);                    // Python-slice clamped (variation_classifier.py:87-88)
            uint32_t s0 = 0u, s1 = 0u;
            for (int j = 0; j < alen && j < 16; ++j) {
                const uint32_t code = read_code(rec, irp + j);
                if (j < 8) s0 |= code << (4 * j); else s1 |= code << (4 * (j - 8));
            }
            sm->o_meta[slot] = meta;
            sm->o_read[slot] = (uint32_t)i | ((uint32_t)alen << 16);
            sm->o_irp[slot] = irp;
            sm->o_sig0[slot] = s0; sm->o_sig1[slot] = s1;
            // push on the column's chain: 16-bit heads are updated with a CAS on the containing word
            {
                uint32_t* hw = reinterpret_cast<uint32_t*>(sm->ihead) + (col >> 1);
                const int shift = (col & 1) * 16;
                uint32_t old = *hw, assumed;
                do {
                    assumed = old;
                    sm->o_next[slot] = (int16_t)((assumed >> shift) & 0xffffu);
                    __threadfence_block();
                    old = atomicCAS(hw, assumed, (assumed & ~(0xffffu << shift)) | ((slot & 0xffffu) << shift));
                } while (old != assumed);
            }
            if (op == 1u) { q += ln; rcb += ln; } else { rc += ln; ccl += ln; rcb -= ln; }
        } else if (op == 3u) { rc += ln; ccl += ln; }
        else if (op == 4u) { q += ln; rcb += ln; }
        else if (op == 5u) { rcb += ln; }
    }
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
template <class SM> __device__ __noinline__ int collect_edits(const SessCtx& c, const SM* sm, int k, int L, Edit* edits, int* n_edits, int* n_dels, bool* too_many) {
    int16_t slots[GA_MAX_EDITS];
    int ns = 0;
#pragma unroll 1
    for (int o = sm->mhead[k]; o >= 0; o = sm->o_rnext[o]) {
        if (ns >= GA_MAX_EDITS) { *too_many = true; break; }
        int p = ns++;
#pragma unroll 1
        while (p > 0 && slots[p - 1] > o) { slots[p] = slots[p - 1]; --p; }
        slots[p] = (int16_t)o;
    }
    int ne = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll 1
        for (int q = 0; q < ns; ++q) {
            const int o = slots[q];
            const uint32_t m = sm->o_meta[o];
            if (((m & kMetaIns) != 0u) != (pass == 1)) continue;
            edits[ne].irp = sm->o_irp[o]; edits[ne].len = (int)(m & kMetaLenMask); edits[ne].pos = sm->o_col[o] + c.d.col_begin;
            edits[ne].is_ins = pass ? 1u : 0u; edits[ne].mean = 0u; ++ne;
        }
        if (pass == 0) *n_dels = ne;
    }
    const int n_del = *n_dels;
    int cur = L;
#pragma unroll 1
    for (int q = 0; q < n_del; ++q) {
        edits[q].p_eff = edits[q].irp < cur ? edits[q].irp : cur;
        edits[q].e_eff = edits[q].p_eff + edits[q].len;
        cur += edits[q].len;
    }
#pragma unroll 1
    for (int q = n_del; q < ne; ++q) {
        const int p = edits[q].irp < cur ? edits[q].irp : cur;
        const int e = edits[q].irp + edits[q].len < cur ? edits[q].irp + edits[q].len : cur;
        edits[q].p_eff = p; edits[q].e_eff = e > p ? e : p;
        cur -= (edits[q].e_eff - p);
    }
    *n_edits = ne;
    return cur;
}

// ------------------------------------------------------------------ warp-cooperative emission
// Lane w owns query words w, w+32, ...: it takes the record word, finds the aligned (M/=/X) segments that
// overlap its 8 bases, and replaces every base whose allele is in the germline set by the reference base
// (anonymizer_methods.py:170-176).  Loads and stores are coalesced.
template <class Store>
__device__ __forceinline__ void masked_words(const SessCtx& c, const Smem2* sm, int64_t r, int pos, int L, uint32_t c0, uint32_t c1,
                                             int n_words, int lane, int stride, Store&& store) {
    const uint32_t* rec = rec_of(c, r);
    for (int w = lane; w < n_words; w += stride) {
        const int qb = w << 3;
        uint32_t v = qb < L ? (__ldg(rec + w) & tail_mask(L, w)) : 0u;
        if (qb < L) {
            int rc = pos, q = 0;
            for (uint32_t ci = c0; ci < c1; ++ci) {
                const uint32_t cw = __ldg(c.B.cigar + ci), op = cw & 15u;
                const int ln = (int)(cw >> 4);
                if (op == 0u || op == 7u || op == 8u) {
                    const int lo = max(q, qb), hi = min(min(q + ln, qb + 8), L);
                    if (lo < hi) {
                        const int p0 = rc - q + qb;                       // reference position of query base qb under this segment
                        const uint32_t fw = ref_word(c.B.ref4, (int64_t)p0);
                        uint32_t mask = 0xffffffffu;
                        if (lo > qb) mask &= 0xffffffffu << ((lo - qb) * 4);
                        if (hi < qb + 8) mask &= 0xffffffffu >> ((qb + 8 - hi) * 4);
                        uint32_t x = (v ^ fw) & mask;
                        while (x) {
                            const int n = (__ffs(x) - 1) >> 2;
                            x &= ~(0xfu << (n * 4));
                            const uint32_t b = (v >> (n * 4)) & 15u;
                            if (b != 15u && ((sm->snv[p0 + n - c.d.col_begin] >> b) & 1u))
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

__device__ __forceinline__ void write_record_meta(const ResultView& O, uint64_t rec_idx, int s, int64_t r, int new_len, uint64_t seq16, uint32_t qual16) {
    O.mod_session[rec_idx] = s;
    O.mod_read[rec_idx] = (int32_t)r;
    O.mod_len[rec_idx] = (uint32_t)new_len;
    O.mod_seq_off16[rec_idx] = (uint32_t)seq16;
    O.mod_qual_off16[rec_idx] = qual16;
}

// Base code at original query index j after SNV masking (anonymizer_methods.py:170-176).
__device__ __forceinline__ uint32_t masked_base2(const SessCtx& c, const Smem2* sm, const uint32_t* rec, uint32_t c0, uint32_t c1, int pos, int j) {
    const uint32_t b = read_code(rec, j);
    if (b == 15u) return b;
    int rc = pos, q = 0;
    for (uint32_t ci = c0; ci < c1; ++ci) {
        const uint32_t w = __ldg(c.B.cigar + ci), op = w & 15u;
        const int ln = (int)(w >> 4);
        if (op == 0u || op == 7u || op == 8u) {
            if (j < q + ln) {
                const int rp = rc + (j - q);
                return ((sm->snv[rp - c.d.col_begin] >> b) & 1u) ? ref_code(c.B.ref4, rp) : b;
            }
            q += ln; rc += ln;
        } else if (op == 1u || op == 4u) { if (j < q + ln) return b; q += ln; }
        else if (op == 2u || op == 3u) rc += ln;
    }
    return b;
}

constexpr int kGroupStage = 32;        // words of SNV-masked input staged per lane group (reads up to 256 bases)

// General form (any number of edits, kept in a local array): used for the rare records with more than two edits.
// Indel-masked records, one per group of kGroup lanes (four records per warp at a time).  The SNV-masked input is
// staged in shared memory; every output word is then pulled through the backward index map of the record's
// edits (all DELs, then all INSs, at original offsets: anonymizer_methods.py:254-270, 178-203).  A word whose
// 8 bases (4 qualities) come from consecutive source positions is one funnel shift of two staged words;
// words that straddle an edit are assembled base by base.  `act` is false for the lanes of a group without a
// record; they only take part in the shuffles.
__device__ __noinline__ void emit_indel_group_slow(const SessCtx& c, Smem2* sm, const ResultView& O, bool act, int k, int i, uint64_t seq16, uint64_t qual16,
                                 int new_len, int glane, int group) {
    const int64_t r = act ? read_of(c, i) : 0;
    uint32_t lf = 0u, c0 = 0u, c1 = 0u; int pos = 0;
    if (act) { lf = __ldg(c.B.len_flag + r); c0 = __ldg(c.B.cigar_off + r); c1 = __ldg(c.B.cigar_off + r + 1); pos = __ldg(c.B.pos + r); }
    const int L = (int)(lf & 0xffffu);
    Edit edits[GA_MAX_EDITS];
    int ne = 0, n_del = 0; bool too_many = false;
    const uint8_t* qrec = nullptr;
    if (act) {
        collect_edits(c, sm, k, L, edits, &ne, &n_del, &too_many);
        qrec = qual_record_in(c.B, r, i < c.nt ? c.d.qt_begin : c.d.qn_begin, i < c.nt ? c.d.qt_end : c.d.qn_end);
        if (!qrec) { if (glane == 0) raise_error(c.totals, GA_ERR_BAD_ARGUMENT, (uint32_t)r); act = false; }
    }
    const bool reverse = ((lf >> 16) & 0x10u) != 0u;
    const uint32_t* rec = act ? rec_of(c, r) : nullptr;
    // ---- stage the SNV-masked input (o_sig0/o_sig1 are dead after phase R: 4 KB = 32 groups x 32 words)
    uint32_t* stage = sm->o_sig0 + group * kGroupStage;
    const bool staged = act && ((L + 7) >> 3) <= kGroupStage - 1;
    if (staged) {
        masked_words(c, sm, r, pos, L, c0, c1, (L + 7) >> 3, glane, kGroup, [&](int w, uint32_t v) { stage[w] = v; });
        if (glane == 0) stage[(L + 7) >> 3] = 0u;                      // the funnel shift may touch one word past the end
    }
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
        for (int q = 0; q < n_del; ++q) {
            const uint32_t m = n ? sum / n : 0u;
            edits[q].mean = m;
            sum += m * (uint32_t)edits[q].len; n += (uint32_t)edits[q].len;
            if (glane == 0 && (int64_t)edits[q].pos + edits[q].len > c.B.ref_len) raise_error(c.totals, GA_ERR_LENGTH_MISMATCH, (uint32_t)r);
        }
    }
    __syncwarp();                                                     // staged words visible to the group
    if (!act) return;
    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
    // ---- sequence words
    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
    for (int w = glane; w < units * 4; w += kGroup) {
        const int j0 = w << 3;
        uint32_t v = 0u;
        if (j0 < new_len) {
            int kin0 = 0, kin7 = 0;
            const int jl = min(j0 + 7, new_len - 1);
            const int s0 = map_back(edits, n_del, ne, j0, &kin0), s7 = map_back(edits, n_del, ne, jl, &kin7);
            if (staged && ne == 1 && s0 >= 0 && s7 - s0 == jl - j0) {  // one contiguous run of input bases (exact for a single edit)
                v = __funnelshift_r(stage[s0 >> 3], stage[(s0 >> 3) + 1], (uint32_t)(s0 & 7) * 4u);
                if (jl - j0 < 7) v &= 0xffffffffu >> ((7 - (jl - j0)) * 4);
            } else {
                for (int n = 0; n <= jl - j0; ++n) {
                    int kin = 0;
                    const int src = map_back(edits, n_del, ne, j0 + n, &kin);
                    uint32_t code;
                    if (src < 0) code = ref_code(c.B.ref4, (int64_t)edits[-1 - src].pos + kin);
                    else if (staged) code = (stage[src >> 3] >> ((src & 7) * 4)) & 15u;
                    else code = masked_base2(c, sm, rec, c0, c1, pos, src);
                    v |= code << (n * 4);
                }
            }
        }
        oseq[w] = v;
    }
    // ---- quality words (printed order = reversed forward-orientation array for reverse reads, AM.py:95,213)
    uint32_t* oq = reinterpret_cast<uint32_t*>(O.out_qual + 32ull * qual16);
    const uint32_t* qw = reinterpret_cast<const uint32_t*>(qrec);
    for (int w = glane; w < units * 8; w += kGroup) {
        const int p0 = w << 2;
        uint32_t v = 0u;
        if (p0 < new_len) {
            const int pl = min(p0 + 3, new_len - 1);
            int kin0 = 0, kin3 = 0;
            const int f0 = reverse ? new_len - 1 - p0 : p0, f3 = reverse ? new_len - 1 - pl : pl;
            const int s0 = map_back(edits, n_del, ne, f0, &kin0), s3 = map_back(edits, n_del, ne, f3, &kin3);
            const int b0 = reverse ? L - 1 - s0 : s0;                   // byte of the BAM-order quality record
            const bool run = ne == 1 && s0 >= 0 && s3 >= 0 && (reverse ? s0 - s3 : s3 - s0) == pl - p0;
            if (run) {
                const uint32_t lo = __ldg(qw + (b0 >> 2)), hi = (b0 & 3) ? __ldg(qw + (b0 >> 2) + 1) : 0u;
                v = __funnelshift_r(lo, hi, (uint32_t)(b0 & 3) * 8u);
                if (pl - p0 < 3) v &= 0xffffffffu >> ((3 - (pl - p0)) * 8);
            } else {
                for (int n = 0; n <= pl - p0; ++n) {
                    int kin = 0;
                    const int f = reverse ? new_len - 1 - (p0 + n) : p0 + n;
                    const int src = map_back(edits, n_del, ne, f, &kin);
                    const uint32_t qv = src >= 0 ? (uint32_t)qrec[reverse ? L - 1 - src : src] : edits[-1 - src].mean;
                    v |= qv << (n * 8);
                }
            }
        }
        oq[w] = v;
    }
}

__device__ __noinline__ int new_len_slow(const SessCtx& c, const Smem2* sm, const ResultView& O, int k, int L, int64_t r) {
    Edit edits[GA_MAX_EDITS];
    int ne = 0, nd = 0; bool too_many = false;
    const int new_len = collect_edits(c, sm, k, L, edits, &ne, &nd, &too_many);
    if (too_many) raise_error(O.totals, GA_ERR_UNSUPPORTED, (uint32_t)r);
    return new_len;
}

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
template <class SM> __device__ __forceinline__ bool collect2(const SessCtx& c, const SM* sm, int k, int L, Ed2& E, int* new_len) {
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
        E.pos[q] = has ? sm->o_col[o[q]] + c.d.col_begin : 0; E.mean[q] = 0u;
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

// In-kernel form: the germline test is the session's shared-memory table.
__device__ void emit_indel_group(const SessCtx& c, Smem2* sm, const ResultView& O, bool act, const Ed2& E, int i, uint64_t seq16, uint64_t qual16,
                                 int new_len, int glane, int group) {
    const int64_t r = act ? read_of(c, i) : 0;
    emit_indel_group_t(c.B, c.totals, O, act, E, r, c.d.col_begin, i < c.nt ? c.d.qt_begin : c.d.qn_begin, i < c.nt ? c.d.qt_end : c.d.qn_end,
                       sm->o_sig0 + group * kGroupStage, seq16, qual16, new_len, glane,
                       [&](int col, uint32_t b) { return ((sm->snv[col] >> b) & 1u) != 0u; });
}

// Block-wide exclusive scan of a 64-bit value per thread.
__device__ __forceinline__ unsigned long long block_exclusive_scan64(unsigned long long v, unsigned long long* tmp, unsigned long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long s = lane < kThreads / 32 ? tmp[lane] : 0ull;
        unsigned long long si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long n = __shfl_up_sync(0xffffffffu, si, d);
            if (lane >= d) si += n;
        }
        if (lane < kThreads / 32) tmp[lane] = si - s;
        if (lane == kThreads / 32 - 1) tmp[kThreads / 32] = si;
    }
    __syncthreads();
    *total = tmp[kThreads / 32];
    return tmp[warp] + inc - v;
}

// ------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kThreads, 4) session_kernel_v2(BatchView B, SessView S, const SessionDesc* __restrict__ descs,
                                                                 int32_t* __restrict__ big_list, int32_t* __restrict__ n_big,
                                                                 ResultView O, unsigned int* __restrict__ ticket, int stop_after, EmitScratch X) {
    // stop_after: profiling knob (GA_STOP_AFTER, 0 = run everything): low byte = sessions end after phase 1=A1 2=A2 3=R 4=M 5=L
    // 6=B1 71/72=inside B2 7=B2; bit 8 switches the L2 bulk prefetch of the next session on
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Smem2* sm = reinterpret_cast<Smem2*>(smem_raw);
    __shared__ unsigned long long s_scan64[kThreads / 32 + 1];
    __shared__ uint32_t s_scan[kThreads / 32 + 1];
    __shared__ uint32_t s_nobs, s_nent, s_nwords, s_ngen, s_nx, s_ngerm, s_reads, s_bases, s_cnt[3], s_overflow;
    __shared__ int s_next_session;
    __shared__ unsigned long long s_base[3];

    SessCtx c;
    c.B = B;
    c.totals = O.totals;
    uint32_t* const clist = sm->wlist;
    uint32_t* const msize = sm->lists; uint32_t* const mseq = sm->lists + kMod2; uint32_t* const mqual = sm->lists + 2 * kMod2;
    memset(&c.T, 0, sizeof c.T);                                      // the shared-memory tables are reached through `sm`
    Queues Q; Q.n_words = &s_nwords; Q.n_gen = &s_ngen; Q.n_ent = &s_nent; Q.n_obs = &s_nobs; Q.overflow = &s_overflow;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int n_work = S.n_sessions;

    if (tid == 0) s_next_session = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    for (;;) {
        const int s = s_next_session;
        if (s >= n_work) break;
        c.d = descs[s];
        const int n_cols = c.d.n_cols;
        c.s = s;
        c.nt = c.d.t_end - c.d.t_begin;
        c.n_range = c.nt + (c.d.n_end - c.d.n_begin);
        const int n_cw = (c.n_range + 31) >> 5;
        // ---- zero the working set (skipped for sessions the fallback kernel owns)
        if (!c.d.big) {
            for (int k = tid; k < n_cols; k += kThreads) { sm->snv[k] = 0u; sm->ihead[k] = (int16_t)-1; }
            for (int k = tid; k < n_cw; k += kThreads) { sm->modbits[k] = 0u; sm->indelbits[k] = 0u; sm->genbits[k] = 0u; }
            const uint32_t* gref = B.ref4 + ((c.d.col_begin + 8) >> 3);
            for (int k = tid; k < (n_cols >> 3) + 8; k += kThreads) sm->sref[k] = __ldg(gref + k);   // + record padding, + funnel-shift lookahead
        }
        __syncthreads();                                              // every thread has read s_next_session
        if (tid == kThreads - 32) {                                   // next ticket, and its reads on their way into L2
            const int nx = (int)atomicAdd(ticket, 1u);
            s_next_session = nx;
            if (nx < n_work && (stop_after & 0x100)) prefetch_session(B, descs[nx], !(stop_after & 0x200));   // knob bit 8: L2 prefetch of the next session (off by default: measured slower), bit 9: meta only
        }
        if (tid == 0) {
            s_nobs = 0; s_nent = 0; s_nwords = 0; s_ngen = 0; s_nx = 0; s_ngerm = 0; s_reads = 0; s_bases = 0; s_cnt[0] = s_cnt[1] = s_cnt[2] = 0; s_overflow = 0;
        }
        if (c.d.big) { __syncthreads(); continue; }
        c.first = S.first[s];
        c.keep_type = S.keep_type[s]; c.keep_pos = S.keep_pos[s]; c.keep_end = S.keep_end[s]; c.keep_len = S.keep_len[s];
        c.keep_allele = S.keep_alleles + S.keep_allele_off[s];
        c.keep_alen = (int)(S.keep_allele_off[s + 1] - S.keep_allele_off[s]);
        __syncthreads();

        // ---- phase A1: scan every read, queue what needs a closer look
        uint32_t n_reads = 0, n_bases = 0;
        phase_scan(c, sm, Q, n_reads, n_bases);
        __syncthreads();
        if ((stop_after & 0xff) == 1) { __syncthreads(); continue; }
        // ---- phase A2: dense discovery over the queues
        {
            const int nw = min((int)s_nwords, kWords2), ng = min((int)s_ngen, kGen2);
            for (int k = tid; k < nw; k += kThreads) discover_word(c, sm, Q, sm->wlist[k]);
            for (int k = warp; k < ng; k += kThreads / 32) discover_generic_warp(c, sm, Q, (int)sm->glist[k], lane, n_reads, n_bases);
            const uint32_t wr = warp_sum(n_reads), wb = warp_sum(n_bases);
            if (lane == 0 && wr) { atomicAdd(&s_reads, wr); atomicAdd(&s_bases, wb); }
        }
        __syncthreads();
        if (s_overflow) {                                             // queues or tables too small: hand the session to the fallback kernel
            if (tid == 0) big_list[atomicAdd(n_big, 1)] = s;
            __syncthreads();
            continue;
        }
        const int n_obs = (int)s_nobs;
        const int n_ent = (int)s_nent;

        if ((stop_after & 0xff) == 2) { __syncthreads(); continue; }
        // ---- phase R: germline = seen in tumor AND normal, minus variant_to_keep (AM.py:546-547)
        {
            uint32_t keep_bit = 0u; int keep_col = -1;
            if (c.keep_type == GA_VT_SNV && c.keep_end == c.keep_pos && c.keep_len == 1 && c.keep_alen == 1) {
                const char* code2asc = "=ACMGRSVTWYHKDBN";
                const uint8_t ch = c.keep_allele[0];
                for (int k = 0; k < 16; ++k) if ((uint8_t)code2asc[k] == ch) { keep_bit = 1u << k; keep_col = c.keep_pos - c.d.col_begin; }
            }
            uint32_t cnt = 0;
            for (int k = tid; k < n_cols; k += kThreads) {
                const uint32_t w = sm->snv[k];
                uint32_t g = (w & (w >> 16)) & 0xffffu;
                if (k == keep_col) g &= ~keep_bit;
                sm->snv[k] = g;
                cnt += __popc(g);
                while (g) {                                           // hand the allele to the emission kernel
                    const uint32_t code = (uint32_t)(__ffs(g) - 1); g &= g - 1;
                    const uint32_t e = atomicAdd(&s_ngerm, 1u);
                    if (e < (uint32_t)kGermCap) X.germ[(size_t)s * kGermCap + e] = ((uint32_t)k << 4) | code;
                }
            }
            cnt = warp_sum(cnt);
            if (lane == 0 && cnt) atomicAdd(&s_cnt[0], cnt);
        }
        for (int o = tid; o < n_obs; o += kThreads) {                 // indels: exact key equality (variants.py:83-96)
            const uint32_t m = sm->o_meta[o];
            bool germ = false, rep = true;
            for (int o2 = sm->ihead[sm->o_col[o]]; o2 >= 0; o2 = sm->o_next[o2]) {
                if (o2 == o) continue;
                if (!obs_equal2(c, sm, o, o2)) continue;
                if ((sm->o_meta[o2] ^ m) & kMetaDs) germ = true;
                if (o2 < o) rep = false;
            }
            if (germ && obs_equals_keep2(c, sm, o)) germ = false;
            if (germ) {
                atomicOr(&sm->o_meta[o], kMetaGerm | (rep ? kMetaRep : 0u));
                if (rep) atomicAdd(&s_cnt[(m & kMetaIns) ? 2 : 1], 1u);
                const uint32_t i = (uint32_t)obs_read(sm, o);
                atomicOr(&sm->modbits[i >> 5], 1u << (i & 31));
                atomicOr(&sm->indelbits[i >> 5], 1u << (i & 31));
            }
        }
        __syncthreads();

        if ((stop_after & 0xff) == 3) { __syncthreads(); continue; }
        // ---- phase M: reads that carry a germline SNV allele
        for (int e = tid; e < n_ent; e += kThreads) {
            const uint32_t w = sm->ent[e];
            if ((sm->snv[(w >> 4) & 0xfffu] >> (w & 15u)) & 1u) {
                const uint32_t i = w >> 16;
                atomicOr(&sm->modbits[i >> 5], 1u << (i & 31));
            }
        }
        __syncthreads();

        if ((stop_after & 0xff) == 4) { __syncthreads(); continue; }
        // ---- phase L: ordered list of the modified reads (n_cw <= 128 bitmap words, one per thread)
        uint32_t n_mod;
        {
            const uint32_t bits = tid < n_cw ? sm->modbits[tid] : 0u;
            uint32_t off = block_exclusive_scan(__popc(bits), s_scan, &n_mod);
            if (tid < n_cw) sm->woff[tid] = off;
            uint32_t b = bits;
            while (b) {
                const int k = __ffs(b) - 1; b &= b - 1;
                if (off < (uint32_t)kMod2) clist[off] = (uint32_t)(tid * 32 + k);
                ++off;
            }
        }
        if (n_mod > (uint32_t)kMod2) {
            if (tid == 0) big_list[atomicAdd(n_big, 1)] = s;
            __syncthreads();
            continue;
        }
        for (int k = tid; k < (int)n_mod; k += kThreads) sm->mhead[k] = -1;
        __syncthreads();
        if (n_obs > 0) {                                              // hang every germline observation on its modified read
            for (int o = tid; o < n_obs; o += kThreads) {
                if (!(sm->o_meta[o] & kMetaGerm)) continue;
                const uint32_t i = (uint32_t)obs_read(sm, o);
                const uint32_t k = sm->woff[i >> 5] + __popc(sm->modbits[i >> 5] & ((1u << (i & 31)) - 1u));
                sm->o_rnext[o] = (int16_t)atomicExch(&sm->mhead[k], o);
            }
            __syncthreads();
        }

        if ((stop_after & 0xff) == 5) { __syncthreads(); continue; }
        // ---- phase B1: new length of every modified read; indel-masked reads need the edit analysis.  Records are
        // written by the emission kernel unless the session has more germline alleles than the hand-over holds
        // (then everything is emitted here) or the record has more than two edits.
        const bool inline_emit = s_ngerm > (uint32_t)kGermCap;
        const int per = ((int)n_mod + kThreads - 1) / kThreads;
        const int k0 = min(tid * per, (int)n_mod), k1 = min(k0 + per, (int)n_mod);
        unsigned long long mine = 0ull;                               // [records:16 | seq units:24 | qual units:24]
        for (int k = k0; k < k1; ++k) {
            const int i = (int)clist[k];
            const int64_t r = read_of(c, i);
            uint32_t m;
            bool slow = false;
            if ((sm->indelbits[i >> 5] >> (i & 31)) & 1u) {
                const int L0 = (int)(__ldg(c.B.len_flag + r) & 0xffffu);
                Ed2 E2;
                int new_len = L0;
                if (!collect2(c, sm, k, L0, E2, &new_len)) { new_len = new_len_slow(c, sm, O, k, L0, r); slow = true; }
                m = kModFlag | kQualFlag | ((uint32_t)new_len & kLen2);
            } else {
                m = kModFlag | (__ldg(c.B.len_flag + r) & 0xffffu);
            }
            if (slow) m |= kSlowFlag;
            msize[k] = m;
            const bool special = (m & kQualFlag) || ((sm->genbits[i >> 5] >> (i & 31)) & 1u);
            if (inline_emit ? special : slow) sm->glist[atomicAdd(&s_nx, 1u)] = (uint16_t)k;
            uint32_t units = ((m & kLen2) + 31u) / 32u; if (units < 1u) units = 1u;
            mine += (1ull << 48) | ((unsigned long long)units << 24) | ((m & kQualFlag) ? (unsigned long long)units : 0ull);
        }
        unsigned long long total;
        const unsigned long long off = block_exclusive_scan64(mine, s_scan64, &total);
        const uint32_t tot_rec = (uint32_t)(total >> 48), tot_seq = (uint32_t)((total >> 24) & 0xffffffu), tot_qual = (uint32_t)(total & 0xffffffu);
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
        {   // per-record output offsets (session-relative)
            uint32_t so = (uint32_t)((off >> 24) & 0xffffffu), qo = (uint32_t)(off & 0xffffffu);
            for (int k = k0; k < k1; ++k) {
                const uint32_t m = msize[k];
                uint32_t units = ((m & kLen2) + 31u) / 32u; if (units < 1u) units = 1u;
                mseq[k] = so; mqual[k] = qo;
                so += units; if (m & kQualFlag) qo += units;
            }
        }
        __syncthreads();
        const bool fits = (int64_t)(s_base[0] + tot_rec) <= O.cap_records && (int64_t)(s_base[1] + tot_seq) <= O.cap_seq16 &&
                          (int64_t)(s_base[2] + tot_qual) <= O.cap_qual16;
        if (!fits) { if (tid == 0) raise_error(O.totals, GA_ERR_CAPACITY, 0xffffffffu); __syncthreads(); continue; }

        if ((stop_after & 0xff) == 6) { __syncthreads(); continue; }
        // ---- phase B2: record headers and the hand-over to the emission kernel; in-kernel emission only for what the
        // hand-over cannot describe
        if (inline_emit) {
            // clean SNV-only records are plain copies, one 16-byte unit per thread and iteration (hits patched in B3)
            for (uint32_t idx = tid; idx < tot_seq; idx += kThreads) {
                int lo = 0, hi = (int)n_mod;                          // last record with mseq[k] <= idx
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (mseq[mid] <= idx) lo = mid; else hi = mid; }
                const uint32_t m = msize[lo];
                const int i = (int)clist[lo];
                if ((m & kQualFlag) || ((sm->genbits[i >> 5] >> (i & 31)) & 1u)) continue;
                const int u = (int)(idx - mseq[lo]), L = (int)(m & kLen2);
                const int64_t r = read_of(c, i);
                uint4 v = ldg128(reinterpret_cast<const uint4*>(c.B.seq4 + 16ull * __ldg(c.B.seq_off16 + r)) + u);
                if (32 * u + 32 > L) { v.x &= tail_mask(L, 4 * u); v.y &= tail_mask(L, 4 * u + 1); v.z &= tail_mask(L, 4 * u + 2); v.w &= tail_mask(L, 4 * u + 3); }
                *reinterpret_cast<uint4*>(O.out_seq4 + 16ull * (s_base[1] + idx)) = v;
            }
        }
        if ((stop_after & 0xff) == 71) { __syncthreads(); continue; }
        uint32_t n_q = 0;
        if (tid == 0) X.germ_n[s] = inline_emit ? 0u : s_ngerm;
        for (int k = tid; k < (int)n_mod; k += kThreads) {            // record headers, coalesced
            const uint32_t m = msize[k];
            const bool q = (m & kQualFlag) != 0u;
            const int i = (int)clist[k];
            n_q += q ? 1u : 0u;
            const uint32_t qual16 = q ? (uint32_t)(s_base[2] + mqual[k]) : 0xffffffffu;
            write_record_meta(O, s_base[0] + k, s, read_of(c, i), (int)(m & kLen2), s_base[1] + mseq[k], qual16);
            uint8_t kind = 0;
            if (!inline_emit && !(m & kSlowFlag)) kind = q ? 3 : (((sm->genbits[i >> 5] >> (i & 31)) & 1u) ? 2 : 1);
            X.kind[s_base[0] + k] = kind;
            if (kind == 3) {                                          // the edits travel in the record's (still unused) quality slot
                Ed2 E2; int nl = 0;
                collect2(c, sm, k, (int)(__ldg(c.B.len_flag + read_of(c, i)) & 0xffffu), E2, &nl);
                EditAux a;
                a.irp0 = E2.irp[0]; a.pos0 = E2.pos[0]; a.len0 = (uint32_t)E2.len[0] | ((E2.ne >= 1 && E2.n_del < 1) ? 0x80000000u : 0u);
                a.irp1 = E2.irp[1]; a.pos1 = E2.pos[1]; a.len1 = (uint32_t)E2.len[1] | ((E2.ne >= 2 && E2.n_del < 2) ? 0x80000000u : 0u);
                a.ne = (uint32_t)E2.ne; a.n_del = (uint32_t)E2.n_del;
                uint4* dst = reinterpret_cast<uint4*>(O.out_qual + 32ull * qual16);
                const uint4* src = reinterpret_cast<const uint4*>(&a);
                dst[0] = src[0]; dst[1] = src[1];
            }
        }
        if ((stop_after & 0xff) == 72) { __syncthreads(); continue; }
        {   // records that are not plain copies: one group of kGroup lanes each
            const int nx = (int)s_nx, group = tid / kGroup, glane = tid % kGroup;
            for (int base = 0; base < nx; base += kThreads / kGroup) {
                const int xi = base + group;
                const bool have = xi < nx;
                const int k = have ? (int)sm->glist[xi] : 0;
                const uint32_t m = have ? msize[k] : 0u;
                const int i = have ? (int)clist[k] : 0;
                const int L = (int)(m & kLen2);
                const uint64_t seq16 = s_base[1] + (have ? mseq[k] : 0u);
                const bool indel = have && (m & kQualFlag);
                if (__any_sync(0xffffffffu, indel)) {
                    Ed2 E2; E2.ne = 0; E2.n_del = 0;
                    int nl = L;
                    const bool fast = indel && collect2(c, sm, k, (int)(__ldg(c.B.len_flag + read_of(c, i)) & 0xffffu), E2, &nl);
                    const uint64_t qual16 = s_base[2] + (have ? mqual[k] : 0u);
                    emit_indel_group(c, sm, O, fast, E2, i, seq16, qual16, L, glane, group);
                    if (__any_sync(0xffffffffu, indel && !fast)) emit_indel_group_slow(c, sm, O, indel && !fast, k, i, seq16, qual16, L, glane, group);
                }
                if (have && !indel) {                                 // other CIGARs, SNV-only: re-walk, mask while copying
                    const int64_t r = read_of(c, i);
                    int units = (L + 31) >> 5; if (units < 1) units = 1;
                    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
                    const uint32_t c0 = __ldg(c.B.cigar_off + r), c1 = __ldg(c.B.cigar_off + r + 1);
                    masked_words(c, sm, r, __ldg(c.B.pos + r), L, c0, c1, units * 4, glane, kGroup, [&](int w, uint32_t v) { oseq[w] = v; });
                }
            }
        }
        n_q = warp_sum(n_q);
        if (lane == 0 && n_q) atomicAdd((unsigned long long*)&O.totals->indel_records, (unsigned long long)n_q);
        __syncthreads();                                              // the copies are in place

        if ((stop_after & 0xff) == 7) { __syncthreads(); continue; }
        // ---- phase B3 (in-kernel emission only): SNV masking of the copied clean reads, one 4-bit XOR per germline hit
        if (inline_emit)
        for (int e = tid; e < n_ent; e += kThreads) {
            const uint32_t w = sm->ent[e];
            const uint32_t col = (w >> 4) & 0xfffu, b = w & 15u, i = w >> 16;
            if (!((sm->snv[col] >> b) & 1u)) continue;
            if ((sm->genbits[i >> 5] >> (i & 31)) & 1u) continue;     // re-walked by its warp above
            const uint32_t k = sm->woff[i >> 5] + __popc(sm->modbits[i >> 5] & ((1u << (i & 31)) - 1u));
            const int64_t r = read_of(c, (int)i);
            const int q = (int)col + c.d.col_begin - __ldg(c.B.pos + r);
            const uint32_t rf = ref_code(c.B.ref4, (int64_t)col + c.d.col_begin);
            uint32_t* word = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * (s_base[1] + mseq[k])) + (q >> 3);
            atomicXor(word, (b ^ rf) << ((q & 7) * 4));
        }
        __syncthreads();                                              // tables are reused by the next session
    }
}

}  // namespace ga
