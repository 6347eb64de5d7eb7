// ga_scan_kernel.cuh - stage 1 of the streaming pipeline: allele discovery over every session read
// (north_star jobs (2) + (3a): CIGAR walk and comparison against the reference).
//
// Work item = the tumor (or normal) candidate reads of one session, taken by ONE WARP from a ticket counter; no
// block-level synchronisation anywhere.  The warp walks the item in tiles of up to 31 reads whose records are
// contiguous in seq4: the tile's record bytes are fetched by one TMA bulk copy (cp.async.bulk + mbarrier) into
// the warp's private two-stage shared-memory ring while the previous tile is being compared, the per-read
// meta words (pos, length, record offset, CIGAR offsets) are prefetched two tiles ahead into registers, and the
// session's reference window is staged once per item.  Lane = read: a clean read (one M/=/X op spanning the
// read - nearly all of them) is compared 32 bases per 128-bit shared-memory load; mismatching 8-base words are
// resolved right away from the staged bytes (variation_classifier.py:144-150); reads with any other CIGAR are
// walked by the whole warp (variation_classifier.py:52-107).
//
// Output (engine scratch, one fixed region per item, written by its warp only - no atomics):
//   ent[item][..]   SNV candidate entries  (read << 16) | (column << 4) | base code, bit 28 = read has another CIGAR,
//                   bits 29-30 = reference base (index into ACGT)
//   obs[item][..]   indel observations (ObsRec)
//   cnt[item]       {entries, observations, session reads, session bases}; entries == kCntOverflow hands the
//                   session to the fallback kernel
// The resolve kernel (ga_resolve_kernel.cuh) turns these into the germline set and the modified-record list.
#pragma once
#include "ga_record_ops.cuh"

namespace ga {

constexpr int kEntHalf = 2560;           // SNV candidate entries per item (tumor or normal half of a session): room for a 1 % mismatch rate at 60x
constexpr int kObsHalf = 384;            // indel observations per item
constexpr int kTileUnits = 156;          // 16-byte units staged per tile (31 reads of 150 bp and room to spare)
constexpr int kWbuf = 168;                // entries buffered per warp between flushes
constexpr uint32_t kEntGen = 1u << 28;   // entry flag: the read is not a clean single-op read
constexpr uint32_t kCntOverflow = 0xffffffffu;
constexpr int kScanThreads = 256;
constexpr int kSrefPad = 8;              // words of reference staged in front of the session's first column
constexpr int kLaneOps = 12;             // CIGAR ops of a read that the lane-per-read walk takes (more: the whole warp walks the read)

// Nibble masks of a partial 32-base unit: row nv = the four 8-base words of a unit of which only the first nv bases exist.
struct UnitMasks { uint32_t m[32][4]; };
constexpr UnitMasks make_unit_masks() {
    UnitMasks t{};
    for (int nv = 0; nv < 32; ++nv)
        for (int k = 0; k < 4; ++k) {
            const int left = nv - 8 * k;
            t.m[nv][k] = left >= 8 ? 0xffffffffu : (left <= 0 ? 0u : (0xffffffffu >> ((8 - left) * 4)));
        }
    return t;
}
__constant__ UnitMasks c_unit_masks = make_unit_masks();

struct ObsRec { int32_t col; uint32_t meta; uint32_t read_alen; int32_t irp; uint32_t sig0, sig1, qord, pad1; };   // qord: the read's ordinal among the item's reads with an I/D op
static_assert(sizeof(ObsRec) == 32, "ObsRec is two 16-byte stores");

struct ScanScratch {
    uint32_t* ent;      // [2 * n_sessions][kEntHalf]
    ObsRec* obs;        // [2 * n_sessions][kObsHalf]
    uint4* cnt;         // [2 * n_sessions]
};

struct WarpSmem {
    uint4 ring[2][kTileUnits];           // record bytes of the tile in flight and the tile being compared
    uint32_t sref[kCols2 / 8 + 16];      // 4-bit reference of the session's columns: word kSrefPad = the ref4 word of column col_begin, 64 bases of
                                         // padding in front (a segment's first 32-base unit may start before the window) and 64 behind
    uint32_t wbuf[kWbuf];                // entries waiting for the next coalesced flush
    uint64_t bar[2];                     // mbarriers of the two ring stages
    uint32_t wcnt, ovf;
    uint32_t n_ent, n_obs, n_qord;       // the item's running counts (lane 0 writes them; rare paths only, so not in registers)
    // warp-uniform facts about the item, read where they are used (one LDS each) instead of living in registers:
    // the tile pipeline keeps three tiles of per-read meta words in flight and needs the registers
    int begin;                           // first read of the item (read indices fit 31 bits: ga_result.mod_read is int32)
    int n;                               // reads of the item
    int i_base;                          // session-relative id of read `begin`
    int first;                           // region start of the session
    int relbase;                         // ((col_begin + 8) & ~7) - 8 * kSrefPad: reference nibble index of sref word 0
    uint32_t item;                       // 2 * session + dataset; the ent / obs regions are derived from it
    uint32_t pad;
};
static_assert(sizeof(WarpSmem) % 16 == 0, "per-warp slices stay 16-byte aligned");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP.S.G)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Per-lane meta words of one read of a tile: four registers, three tiles' worth live at a time without spilling,
// and nothing computed from them at load time (the loads stay in flight for two tiles).  A tile has at most 31
// reads: the lane behind the last read loads the CIGAR offset that closes the last read's op range, so a read's op
// count is the neighbour lane's c0 minus its own.
struct TileMeta { int pos; uint32_t lf, so, c0; };
__device__ __forceinline__ uint32_t meta_len(const TileMeta& m) { return m.lf & 0xffffu; }
__device__ __forceinline__ uint32_t meta_ops(const TileMeta& m) { return __shfl_down_sync(0xffffffffu, m.c0, 1) - m.c0; }   // lanes holding a read

// Everything a warp knows about the item it is working on.
struct ItemCtx {
    BatchView B;
    WarpSmem* ws;
    ga_totals* totals;
    int col_begin, n_cols;
    bool table_in_ref;      // col_begin >= 0 and col_begin + n_cols <= ref_len
    ScanScratch X;
    uint32_t n_reads, n_bases;                 // per lane (summed at the end)
};

__device__ __forceinline__ void push_entry_w(const ItemCtx& c, int i, int col, uint32_t b, uint32_t rf, uint32_t gen) {
    const uint32_t slot = atomicAdd(&c.ws->wcnt, 1u);
    if (slot < (uint32_t)kWbuf) c.ws->wbuf[slot] = gen | ((uint32_t)(__ffs(rf) - 1) << 29) | ((uint32_t)i << 16) | ((uint32_t)col << 4) | b;
    else c.ws->ovf = 1u;
}

// Coalesced flush of the buffered entries into the item's region.
__device__ __forceinline__ void flush_entries(ItemCtx& c, int lane) {
    __syncwarp();
    const uint32_t n = min(c.ws->wcnt, (uint32_t)kWbuf), n_ent = c.ws->n_ent;
    if (n_ent + n > (uint32_t)kEntHalf) c.ws->ovf = 1u;
    else { uint32_t* ent = c.X.ent + (size_t)c.ws->item * kEntHalf + n_ent; for (uint32_t k = lane; k < n; k += 32) ent[k] = c.ws->wbuf[k]; }
    __syncwarp();
    if (lane == 0) { c.ws->wcnt = 0u; c.ws->n_ent = n_ent + n; }
    __syncwarp();
}

// One read with a CIGAR of more than eight ops, walked by the whole warp (all arguments warp-uniform; rare).  Lane w
// owns query words w, w+32, ...: it compares the part of every aligned segment that overlaps its 8 bases with the
// reference; lane 0 records the I/D observations (variation_classifier.py:52-107: pos, in_read_pos with the H/N quirk,
// Python-slice clamped allele).  The caller has counted the read, checked its span and knows its quality ordinal.
__device__ __forceinline__ void scan_generic_read(ItemCtx& c, int i, int64_t r, int pos, int L, uint32_t c0, uint32_t c1, const uint32_t* rec, uint32_t qord, int lane) {
    const BatchView& B = c.B;
    // ---- SNV candidates, one 8-base word per lane
    for (int w = lane; w < ((L + 7) >> 3); w += 32) {
        const int qb = w << 3;
        const uint32_t v = rec[w];
        int rc = pos, q = 0;
        for (uint32_t ci = c0; ci < c1; ++ci) {
            const uint32_t cw = __ldg(B.cigar + ci), op = cw & 15u;
            const int ln = (int)(cw >> 4);
            if (op == 0u || op == 7u || op == 8u) {
                const int lo = max(q, qb), hi = min(min(q + ln, qb + 8), L);
                if (lo < hi) {
                    const int p0 = rc - q + qb;                           // reference position of query base qb under this segment
                    const uint32_t fw = ref_word(B.ref4, (int64_t)p0);
                    uint32_t mask = 0xffffffffu;
                    if (lo > qb) mask &= 0xffffffffu << ((lo - qb) * 4);
                    if (hi < qb + 8) mask &= 0xffffffffu >> ((qb + 8 - hi) * 4);
                    uint32_t x = (v ^ fw) & mask;
                    while (x) {
                        const int n = (__ffs(x) - 1) >> 2;
                        x &= ~(0xfu << (n * 4));
                        const uint32_t b = (v >> (n * 4)) & 15u, rf = (fw >> (n * 4)) & 15u;
                        if (b != 15u && is_acgt(rf)) push_entry_w(c, i, p0 + n - c.col_begin, b, rf, kEntGen);   // variation_classifier.py:147-150
                    }
                }
                q += ln; rc += ln;
            } else if (op == 1u || op == 4u) q += ln;
            else if (op == 2u || op == 3u) rc += ln;
            if (q >= qb + 8) break;
        }
    }
    // ---- indel observations (lane 0; the count stays warp-uniform through the shuffle below)
    if (lane == 0) {
        uint32_t n_obs = c.ws->n_obs;
        const int span_all = ref_span_of(B.cigar, c0, c1);               // an insertion behind the last aligned base is marked (kMetaTrail)
        int rc = pos, q = 0, ccl = 0, rcb = 0;
        for (uint32_t ci = c0; ci < c1; ++ci) {
            const uint32_t w = __ldg(B.cigar + ci), op = w & 15u;
            const int ln = (int)(w >> 4);
            if (op == 0u || op == 7u || op == 8u) {
                if (q + ln > L) { raise_error(c.totals, GA_ERR_OFFSET_RANGE, (uint32_t)r); break; }   // IndexError in variation_classifier.py:148
                q += ln; rc += ln; ccl += ln;
            } else if (op == 1u || op == 2u) {
                if (n_obs >= (uint32_t)kObsHalf) { c.ws->ovf = 1u; break; }
                const uint32_t meta = (op == 1u ? kMetaIns | (rc - pos == span_all ? kMetaTrail : 0u) : 0u) | ((c.ws->item & 1u) ? kMetaDs : 0u) | ((uint32_t)ln & kMetaLenMask);
                const int irp = ccl + rcb;                                // variation_classifier.py:82
                const int alen = allele_len(meta, irp, L);                // Python-slice clamped (variation_classifier.py:87-88)
                uint32_t s0 = 0u, s1 = 0u;                                 // signature: the first 16 allele bases
                if (alen > 0) {
                    const int w0 = irp >> 3;
                    const uint32_t sh = (uint32_t)(irp & 7) * 4u;
                    const uint32_t a0 = rec[w0], a1 = 8 * (w0 + 1) < L ? rec[w0 + 1] : 0u, a2 = 8 * (w0 + 2) < L ? rec[w0 + 2] : 0u;
                    s0 = __funnelshift_r(a0, a1, sh) & tail_mask(alen, 0);
                    s1 = __funnelshift_r(a1, a2, sh) & tail_mask(alen, 1);
                }
                uint4* dst = reinterpret_cast<uint4*>(c.X.obs + (size_t)c.ws->item * kObsHalf + n_obs);
                dst[0] = make_uint4((uint32_t)(rc - c.col_begin), meta, (uint32_t)i | ((uint32_t)alen << 16), (uint32_t)irp);
                dst[1] = make_uint4(s0, s1, qord, 0u);
                ++n_obs;
                if (op == 1u) { q += ln; rcb += ln; } else { rc += ln; ccl += ln; rcb -= ln; }
            } else if (op == 3u) { rc += ln; ccl += ln; }
            else if (op == 4u) { q += ln; rcb += ln; }
            else if (op == 5u) { rcb += ln; }
        }
        c.ws->n_obs = n_obs;
    }
}

// Reads with a CIGAR of at most eight ops, FOUR AT A TIME: a group of 8 lanes per read (north_star job (2)).
//   * one lane per op: the reference- and query-consumed lengths and the in_read_pos terms are turned into each op's
//     start offsets by a warp-shuffle inclusive scan over the 8 lanes of the group (variation_classifier.py:69-82 sums
//     them op by op; htslib's query_position is the same sum);
//   * every I / D lane writes its observation in parallel (slots ascend in CIGAR order, reads in tile order);
//   * lane k of the group then owns 8-base words k, k+8, ...: each aligned segment (broadcast from its op lane by
//     shuffle) is compared against the session's staged reference window where it overlaps the word.
// sel: the four source lanes (one byte each, 0xff = none).  Per-lane inputs describe the read the LANE holds.
__device__ __forceinline__ void scan_generic_quad(ItemCtx& c, uint32_t sel, int i_lane, int pos, int L, uint32_t c0, uint32_t n_ops, uint32_t so, uint32_t qord,
                                                  bool staged, int b, uint32_t sof, int lane) {
    const BatchView& B = c.B;
    const int gl = lane & 7, gbase = lane & ~7;
    const uint32_t src = (sel >> (8 * (lane >> 3))) & 0xffu;
    const bool act = src != 0xffu;
    const int sl = act ? (int)src : 0;
    const int g_pos = __shfl_sync(0xffffffffu, pos, sl), g_L = __shfl_sync(0xffffffffu, L, sl), g_i = __shfl_sync(0xffffffffu, i_lane, sl);
    const uint32_t g_c0 = __shfl_sync(0xffffffffu, c0, sl), g_nops = __shfl_sync(0xffffffffu, n_ops, sl);
    const uint32_t g_so = __shfl_sync(0xffffffffu, so, sl), g_qord = __shfl_sync(0xffffffffu, qord, sl);
    const uint32_t* rec = staged ? reinterpret_cast<const uint32_t*>(c.ws->ring[b] + (g_so - sof))
                                 : reinterpret_cast<const uint32_t*>(B.seq4 + 16ull * g_so);
    // ---- K2: one lane per op, inclusive scans over the group
    const bool has_op = act && (uint32_t)gl < g_nops;
    const uint32_t cw = has_op ? __ldg(B.cigar + g_c0 + gl) : 0u;
    const uint32_t op = has_op ? (cw & 15u) : 15u;
    const int ln = has_op ? (int)(cw >> 4) : 0;
    const bool aligned = op == 0u || op == 7u || op == 8u;
    int sq = (aligned || op == 1u || op == 4u) ? ln : 0;                   // query consumed: M = X I S
    int sr = (aligned || op == 2u || op == 3u) ? ln : 0;                   // reference consumed: M = X D N (= the reference's cigar-consumed length)
    int sb = (op == 1u || op == 4u || op == 5u) ? ln : (op == 2u ? -ln : 0);   // read-consuming bases: +I +S +H -D (quirk Q7)
    const int q_own = sq, r_own = sr, b_own = sb;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
        const int tq = __shfl_up_sync(0xffffffffu, sq, d, 8), tr = __shfl_up_sync(0xffffffffu, sr, d, 8), tb = __shfl_up_sync(0xffffffffu, sb, d, 8);
        if (gl >= d) { sq += tq; sr += tr; sb += tb; }
    }
    const int q0 = sq - q_own, r0 = sr - r_own, b0 = sb - b_own;          // offsets at which this op starts
    const int g_span = __shfl_sync(0xffffffffu, sr, gbase + 7);          // the read's reference span: an insertion that starts there ends the alignment (kMetaTrail)
    if (aligned && q0 + ln > g_L) raise_error(c.totals, GA_ERR_OFFSET_RANGE, (uint32_t)(c.ws->begin + (g_i - c.ws->i_base)));   // IndexError in variation_classifier.py:148
    // ---- indel observations
    const bool is_id = op == 1u || op == 2u;
    const uint32_t idm = __ballot_sync(0xffffffffu, is_id);
    if (idm) {
        const uint32_t n_obs0 = c.ws->n_obs;
        const uint32_t slot = n_obs0 + __popc(idm & ((1u << lane) - 1u));
        if (is_id) {
            if (slot >= (uint32_t)kObsHalf) c.ws->ovf = 1u;
            else {
                const uint32_t meta = (op == 1u ? kMetaIns | (r0 == g_span ? kMetaTrail : 0u) : 0u) | ((c.ws->item & 1u) ? kMetaDs : 0u) | ((uint32_t)ln & kMetaLenMask);
                const int irp = r0 + b0;                                  // variation_classifier.py:82
                const int alen = allele_len(meta, irp, g_L);              // Python-slice clamped (variation_classifier.py:87-88)
                uint32_t s0 = 0u, s1 = 0u;                                 // signature: the first 16 allele bases
                if (alen > 0) {
                    const int w0 = irp >> 3;
                    const uint32_t sh = (uint32_t)(irp & 7) * 4u;
                    const uint32_t a0 = rec[w0], a1 = 8 * (w0 + 1) < g_L ? rec[w0 + 1] : 0u, a2 = 8 * (w0 + 2) < g_L ? rec[w0 + 2] : 0u;
                    s0 = __funnelshift_r(a0, a1, sh) & tail_mask(alen, 0);
                    s1 = __funnelshift_r(a1, a2, sh) & tail_mask(alen, 1);
                }
                uint4* dst = reinterpret_cast<uint4*>(c.X.obs + (size_t)c.ws->item * kObsHalf + slot);
                dst[0] = make_uint4((uint32_t)(g_pos + r0 - c.col_begin), meta, (uint32_t)g_i | ((uint32_t)alen << 16), (uint32_t)irp);
                dst[1] = make_uint4(s0, s1, g_qord, 0u);
            }
        }
        __syncwarp();
        if (lane == 0) c.ws->n_obs = min(n_obs0 + (uint32_t)__popc(idm), (uint32_t)kObsHalf);
        __syncwarp();
    }
    // ---- SNV candidates (variation_classifier.py:147-150)
    // an aligned op travels as two words: (query start | length << 16) and the diagonal (reference minus query
    // offset); every other op travels with length 0 and drops out of the overlap test by itself
    const uint32_t seg_ql = aligned ? ((uint32_t)q0 | ((uint32_t)ln << 16)) : 0u;
    const int seg_diag = r0 - q0;
    const uint32_t al_mask = __ballot_sync(0xffffffffu, aligned);
    const int nw = act ? (g_L + 7) >> 3 : 0;
    const int nw_max = __reduce_max_sync(0xffffffffu, nw);
    const int relbase = c.ws->relbase;
    // op slots that hold an aligned op in at least one of the four groups
    const uint32_t slots = (al_mask | (al_mask >> 8) | (al_mask >> 16) | (al_mask >> 24)) & 0xffu;
    for (int wb = 0; wb < nw_max; wb += 8) {
        const int w = wb + gl;
        const bool mine = w < nw;
        const int qb = w << 3;
        const uint32_t v = mine ? rec[w] : 0u;
        uint32_t todo = slots;
#pragma unroll 1
        while (todo) {
            const int j = __ffs(todo) - 1; todo &= todo - 1u;
            const uint32_t o_ql = __shfl_sync(0xffffffffu, seg_ql, gbase + j);
            const int o_diag = __shfl_sync(0xffffffffu, seg_diag, gbase + j);
            const int o_q0 = (int)(o_ql & 0xffffu), o_end = o_q0 + (int)(o_ql >> 16);
            const int lo = max(o_q0, qb), hi = min(min(o_end, qb + 8), g_L);
            if (!mine || lo >= hi) continue;
            const int p0 = g_pos + o_diag + qb;                           // reference position of query base qb under this segment (>= pos - 7)
            const int nib = p0 + 8 - relbase;                             // nibble offset in the staged window (>= -7)
            const uint32_t fw = nib >= 0 ? __funnelshift_r(c.ws->sref[nib >> 3], c.ws->sref[(nib >> 3) + 1], (uint32_t)(nib & 7) * 4u)
                                         : (c.ws->sref[0] << ((uint32_t)(-nib) * 4u));
            const uint32_t mask = (0xffffffffu << ((lo - qb) * 4)) & (0xffffffffu >> ((qb + 8 - hi) * 4));
            uint32_t x = (v ^ fw) & mask;
            while (x) {
                const int n = (__ffs(x) - 1) >> 2;
                x &= ~(0xfu << (n * 4));
                const uint32_t bb = (v >> (n * 4)) & 15u, rf = (fw >> (n * 4)) & 15u;
                if (bb != 15u && is_acgt(rf)) push_entry_w(c, g_i, p0 + n - c.col_begin, bb, rf, kEntGen);
            }
        }
    }
}

// Reads with another CIGAR, LANE = READ (the records of the tile are staged in the ring).  Every lane walks the ops of its
// own read; the warp moves in ROUNDS, one aligned segment per lane and round (north_star job (2): the running sums below
// are the reference's cigar-consumed lengths, variation_classifier.py:69-82):
//   * an I / D op: the lane writes the observation (slots were handed out by a warp scan of the reads' I / D counts, so
//     they ascend in read order and, inside a read, in CIGAR order - the emission relies on it);
//   * an aligned op: the segment's 32-base units are compared against the staged reference window exactly like a clean
//     read's (128-bit shared-memory loads, one funnel shift per word) on the segment's own diagonal; words outside the
//     segment are dropped from the mismatch mask, the boundary words are cut to the segment nibble by nibble.
// A soft-clipped read costs one compare round, a read with one indel two: tens of warp instructions per tile instead of
// hundreds per read (the 8-lanes-per-read walk, kept for tiles that could not be staged).
__device__ __forceinline__ void scan_generic_lanes(ItemCtx& c, bool act, int i, int idx, int pos, int L, int span, uint32_t c0, uint32_t n_ops, uint32_t n_id, uint32_t cw0,
                                                   const uint32_t* rec, uint32_t qord, int lane) {
    const BatchView& B = c.B;
    uint32_t tot_id;
    const uint32_t n_obs0 = c.ws->n_obs;
    uint32_t slot = n_obs0 + warp_excl_scan(act ? n_id : 0u, lane, &tot_id);
    const int relbase = c.ws->relbase;
    const uint4* rec4 = reinterpret_cast<const uint4*>(rec);
    int q = 0, rr = 0, bsum = 0;                                         // query consumed, reference consumed, +I +S +H -D (quirk Q7)
    uint32_t ci = 0u;                                                    // the lane's next op
#pragma unroll 1
    for (;;) {
        // ---- every lane moves on to its next aligned op; what lies in front of it is handled on the way
        bool seg = false;
        int ln = 0;
#pragma unroll 1
        while (__any_sync(0xffffffffu, act && !seg && ci < n_ops)) {
            if (act && !seg && ci < n_ops) {
                const uint32_t cw = ci ? __ldg(B.cigar + c0 + ci) : cw0;
                const uint32_t op = cw & 15u;
                const int l = (int)(cw >> 4);
                ++ci;
                if (op == 0u || op == 7u || op == 8u) {
                    if (q + l > L) raise_error(c.totals, GA_ERR_OFFSET_RANGE, (uint32_t)(c.ws->begin + idx));   // IndexError in variation_classifier.py:148
                    else if (l > 0) { seg = true; ln = l; }
                } else if (op == 1u || op == 2u) {                       // indel observation (variation_classifier.py:52-107)
                    if (slot >= (uint32_t)kObsHalf) c.ws->ovf = 1u;
                    else {
                        const uint32_t meta = (op == 1u ? kMetaIns | (rr == span ? kMetaTrail : 0u) : 0u) | ((c.ws->item & 1u) ? kMetaDs : 0u) | ((uint32_t)l & kMetaLenMask);
                        const int irp = rr + bsum;                       // variation_classifier.py:82
                        const int alen = allele_len(meta, irp, L);       // Python-slice clamped (variation_classifier.py:87-88)
                        uint32_t s0 = 0u, s1 = 0u;                        // signature: the first 16 allele bases
                        if (alen > 0) {
                            const int w0 = irp >> 3;
                            const uint32_t sh = (uint32_t)(irp & 7) * 4u;
                            const uint32_t a0 = rec[w0], a1 = 8 * (w0 + 1) < L ? rec[w0 + 1] : 0u, a2 = 8 * (w0 + 2) < L ? rec[w0 + 2] : 0u;
                            s0 = __funnelshift_r(a0, a1, sh) & tail_mask(alen, 0);
                            s1 = __funnelshift_r(a1, a2, sh) & tail_mask(alen, 1);
                        }
                        uint4* dst = reinterpret_cast<uint4*>(c.X.obs + (size_t)c.ws->item * kObsHalf + slot);
                        dst[0] = make_uint4((uint32_t)(pos + rr - c.col_begin), meta, (uint32_t)i | ((uint32_t)alen << 16), (uint32_t)irp);
                        dst[1] = make_uint4(s0, s1, qord, 0u);
                    }
                    ++slot;
                    if (op == 1u) { q += l; bsum += l; } else { rr += l; bsum -= l; }
                } else if (op == 4u) { q += l; bsum += l; }
                else if (op == 3u) rr += l;
                else if (op == 5u) bsum += l;
            }
        }
        if (!__any_sync(0xffffffffu, seg)) break;
        // ---- one compare round: the lanes' segments on their own diagonals (reference = pos + rr - q + query)
        const int qs = q, qe = q + ln;
        const int u_lo = qs >> 5, n_u = seg ? ((qe - 1) >> 5) - u_lo + 1 : 0;
        const int rel = pos + rr - q + 32 * u_lo + 8 - relbase;          // nibble offset of query base 32 * u_lo in the staged window (>= 0: kSrefPad)
        const int n_max = __reduce_max_sync(0xffffffffu, n_u);
        uint32_t wm = 0u;                                                // bit k: 8-base word 4 * u_lo + k differs from the reference
        {
            const uint32_t* rp = c.ws->sref + (seg ? (rel >> 3) : 0);
            const uint32_t sh = (uint32_t)(rel & 7) * 4u;
            uint32_t prev = rp[0];
#pragma unroll 1
            for (int uu = 0; uu < n_max; ++uu) {
                if (uu < n_u) {
                    const uint4 v = rec4[u_lo + uu];
                    const uint32_t r1 = rp[4 * uu + 1], r2 = rp[4 * uu + 2], r3 = rp[4 * uu + 3], r4 = rp[4 * uu + 4];
                    const uint32_t x0 = v.x ^ __funnelshift_r(prev, r1, sh), x1 = v.y ^ __funnelshift_r(r1, r2, sh);
                    const uint32_t x2 = v.z ^ __funnelshift_r(r2, r3, sh), x3 = v.w ^ __funnelshift_r(r3, r4, sh);
                    prev = r4;
                    wm |= ((x0 ? 1u : 0u) | (x1 ? 2u : 0u) | (x2 ? 4u : 0u) | (x3 ? 8u : 0u)) << (4 * uu);
                }
            }
        }
        if (seg) {                                                       // only the words that hold bases of the segment
            const int lo_w = (qs >> 3) - 4 * u_lo, hi_w = ((qe - 1) >> 3) - 4 * u_lo;
            wm &= (0xffffffffu << lo_w) & (0xffffffffu >> (31 - hi_w));
        } else wm = 0u;
        while (wm) {                                                     // SNV candidates (variation_classifier.py:147-150)
            const int k = __ffs(wm) - 1; wm &= wm - 1;
            const int kw = 4 * u_lo + k, qb = kw << 3;
            const uint32_t rw = rec[kw];
            const int nib = rel + 8 * k;
            const uint32_t fw = __funnelshift_r(c.ws->sref[nib >> 3], c.ws->sref[(nib >> 3) + 1], (uint32_t)(nib & 7) * 4u);
            const int lo = max(qs, qb), hi = min(qe, qb + 8);
            uint32_t x = (rw ^ fw) & (0xffffffffu << ((lo - qb) * 4)) & (0xffffffffu >> ((qb + 8 - hi) * 4));
            const int colb = pos + rr - q + qb - c.col_begin;
            while (x) {
                const int n = (__ffs(x) - 1) >> 2;
                x &= ~(0xfu << (n * 4));
                const uint32_t bb = (rw >> (n * 4)) & 15u, rf = (fw >> (n * 4)) & 15u;
                if (bb != 15u && is_acgt(rf)) push_entry_w(c, i, colb + n, bb, rf, kEntGen);
            }
        }
        if (seg) { q += ln; rr += ln; }
    }
    __syncwarp();
    if (lane == 0 && tot_id) c.ws->n_obs = min(n_obs0 + tot_id, (uint32_t)kObsHalf);
    __syncwarp();
}

__device__ __forceinline__ TileMeta load_tile_meta(const ItemCtx& c, int t, int TR, int lane) {
    TileMeta m = {0, 0u, 0u, 0u};
    const int i = t * TR + lane;
    const int64_t r = (int64_t)c.ws->begin + i;
    if (lane < TR && i < c.ws->n) { m.pos = __ldg(c.B.pos + r); m.lf = __ldg(c.B.len_flag + r); m.so = __ldg(c.B.seq_off16 + r); }
    if (lane <= TR && i <= c.ws->n) m.c0 = __ldg(c.B.cigar_off + r);      // one lane more: cigar_off has n_reads + 1 entries
    return m;
}

// Starts the TMA copy of the tile's record bytes into ring stage `b` when the records are contiguous and fit.
__device__ __forceinline__ uint32_t issue_tile(const ItemCtx& c, const TileMeta& m, bool valid, int b, bool tma_ok, int lane) {
    const uint32_t L = meta_len(m);
    const uint32_t units = valid ? (L + 31u) >> 5 : 0u;
    const uint32_t sof = __shfl_sync(0xffffffffu, m.so, 0);
    const uint32_t rel = m.so - sof;                                  // huge when the record lies before the tile's first one
    const uint32_t end = valid ? (rel <= (uint32_t)kTileUnits ? rel + units : 0xffffffffu) : 0u;
    const uint32_t total = __reduce_max_sync(0xffffffffu, end);       // units from the first record to the end of the last
    const uint32_t staged = (tma_ok && total > 0u && total <= (uint32_t)kTileUnits) ? 1u : 0u;
    if (staged && lane == 0) {
        mbar_expect_tx(&c.ws->bar[b], total * 16u);
        bulk_g2s(c.ws->ring[b], c.B.seq4 + 16ull * sof, total * 16u, &c.ws->bar[b]);
    }
    return staged;
}

// Compare + discover one tile (lane = read).
// `prefetch` is called once, between the comparison and the mismatch handling: the kernel loads the meta words of
// the tiles ahead there, away from the first use of this tile's own prefetched words.
template <class Prefetch>
__device__ __forceinline__ void scan_tile_w(ItemCtx& c, int t, int TR, const TileMeta& m, uint32_t cw0, bool staged, int b, int lane, Prefetch&& prefetch) {
    const uint32_t sof = __shfl_sync(0xffffffffu, m.so, 0);                // first record unit of the tile (what issue_tile copied from)
    const int idx = t * TR + lane;
    // a placed-unmapped mate (flag 0x4) is in no session: htslib's pileup drops BAM_FUNMAP records
    const bool valid = lane < TR && idx < c.ws->n && !((m.lf >> 16) & 0x4u);
    const int i = c.ws->i_base + idx;
    const int pos = m.pos;
    const int L = (int)meta_len(m);
    const uint32_t n_ops = meta_ops(m);
    const bool one_op = valid && n_ops == 1u;
    // a single-op read whose span L stays inside the reference and the session table
    // (when the whole table lies inside the reference - checked once per item - two compares say the same)
    const bool spec = one_op && L <= 256 &&
                      (c.table_in_ref ? (uint32_t)(pos - c.col_begin) < (uint32_t)(c.n_cols - L) && L < c.n_cols
                                      : pos >= 0 && (int64_t)pos + L <= c.B.ref_len && pos >= c.col_begin && pos + L - c.col_begin < c.n_cols);
    const uint32_t op0 = cw0 & 15u;
    // lane = read works on records staged in the ring only (shared-memory loads); a tile that could not be staged
    // (records not contiguous, or longer than the ring) goes read by read through the whole-warp walk below
    const bool clean = staged && spec && (op0 == 0u || op0 == 7u || op0 == 8u) && ((int)(cw0 >> 4) == L);
    const bool in_sess = clean && pos + L > c.ws->first;                     // fetched by range but not reaching the region: skipped
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(c.ws->ring[b] + (clean ? m.so - sof : 0u));
    if (in_sess) { c.n_reads += 1u; c.n_bases += (uint32_t)L; }
    const int rel = pos + 8 - c.ws->relbase;                                 // nibble offset of base `pos` inside the staged window
    // Full 32-base units are compared without any masking; the last, partial unit once, after the loop.
    const int full = in_sess ? L >> 5 : 0;
    const int fmax = __reduce_max_sync(0xffffffffu, full);
    uint32_t wm = 0u;                                                    // bit k: 8-base word k differs from the reference
    {
        const uint32_t* rp = c.ws->sref + (in_sess ? (rel >> 3) : 0);
        const uint32_t sh = (uint32_t)(rel & 7) * 4u;
        uint32_t prev = rp[0];
        const uint4* rec4 = reinterpret_cast<const uint4*>(rec);
#define GA_CMP_UNIT(U)                                                                                                     \
        {                                                                                                                  \
            const uint4 v = rec4[U];                                                                                       \
            const uint32_t r1 = rp[4 * (U) + 1], r2 = rp[4 * (U) + 2], r3 = rp[4 * (U) + 3], r4 = rp[4 * (U) + 4];         \
            const uint32_t x0 = v.x ^ __funnelshift_r(prev, r1, sh), x1 = v.y ^ __funnelshift_r(r1, r2, sh);               \
            const uint32_t x2 = v.z ^ __funnelshift_r(r2, r3, sh), x3 = v.w ^ __funnelshift_r(r3, r4, sh);                 \
            prev = r4;                                                                                                     \
            wm |= ((x0 ? 1u : 0u) | (x1 ? 2u : 0u) | (x2 ? 4u : 0u) | (x3 ? 8u : 0u)) << (4 * (U));                        \
        }
        if (fmax == 4 && __all_sync(0xffffffffu, full == 4 || full == 0)) {   // 129..159-base reads (the common shape): unrolled, constant offsets
            if (full) { GA_CMP_UNIT(0) GA_CMP_UNIT(1) GA_CMP_UNIT(2) GA_CMP_UNIT(3) }
        } else {
#pragma unroll 1
            for (int u = 0; u < fmax; ++u)
                if (u < full) GA_CMP_UNIT(u)
        }
#undef GA_CMP_UNIT
        if (in_sess && (L & 31)) {                                       // the partial unit: padding nibbles masked
            const int u = full, nv = L & 31;                             // nv valid nibbles in this unit
            const uint4 v = rec4[u];
            const uint32_t r1 = rp[4 * u + 1], r2 = rp[4 * u + 2], r3 = rp[4 * u + 3], r4 = rp[4 * u + 4];
            const uint32_t* um = c_unit_masks.m[nv];                     // usually the same row for the whole warp
            const uint32_t x0 = (v.x ^ __funnelshift_r(prev, r1, sh)) & um[0], x1 = (v.y ^ __funnelshift_r(r1, r2, sh)) & um[1];
            const uint32_t x2 = (v.z ^ __funnelshift_r(r2, r3, sh)) & um[2], x3 = (v.w ^ __funnelshift_r(r3, r4, sh)) & um[3];
            wm |= ((x0 ? 1u : 0u) | (x1 ? 2u : 0u) | (x2 ? 4u : 0u) | (x3 ? 8u : 0u)) << (4 * u);
        }
    }
    prefetch();
    // ---- mismatching words: SNV candidates (variation_classifier.py:147-150), resolved from the staged bytes
    while (wm) {
        const int k = __ffs(wm) - 1; wm &= wm - 1;
        const uint32_t rw = rec[k];
        const int nib = rel + 8 * k;
        const uint32_t fw = __funnelshift_r(c.ws->sref[nib >> 3], c.ws->sref[(nib >> 3) + 1], (uint32_t)(nib & 7) * 4u);
        uint32_t x = rw ^ fw;
        if (8 * k + 8 > L) x &= tail_mask(L, k);                        // only the read's last word has padding nibbles
        const int colb = pos - c.col_begin + 8 * k;
        while (x) {
            const int n = (__ffs(x) - 1) >> 2;
            x &= ~(0xfu << (n * 4));
            const uint32_t bb = (rw >> (n * 4)) & 15u, rf = (fw >> (n * 4)) & 15u;
            if (bb != 15u && is_acgt(rf)) push_entry_w(c, i, colb + n, bb, rf, 0u);
        }
    }
    // ---- reads with any other CIGAR (and every read of a tile that could not be staged)
    const bool gen = valid && !clean;
    if (__any_sync(0xffffffffu, gen)) {
        // lane = read: reference span and I/D presence of its ops, then the checks the reference's walk implies
        int span = 0;
        uint32_t n_id = 0u;
        if (gen) {
            for (uint32_t ci = 0; ci < n_ops; ++ci) {
                const uint32_t w = ci ? __ldg(c.B.cigar + m.c0 + ci) : cw0, op = w & 15u;
                if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) span += (int)(w >> 4);
                n_id += (op == 1u || op == 2u) ? 1u : 0u;
            }
        }
        const bool has_id = n_id != 0u;
        // the read's ordinal among the item's reads with an I/D op = its slot in the sparse quality index; counted
        // before any exit: the index lists every such read
        const uint32_t idm = __ballot_sync(0xffffffffu, gen && has_id);
        const uint32_t qord = c.ws->n_qord + __popc(idm & ((1u << lane) - 1u));
        __syncwarp();
        if (lane == 0 && idm) c.ws->n_qord += __popc(idm);
        bool work = gen && pos + span > c.ws->first;                     // fetched by range but not reaching the region: skipped
        if (work) { c.n_reads += 1u; c.n_bases += (uint32_t)L; }
        if (work && ((int64_t)pos + span > c.B.ref_len || pos < 0 || pos < c.col_begin || pos + span - c.col_begin >= c.n_cols)) {
            raise_error(c.totals, GA_ERR_OFFSET_RANGE, (uint32_t)(c.ws->begin + idx));
            work = false;
        }
        // staged tiles: lane = read; tiles that could not be staged (and very long CIGARs / reads): the group and whole-warp walks
        const bool by_lane = work && staged && n_ops <= (uint32_t)kLaneOps && L <= 256;
        if (__any_sync(0xffffffffu, by_lane))
            scan_generic_lanes(c, by_lane, i, idx, pos, L, span, m.c0, n_ops, n_id, cw0, reinterpret_cast<const uint32_t*>(c.ws->ring[b] + (by_lane ? m.so - sof : 0u)), qord, lane);
        uint32_t m_short = __ballot_sync(0xffffffffu, work && !by_lane && n_ops <= 8u), m_long = __ballot_sync(0xffffffffu, work && !by_lane && n_ops > 8u);
        while (m_short) {                                                // four reads per step, 8 lanes each
            uint32_t sel = 0u;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t s1 = m_short ? (uint32_t)(__ffs(m_short) - 1) : 0xffu;
                m_short &= m_short - 1u;
                sel |= s1 << (8 * g);
            }
            scan_generic_quad(c, sel, i, pos, L, m.c0, n_ops, m.so, qord, staged, b, sof, lane);
        }
        while (m_long) {
            const int src = __ffs(m_long) - 1; m_long &= m_long - 1;
            const int g_pos = __shfl_sync(0xffffffffu, pos, src), g_L = __shfl_sync(0xffffffffu, L, src);
            const uint32_t g_c0 = __shfl_sync(0xffffffffu, m.c0, src), g_c1 = g_c0 + __shfl_sync(0xffffffffu, n_ops, src);
            const uint32_t g_so = __shfl_sync(0xffffffffu, m.so, src), g_qord = __shfl_sync(0xffffffffu, qord, src);
            const uint32_t* g_rec = staged ? reinterpret_cast<const uint32_t*>(c.ws->ring[b] + (g_so - sof))
                                              : reinterpret_cast<const uint32_t*>(c.B.seq4 + 16ull * g_so);
            scan_generic_read(c, c.ws->i_base + t * TR + src, (int64_t)c.ws->begin + t * TR + src, g_pos, g_L, g_c0, g_c1, g_rec, g_qord, lane);
        }
    }
}

__global__ void __launch_bounds__(kScanThreads, 4) scan_kernel(BatchView B, SessView S, const SessionDesc* __restrict__ descs, ScanScratch X,
                                                               unsigned int* __restrict__ ticket, ga_totals* totals) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem* ws = reinterpret_cast<WarpSmem*>(smem_raw) + warp;
    if (lane == 0) {
        mbar_init(&ws->bar[0], 1u); mbar_init(&ws->bar[1], 1u);
        ws->wcnt = 0u; ws->ovf = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0u;                                                // bit b: phase the next wait on bar[b] looks for
    const bool tma_ok = (reinterpret_cast<uintptr_t>(B.seq4) & 15u) == 0u;
    const int64_t ref_words = ((B.ref_len + 7) >> 3) + 9;
    const uint32_t n_items = 2u * (uint32_t)S.n_sessions;

    ItemCtx c;
    c.B = B; c.ws = ws; c.totals = totals; c.X = X;
    uint32_t item = lane == 0 ? atomicAdd(ticket, 1u) : 0u;
    item = __shfl_sync(0xffffffffu, item, 0);
    while (item < n_items) {
        const uint32_t next = lane == 0 ? atomicAdd(ticket, 1u) : 0u;    // next item's ticket travels while this one is processed
        const int s = (int)(item >> 1);
        const uint32_t ds = item & 1u;
        const uint32_t dw = lane < 20 ? __ldg(reinterpret_cast<const uint32_t*>(descs + s) + lane) : 0u;
        c.ws->first = __ldg(S.first + s);
        const int t_begin = (int)__shfl_sync(0xffffffffu, dw, 0), t_end = (int)__shfl_sync(0xffffffffu, dw, 1);
        const int n_begin = (int)__shfl_sync(0xffffffffu, dw, 2), n_end = (int)__shfl_sync(0xffffffffu, dw, 3);
        c.col_begin = (int)__shfl_sync(0xffffffffu, dw, 4); c.n_cols = (int)__shfl_sync(0xffffffffu, dw, 5);
        const int big = (int)__shfl_sync(0xffffffffu, dw, 7);
        const uint32_t seq_n = __shfl_sync(0xffffffffu, dw, ds ? 15 : 13);
        c.ws->begin = ds ? n_begin : t_begin;
        c.ws->n = ds ? n_end - n_begin : t_end - t_begin;
        c.ws->i_base = ds ? t_end - t_begin : 0;
        c.ws->relbase = ((c.col_begin + 8) & ~7) - 8 * kSrefPad;
        c.table_in_ref = c.col_begin >= 0 && (int64_t)c.col_begin + c.n_cols <= B.ref_len;
        c.ws->item = item;
        c.n_reads = 0u; c.n_bases = 0u;
        if (lane == 0) { ws->n_ent = 0u; ws->n_obs = 0u; ws->n_qord = 0u; }
        __syncwarp();
        if (!big && c.ws->n > 0) {
            // ---- the session's reference window (+ record padding, + funnel-shift lookahead)
            {
                const int64_t w0 = (int64_t)((c.col_begin + 8) >> 3) - kSrefPad;
                const int nw = (c.n_cols >> 3) + 8 + kSrefPad;
                for (int k = lane; k < nw; k += 32) ws->sref[k] = (w0 + k >= 0 && w0 + k < ref_words) ? __ldg(B.ref4 + w0 + k) : 0xffffffffu;
            }
            const uint32_t avg_units = max(1u, (seq_n + (uint32_t)c.ws->n - 1u) / (uint32_t)c.ws->n);
            const int TR = (int)max(1u, min(31u, (uint32_t)kTileUnits / avg_units));     // reads per tile (31: see TileMeta)
            const int n_tiles = (c.ws->n + TR - 1) / TR;
            // ---- software pipeline: meta two tiles ahead, record bytes (TMA) one tile ahead
            TileMeta mA = load_tile_meta(c, 0, TR, lane);
            TileMeta mB = load_tile_meta(c, 1, TR, lane);
            __syncwarp();                                                 // the previous item's ring reads are done
            uint32_t stA = issue_tile(c, mA, lane < TR && lane < c.ws->n, 0, tma_ok, lane);
            uint32_t cwA = (meta_ops(mA) && lane < TR && lane < c.ws->n) ? __ldg(B.cigar + mA.c0) : 0u;
            uint32_t stB = 0u;
            if (n_tiles > 1) stB = issue_tile(c, mB, lane < TR && TR + lane < c.ws->n, 1, tma_ok, lane);
            for (int t = 0; t < n_tiles; ++t) {
                const int b = t & 1;
                uint32_t cwB = 0u;
                TileMeta mC = {0, 0u, 0u, 0u};
                if (stA) { mbar_wait(&ws->bar[b], (parity >> b) & 1u); parity ^= 1u << b; }
                scan_tile_w(c, t, TR, mA, cwA, stA != 0u, b, lane, [&]() {
                    const uint32_t opsB = meta_ops(mB);
                    if (opsB && lane < TR && (t + 1) * TR + lane < c.ws->n) {
                        cwB = __ldg(B.cigar + mB.c0);
                        if (opsB > 1u) prefetch_l1(B.cigar + mB.c0 + 1);
                    }
                    mC = load_tile_meta(c, t + 2, TR, lane);
                });
                __syncwarp();                                             // every lane is done with ring stage b
                uint32_t stC = 0u;
                if (t + 2 < n_tiles) stC = issue_tile(c, mC, lane < TR && (t + 2) * TR + lane < c.ws->n, b, tma_ok, lane);
                if (ws->wcnt >= 32u) flush_entries(c, lane);
                mA = mB; cwA = cwB; stA = stB;
                mB = mC; stB = stC;
            }
            flush_entries(c, lane);
        }
        if (!big) {
            const uint32_t nr = warp_sum(c.n_reads), nb = warp_sum(c.n_bases);
            __syncwarp();
            if (lane == 0) {
                X.cnt[item] = make_uint4(ws->ovf ? kCntOverflow : ws->n_ent, ws->n_obs, nr, nb);
                ws->ovf = 0u;
            }
            __syncwarp();
        }
        item = __shfl_sync(0xffffffffu, next, 0);
    }
}

}  // namespace ga
