// ga_emit_kernel.cuh - stage 3 of the streaming pipeline: grid-wide emission of the compacted modified records
// (north_star jobs (3) + (4)).
//
// The resolve kernel (ga_resolve_kernel.cuh) decides WHAT changes: it allocates the output slots, writes the record
// headers and hands over, per record, a 16-byte descriptor and, per session, the list of germline SNV alleles.
// These kernels write the record bodies with full-chip parallelism and no block-level synchronisation.
// emit_kernel: a warp takes 32 records at a time: kinds and descriptors arrive with two coalesced round trips, then
// a group of 8 lanes copies and patches each record of kind 1 with 128-bit loads / stores (four records per step).
// emit_special_kernel: the records of the other kinds, densely from the list the resolve kernels packed.
//   kind 1  clean read, SNV-only   copy; the (at most two) bases whose allele is germline arrive in the descriptor
//                                  together with the reference base that replaces them
//                                  (anonymizer_methods.py:170-176); qualities untouched
//   kind 4  same, any number of hits: every mismatch is looked up in the session's germline list
//   kind 2  other CIGAR, SNV-only  same, walking the CIGAR per 8-base word
//   kind 3  indel-masked           all DELs then all INSs at original offsets with the quality rules of
//                                  anonymizer_methods.py:178-203, 254-270 (emit_indel_group_t)
#pragma once
#include "ga_resolve_kernel.cuh"

namespace ga {

struct GermList {
    const uint32_t* e; uint32_t n;
    __device__ __forceinline__ bool operator()(int col, uint32_t b) const {
        const uint32_t key = ((uint32_t)col << 4) | b;
        for (uint32_t k = 0; k < n; ++k) if (__ldg(e + k) == key) return true;
        return false;
    }
};

__global__ void __launch_bounds__(kThreads) emit_kernel(BatchView B, const SessionDesc* __restrict__ descs, ResultView O, EmitScratch2 E) {
    __shared__ uint32_t stage[kThreads / kGroup][kGroupStage];
    const int tid = threadIdx.x, lane = tid & 31, group = tid / kGroup, glane = tid % kGroup, gw = (lane >> 3);
    if (*E.n_kind1 == 0u) return;                                       // the resolve kernels wrote every body themselves (the usual case)
    const unsigned long long n_all = O.totals->n_modified;
    const int64_t n = (int64_t)(n_all < (unsigned long long)O.cap_records ? n_all : (unsigned long long)O.cap_records);
    const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (tid >> 5);
    const int64_t stride = (int64_t)gridDim.x * (kThreads / 32) * 32;
    for (int64_t k0 = warp_global * 32; k0 < n; k0 += stride) {
        // ---- lane = record: kind, descriptor and output slot (two coalesced round trips)
        const int64_t k = k0 + lane;
        const uint32_t kind = k < n ? E.kind[k] : 0u;
        uint4 d = make_uint4(0u, 0u, 0u, 0u);
        uint32_t dst16 = 0u;
        if (kind == 1u) { d = E.edesc[k]; dst16 = O.mod_seq_off16[k]; }
        const uint32_t m1 = __ballot_sync(0xffffffffu, kind == 1u);
        // ---- kind 1: copy + patch, four records per step, 8 lanes each
#pragma unroll 2
        for (int step = 0; step < 8; ++step) {
            if (!((m1 >> (4 * step)) & 0xfu)) continue;
            const int rr = 4 * step + gw;
            const uint32_t r_kind = __shfl_sync(0xffffffffu, kind, rr);
            const uint32_t r_src = __shfl_sync(0xffffffffu, d.x, rr);
            const int rel0 = (int)__shfl_sync(0xffffffffu, d.y, rr);
            const uint32_t lz = __shfl_sync(0xffffffffu, d.z, rr);
            const uint32_t hits = __shfl_sync(0xffffffffu, d.w, rr);
            const uint32_t r_dst = __shfl_sync(0xffffffffu, dst16, rr);
            if (r_kind != 1u) continue;
            const int L = (int)(lz & 0xffffu), nh = (int)(lz >> 16);
            const uint4* rec = reinterpret_cast<const uint4*>(B.seq4 + 16ull * r_src);
            uint4* out = reinterpret_cast<uint4*>(O.out_seq4 + 16ull * r_dst);
            const int units = (L + 31) >> 5;
            const int q0 = (int)((hits >> 4) & 0xfffu) - rel0, q1 = nh > 1 ? (int)(hits >> 20) - rel0 : -1;
            const uint32_t f0 = hits & 15u, f1 = (hits >> 16) & 15u;
            for (int u = glane; u < units; u += kGroup) {
                uint4 v = ldg128(rec + u);
                if (32 * u + 32 > L) { v.x &= tail_mask(L, 4 * u); v.y &= tail_mask(L, 4 * u + 1); v.z &= tail_mask(L, 4 * u + 2); v.w &= tail_mask(L, 4 * u + 3); }
                if ((q0 >> 5) == u) {                                  // anonymizer_methods.py:170-176: base <- reference base
                    const uint32_t sh = (uint32_t)(q0 & 7) * 4u, keepm = ~(0xfu << sh), ins = f0 << sh;
                    const int w = (q0 >> 3) & 3;
                    if (w == 0) v.x = (v.x & keepm) | ins; else if (w == 1) v.y = (v.y & keepm) | ins;
                    else if (w == 2) v.z = (v.z & keepm) | ins; else v.w = (v.w & keepm) | ins;
                }
                if (q1 >= 0 && (q1 >> 5) == u) {
                    const uint32_t sh = (uint32_t)(q1 & 7) * 4u, keepm = ~(0xfu << sh), ins = f1 << sh;
                    const int w = (q1 >> 3) & 3;
                    if (w == 0) v.x = (v.x & keepm) | ins; else if (w == 1) v.y = (v.y & keepm) | ins;
                    else if (w == 2) v.z = (v.z & keepm) | ins; else v.w = (v.w & keepm) | ins;
                }
                out[u] = v;
            }
        }
    }
}

// ------------------------------------------------------------------ two edits as runs
// After its (at most two) edits a read is a handful of RUNS: pieces of the SNV-masked source and, per masked DEL, the
// deleted reference bases that come back (anonymizer_methods.py:188-195; all DELs, then all INSs, each at its original,
// clamped offset: :254-270).  The leader lane of a group builds the run list in shared memory by applying the edits to
// the list [whole read]; every output word is then merged from the runs it overlaps - no per-element loop.
constexpr int kMaxRuns = 6;
struct RunList { int f[kMaxRuns], len[kMaxRuns], a[kMaxRuns], kind[kMaxRuns]; int n; };   // kind 0: source from offset a; 1 + q: reference from position a, re-inserted by DEL q

__device__ __noinline__ void build_runs(const Ed2& E, int L, RunList* R) {
    int n = 1;
    R->len[0] = L; R->a[0] = 0; R->kind[0] = 0;
    for (int q = 0; q < E.ne; ++q) {
        const int p = E.p[q];
        if (q < E.n_del) {                                              // insert E.len[q] reference bases at p
            int k = 0, acc = 0;
            while (k < n && acc + R->len[k] < p) { acc += R->len[k]; ++k; }
            if (k >= n) { R->len[n] = E.len[q]; R->a[n] = E.pos[q]; R->kind[n] = 1 + q; ++n; continue; }
            const int off = p - acc;
            for (int t = n - 1; t > k; --t) { R->len[t + 2] = R->len[t]; R->a[t + 2] = R->a[t]; R->kind[t + 2] = R->kind[t]; }
            R->len[k + 2] = R->len[k] - off; R->a[k + 2] = R->a[k] + off; R->kind[k + 2] = R->kind[k];
            R->len[k + 1] = E.len[q]; R->a[k + 1] = E.pos[q]; R->kind[k + 1] = 1 + q;
            R->len[k] = off;
            n += 2;
        } else {                                                        // remove [p, e)
            const int e = E.e[q];
            int acc = 0;
            for (int k = 0; k < n; ++k) {
                const int ln = R->len[k], lo = max(acc, p), hi = min(acc + ln, e);
                if (lo < hi) {
                    if (lo == acc) { R->a[k] += hi - acc; R->len[k] = acc + ln - hi; }           // a prefix (or all) of the run goes
                    else if (hi == acc + ln) R->len[k] = lo - acc;                                // a suffix goes
                    else {                                                                         // the middle goes: the run splits
                        for (int t = n - 1; t > k; --t) { R->len[t + 1] = R->len[t]; R->a[t + 1] = R->a[t]; R->kind[t + 1] = R->kind[t]; }
                        R->len[k + 1] = acc + ln - hi; R->a[k + 1] = R->a[k] + (hi - acc); R->kind[k + 1] = R->kind[k];
                        R->len[k] = lo - acc;
                        ++n;
                        acc += ln;
                        ++k;                                                                       // the new right part lies behind e
                        continue;
                    }
                }
                acc += ln;
            }
        }
    }
    int f = 0;
    for (int k = 0; k < n; ++k) { R->f[k] = f; f += R->len[k]; }
    R->n = n;
}

// The body of an indel-masked record with two edits from the staged (SNV-masked) words and quality bytes.
__device__ __noinline__ void emit_runs_group(const BatchView& B, ga_totals* totals, const ResultView& O, bool act, const Ed2& E, int64_t r, int L, bool reverse,
                                             const uint8_t* qrec, const uint32_t* stage, const uint32_t* qstage, RunList* R, uint64_t seq16, uint64_t qual16,
                                             int new_len, int glane) {
    const int nqw = (L + 3) >> 2;
    if (act && !qrec) { if (glane == 0) raise_error(totals, GA_ERR_BAD_ARGUMENT, (uint32_t)r); act = false; }
    uint32_t part = 0;
    if (act) {
        for (int q = glane; q < nqw; q += kGroup) {
            uint32_t v = qstage[q];
            if (4 * q + 4 > L) v &= 0xffffffffu >> ((4 * q + 4 - L) * 8);
            part += __vsadu4(v, 0u);
        }
        if (glane == 0) {
            build_runs(E, L, R);
            for (int q = 0; q < E.n_del; ++q)
                if ((int64_t)E.pos[q] + E.len[q] > B.ref_len) raise_error(totals, GA_ERR_LENGTH_MISMATCH, (uint32_t)r);
        }
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1); part += __shfl_xor_sync(0xffffffffu, part, 2); part += __shfl_xor_sync(0xffffffffu, part, 4);
    __syncwarp();                                                         // the run list is visible to the group
    if (!act) return;
    // quality of re-inserted bases: floor(mean(current qualities)), recomputed after each DEL (anonymizer_methods.py:193)
    uint32_t mean0 = 0u, mean1 = 0u;
    {
        uint32_t sum = part, n = (uint32_t)L;
        if (E.n_del >= 1) { mean0 = n ? sum / n : 0u; sum += mean0 * (uint32_t)E.len[0]; n += (uint32_t)E.len[0]; }
        if (E.n_del >= 2) mean1 = n ? sum / n : 0u;
    }
    auto low_nibbles = [](int cnt) -> uint32_t { return cnt >= 8 ? 0xffffffffu : (cnt <= 0 ? 0u : (0xffffffffu >> ((8 - cnt) * 4))); };
    auto low_bytes = [](int cnt) -> uint32_t { return cnt >= 4 ? 0xffffffffu : (cnt <= 0 ? 0u : (0xffffffffu >> ((4 - cnt) * 8))); };
    const int n_runs = R->n;
    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
    for (int w = glane; w < units * 4; w += kGroup) {
        const int j0 = w << 3;
        uint32_t v = 0u;
        if (j0 < new_len) {
#pragma unroll 1
            for (int k = 0; k < n_runs; ++k) {
                const int f = R->f[k], lo = max(f, j0), hi = min(f + R->len[k], j0 + 8);
                if (lo >= hi) continue;
                const uint32_t m = low_nibbles(hi - j0) & ~low_nibbles(lo - j0);
                const int at = R->a[k] + (j0 - f);                       // >= -7
                uint32_t val;
                if (R->kind[k]) val = ref_word(B.ref4, (int64_t)at);
                else if (at < 0) val = stage[0] << ((-at) * 4);
                else val = __funnelshift_r(stage[at >> 3], stage[(at >> 3) + 1], (uint32_t)(at & 7) * 4u);
                v |= val & m;
            }
        }
        oseq[w] = v;
    }
    // qualities in printed (= BAM) order; the runs index the forward-orientation array (anonymizer_methods.py:95, 187,
    // 195: quirk Q2), so for a reverse read a run [f, f + len) is printed at [new_len - f - len, new_len - f) and its BAM
    // bytes ascend with the printed index
    uint32_t* oq = reinterpret_cast<uint32_t*>(O.out_qual + 32ull * qual16);
    for (int w = glane; w < units * 8; w += kGroup) {
        const int p0 = w << 2;
        uint32_t v = 0u;
        if (p0 < new_len) {
#pragma unroll 1
            for (int k = 0; k < n_runs; ++k) {
                const int ln = R->len[k], f = reverse ? new_len - R->f[k] - ln : R->f[k];
                const int lo = max(f, p0), hi = min(f + ln, p0 + 4);
                if (lo >= hi) continue;
                const uint32_t m = low_bytes(hi - p0) & ~low_bytes(lo - p0);
                uint32_t val;
                const int kind = R->kind[k];
                if (kind) val = (kind == 1 ? mean0 : mean1) * 0x01010101u;
                else {
                    const int at = p0 + (reverse ? L - new_len - R->a[k] + R->f[k] : R->a[k] - R->f[k]);   // BAM byte of printed byte p0 (>= -3)
                    val = at < 0 ? qstage[0] << ((-at) * 8) : __funnelshift_r(qstage[at >> 2], qstage[(at >> 2) + 1], (uint32_t)(at & 3) * 8u);
                }
                v |= val & m;
            }
        }
        oq[w] = v;
    }
}

// Reads with other CIGARs, the common shapes: SNV-only (kind 2, E.ne == 0) and one germline indel (kind 3,
// E.ne == 1), reads of at most 8 * (kGroupStage - 1) bases.  One group of 8 lanes per record; every lane of the warp
// calls this (act = false for groups without such a record).
//   1. the record is staged in shared memory, coalesced;
//   2. SNV masking (anonymizer_methods.py:170-176): every germline allele of the session is carried through the
//      CIGAR to its query offset and, when the read shows it, replaced by the reference base;
//   3. the one edit is applied while copying: DEL -> the deleted reference bases come back with quality
//      floor(mean(qualities)), INS -> the inserted bases and qualities go (anonymizer_methods.py:178-203); edits
//      index the forward-orientation quality array, printed order = BAM order (anonymizer_methods.py:95, 213).
__device__ __forceinline__ void emit_special_group(const BatchView& B, ga_totals* totals, const ResultView& O, bool act, const Ed2& E, int64_t r,
                                                   int pos, int L, uint32_t src_unit, uint32_t c0, uint32_t c1, bool reverse, int col_begin, const GermList& germ, const uint8_t* qrec,
                                                   uint32_t* stage, uint32_t* qstage, uint64_t seq16, uint64_t qual16, int new_len, int glane) {
    const int nw = (L + 7) >> 3, nqw = (L + 3) >> 2;
    if (act) {
        const uint32_t* rec = reinterpret_cast<const uint32_t*>(B.seq4 + 16ull * src_unit);
        for (int w = glane; w < nw; w += kGroup) stage[w] = __ldg(rec + w) & tail_mask(L, w);
        if (glane == 0) stage[nw] = 0u;                                // the funnel shift may touch one word past the end
        if (E.ne >= 1 && qrec) {                                       // the quality record too: it is summed and shifted below
            const uint32_t* qg = reinterpret_cast<const uint32_t*>(qrec);
            for (int q = glane; q < nqw; q += kGroup) qstage[q] = __ldg(qg + q);
            if (glane == 0) qstage[nqw] = 0u;
        }
    }
    __syncwarp();
    if (act) {
        for (uint32_t a = glane; a < germ.n; a += kGroup) {
            const uint32_t key = __ldg(germ.e + a), code = key & 15u;
            const int at = col_begin + (int)(key >> 4);
            int rc = pos, q = 0;
            for (uint32_t ci = c0; ci < c1; ++ci) {
                const uint32_t cw = __ldg(B.cigar + ci), op = cw & 15u;
                const int ln = (int)(cw >> 4);
                if (at < rc) break;                                    // the column lies before what is left of the read
                if (op == 0u || op == 7u || op == 8u) {
                    if (at < rc + ln) {
                        const int qq = q + (at - rc);
                        if (qq < L && ((stage[qq >> 3] >> ((qq & 7) * 4)) & 15u) == code)
                            atomicXor(&stage[qq >> 3], (code ^ ref_code(B.ref4, at)) << ((qq & 7) * 4));
                        break;
                    }
                    q += ln; rc += ln;
                } else if (op == 1u || op == 4u) q += ln;
                else if (op == 2u || op == 3u) { if (at < rc + ln) break; rc += ln; }
            }
        }
    }
    __syncwarp();
    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
    const bool indel = act && E.ne == 1;                              // two edits: the caller continues with emit_runs_group
    if (act && E.ne == 0) for (int w = glane; w < units * 4; w += kGroup) oseq[w] = w < nw ? stage[w] : 0u;
    // ---- one edit
    const bool is_del = E.n_del == 1;
    const int p = E.p[0], len = E.len[0], shift = is_del ? -len : E.e[0] - E.p[0];      // source = final + shift behind the edit
    const int ins_end = is_del ? p + len : p;                                             // [p, ins_end): re-inserted elements
    bool ok = indel;
    if (indel && !qrec) { if (glane == 0) raise_error(totals, GA_ERR_BAD_ARGUMENT, (uint32_t)r); ok = false; }
    uint32_t mean = 0u;
    if (__any_sync(0xffffffffu, ok && is_del)) {                      // quality of re-inserted bases (anonymizer_methods.py:193)
        uint32_t part = 0;
        if (ok && is_del) {
            for (int q = glane; q < nqw; q += kGroup) {
                uint32_t v = qstage[q];
                if (4 * q + 4 > L) v &= 0xffffffffu >> ((4 * q + 4 - L) * 8);
                part += __vsadu4(v, 0u);
            }
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1); part += __shfl_xor_sync(0xffffffffu, part, 2); part += __shfl_xor_sync(0xffffffffu, part, 4);
        mean = L ? part / (uint32_t)L : 0u;
        if (ok && is_del && glane == 0 && (int64_t)E.pos[0] + len > B.ref_len) raise_error(totals, GA_ERR_LENGTH_MISMATCH, (uint32_t)r);
    }
    if (!ok) return;
    // The final array is three pieces: [0, p) = source as is, [p, ins_end) = re-inserted reference bases (DEL only),
    // [ins_end, new_len) = source shifted by `shift`.  Every output word is merged from the (at most three) pieces it
    // overlaps with nibble / byte masks - no per-element loop.
    auto low_nibbles = [](int cnt) -> uint32_t { return cnt >= 8 ? 0xffffffffu : (cnt <= 0 ? 0u : (0xffffffffu >> ((8 - cnt) * 4))); };
    auto low_bytes = [](int cnt) -> uint32_t { return cnt >= 4 ? 0xffffffffu : (cnt <= 0 ? 0u : (0xffffffffu >> ((4 - cnt) * 8))); };
    auto stage_at = [&](int sidx) -> uint32_t {                        // 8 staged bases from source index sidx (>= -7; nibbles outside the read are masked by the caller)
        if (sidx < 0) return stage[0] << ((-sidx) * 4);
        return __funnelshift_r(stage[sidx >> 3], stage[(sidx >> 3) + 1], (uint32_t)(sidx & 7) * 4u);
    };
    for (int w = glane; w < units * 4; w += kGroup) {
        const int j0 = w << 3;
        uint32_t v = 0u;
        if (j0 + 8 <= p) v = stage_at(j0);                               // wholly in front of the edit (the usual word)
        else if (j0 >= ins_end && j0 + 8 <= new_len) v = stage_at(j0 + shift);   // wholly behind it
        else if (j0 < new_len) {
            const uint32_t mA = low_nibbles(p - j0), mAB = low_nibbles(ins_end - j0);
            if (mA) v = stage_at(j0) & mA;
            if (mAB & ~mA) v |= ref_word(B.ref4, (int64_t)E.pos[0] + (j0 - p)) & (mAB & ~mA);
            if (~mAB) v |= stage_at(j0 + shift) & ~mAB;
            v &= low_nibbles(new_len - j0);
        }
        oseq[w] = v;
    }
    // qualities in printed (= BAM) order.  The edit indexes the forward-orientation array (anonymizer_methods.py:95,
    // 187, 195: quirk Q2), so for a reverse read the pieces come in the opposite order: printed [0, b1) = BAM bytes as
    // they are, [b1, b2) = the mean, [b2, new_len) = BAM bytes shifted by d3.
    const int b1 = reverse ? new_len - ins_end : p, b2 = reverse ? new_len - p : ins_end, d3 = reverse ? L - new_len : shift;
    auto qual_at = [&](int bidx) -> uint32_t {                         // 4 staged quality bytes from BAM byte bidx (>= -3)
        if (bidx < 0) return qstage[0] << ((-bidx) * 8);
        return __funnelshift_r(qstage[bidx >> 2], qstage[(bidx >> 2) + 1], (uint32_t)(bidx & 3) * 8u);
    };
    uint32_t* oq = reinterpret_cast<uint32_t*>(O.out_qual + 32ull * qual16);
    for (int w = glane; w < units * 8; w += kGroup) {
        const int p0 = w << 2;
        uint32_t v = 0u;
        if (p0 + 4 <= b1) v = qual_at(p0);                               // wholly in front of the edit (the usual word)
        else if (p0 >= b2 && p0 + 4 <= new_len) v = qual_at(p0 + d3);    // wholly behind it
        else if (p0 < new_len) {
            const uint32_t m1 = low_bytes(b1 - p0), m12 = low_bytes(b2 - p0);
            if (m1) v = qual_at(p0) & m1;
            if (m12 & ~m1) v |= (mean * 0x01010101u) & (m12 & ~m1);
            if (~m12) v |= qual_at(p0 + d3) & ~m12;
            v &= low_bytes(new_len - p0);
        }
        oq[w] = v;
    }
}

// Records of kind >= 2 (reads with other CIGARs, reads with many hits), taken densely from the list the resolve
// kernels packed: four records per warp step, one group of 8 lanes each.
__global__ void __launch_bounds__(kThreads) emit_special_kernel(BatchView B, const SessionDesc* __restrict__ descs, ResultView O, EmitScratch2 E) {
    __shared__ uint32_t stage[kThreads / kGroup][kGroupStage];
    __shared__ uint32_t qstage[kThreads / kGroup][2 * kGroupStage];      // quality bytes of the same reads (4 per word)
    __shared__ RunList runs[kThreads / kGroup];                           // two-edit records: the runs the output is merged from
    const int tid = threadIdx.x, group = tid / kGroup, glane = tid % kGroup;
    const uint32_t n_x = (uint32_t)min((int64_t)*E.n_special, O.cap_records);   // slots past the capacity were never written
    const uint32_t groups_total = gridDim.x * (kThreads / kGroup);
    for (uint32_t jb = blockIdx.x * (kThreads / kGroup) + (tid >> 5) * 4; jb < n_x; jb += groups_total) {   // warp-uniform
        const uint32_t j = jb + ((tid & 31) >> 3);
        const bool have = j < n_x;
        // ---- the record's 64-byte descriptor: one round trip, one 16-byte part per lane of the group
        uint4 part = make_uint4(0u, 0u, 0u, 0u);
        if (have && glane < 4) part = E.sdesc[4ull * j + glane];
        const int gbase = (tid & 31) & ~7;
        uint4 d0, d1, d2, d3;
        d0.x = __shfl_sync(0xffffffffu, part.x, gbase);     d0.y = __shfl_sync(0xffffffffu, part.y, gbase);
        d0.z = __shfl_sync(0xffffffffu, part.z, gbase);     d0.w = __shfl_sync(0xffffffffu, part.w, gbase);
        d1.x = __shfl_sync(0xffffffffu, part.x, gbase + 1); d1.y = __shfl_sync(0xffffffffu, part.y, gbase + 1);
        d1.z = __shfl_sync(0xffffffffu, part.z, gbase + 1); d1.w = __shfl_sync(0xffffffffu, part.w, gbase + 1);
        d2.x = __shfl_sync(0xffffffffu, part.x, gbase + 2); d2.y = __shfl_sync(0xffffffffu, part.y, gbase + 2);
        d2.z = __shfl_sync(0xffffffffu, part.z, gbase + 2);
        d3.x = __shfl_sync(0xffffffffu, part.x, gbase + 3); d3.y = __shfl_sync(0xffffffffu, part.y, gbase + 3);
        const uint32_t r_kind = (d0.z >> 16) & 15u;                      // 0: the slot of a session that did not fit
        const uint32_t r_src = d0.x; const int r_pos = (int)d0.y; const int s = (int)d0.w;
        const int64_t r = (int64_t)d1.x; const int new_len = (int)d1.y; const uint64_t seq16 = d1.z, qual16 = d1.w;
        const uint32_t c0 = d2.x, c1 = d2.y; const int col_begin = (int)d2.z;
        const bool reverse = ((d0.z >> 20) & 1u) != 0u;
        GermList germ; germ.e = E.germ + (size_t)s * kGermStride + 4; germ.n = 0u;
        if (r_kind) germ.n = __ldg(E.germ + (size_t)s * kGermStride);
        {
                if (r_kind == 4u) {
                    const int L = new_len;
                    const uint4* rec = reinterpret_cast<const uint4*>(B.seq4 + 16ull * r_src);
                    uint4* out = reinterpret_cast<uint4*>(O.out_seq4 + 16ull * seq16);
                    const int units = (L + 31) >> 5;
                    for (int u = glane; u < units; u += kGroup) {
                        const uint4 vv = ldg128(rec + u);
                        const int64_t ni = (int64_t)r_pos + 32 * u + 8;
                        const uint32_t* rp = B.ref4 + (ni >> 3);
                        const uint32_t sh = (uint32_t)(ni & 7) * 4u;
                        const uint32_t r0 = __ldg(rp), r1 = __ldg(rp + 1), r2 = __ldg(rp + 2), r3 = __ldg(rp + 3), r4 = __ldg(rp + 4);
                        uint32_t w[4] = {vv.x, vv.y, vv.z, vv.w};
                        const uint32_t f[4] = {__funnelshift_r(r0, r1, sh), __funnelshift_r(r1, r2, sh), __funnelshift_r(r2, r3, sh), __funnelshift_r(r3, r4, sh)};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int wi = 4 * u + q;
                            const uint32_t tm = tail_mask(L, wi);
                            uint32_t v = w[q] & tm;
                            uint32_t x = (v ^ f[q]) & tm;
                            while (x) {
                                const int nb = (__ffs(x) - 1) >> 2;
                                x &= ~(0xfu << (nb * 4));
                                const uint32_t b = (v >> (nb * 4)) & 15u;
                                if (b != 15u && germ(r_pos + 8 * wi + nb - col_begin, b)) v = (v & ~(0xfu << (nb * 4))) | (((f[q] >> (nb * 4)) & 15u) << (nb * 4));
                            }
                            w[q] = v;
                        }
                        out[u] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                const bool indel = r_kind == 3u;
                Ed2 Ed; Ed.ne = 0; Ed.n_del = 0;
#pragma unroll
                for (int q = 0; q < 2; ++q) { Ed.irp[q] = 0; Ed.len[q] = 0; Ed.pos[q] = 0; Ed.mean[q] = 0u; Ed.p[q] = 0; Ed.e[q] = 0; }
                int64_t q_lo = 0, q_hi = 0;
                const uint8_t* qrec = nullptr;
                const int L = (int)(d0.z & 0xffffu);
                if (indel) {
                    const uint4* ap = reinterpret_cast<const uint4*>(O.out_qual + 32ull * qual16);
                    const uint4 x0 = ap[0], x1 = ap[1];              // EditAux written by the resolve kernel
                    Ed.irp[0] = (int)x0.x; Ed.pos[0] = (int)x0.y; Ed.len[0] = (int)(x0.z & 0x7fffffffu);
                    Ed.irp[1] = (int)x0.w; Ed.pos[1] = (int)x1.x; Ed.len[1] = (int)(x1.y & 0x7fffffffu);
                    Ed.ne = (int)(x1.z & 0xffu); Ed.n_del = (int)((x1.z >> 8) & 0xffu);
                    clamp_edits2(Ed, L);
                    q_lo = (int64_t)d3.x; q_hi = (int64_t)d3.y;
                    // the read's quality record: dense upload, or the slot the resolve kernel predicted in the sparse
                    // index (verified; a caller may list more reads than those with I/D ops), or a search of the slice
                    if (B.qual && !B.qual_reads) qrec = B.qual + 32ull * r_src;
                    else if (B.qual) {
                        const int64_t qi = (int64_t)x1.w;
                        if (qi < B.n_qual && __ldg(B.qual_reads + qi) == (int32_t)r) qrec = B.qual + 32ull * __ldg(B.qual_off16 + qi);
                        else qrec = qual_record_in(B, r, q_lo, q_hi);
                    }
                }
                __syncwarp();                                         // every lane of the group has read the aux before it is overwritten
                // the common shapes take the staged path; two edits or very long reads take the general one
                const bool fast = (r_kind == 2u || (indel && (Ed.ne == 1 || Ed.ne == 2))) && ((L + 7) >> 3) <= kGroupStage - 1;
                if (__any_sync(0xffffffffu, fast))
                    emit_special_group(B, O.totals, O, fast, Ed, r, r_pos, L, r_src, c0, c1, reverse, col_begin, germ, qrec, stage[group], qstage[group], seq16, qual16, new_len, glane);
                const bool two = fast && indel && Ed.ne == 2;
                if (__any_sync(0xffffffffu, two)) {
                    __syncwarp();                                     // staged words are SNV-masked
                    emit_runs_group(B, O.totals, O, two, Ed, r, L, reverse, qrec, stage[group], qstage[group], &runs[group], seq16, qual16, new_len, glane);
                }
                if (r_kind == 2u && !fast) {
                    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
                    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
                    masked_words_g(B, r, r_pos, new_len, c0, c1, col_begin, units * 4, glane, kGroup, germ, [&](int wd, uint32_t v) { oseq[wd] = v; });
                }
                const bool slow = indel && !fast;
                if (__any_sync(0xffffffffu, slow)) {
                    __syncwarp();
                    emit_indel_group_t(B, O.totals, O, slow, Ed, r, col_begin, q_lo, q_hi, stage[group], seq16, qual16, new_len, glane, germ);
                }
            }
        }
    }
// Records of kind 5: indel-masked reads with more than two germline indels (rare; the one-CTA resolve kernel lists
// them and stores their edit lists in E.many).  One warp per record, general element-wise emission.
__global__ void __launch_bounds__(kThreads) emit_many_kernel(BatchView B, ResultView O, EmitScratch2 E) {
    const int lane = threadIdx.x & 31;
    const uint32_t n_x = min(*E.n_many_recs, E.cap_many);
    const uint32_t warps_total = gridDim.x * (kThreads / 32);
    for (uint32_t j = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); j < n_x; j += warps_total) {   // warp-uniform
        const uint4* dp = E.sdesc + 4ull * E.many_recs[j];
        const uint4 d0 = dp[0], d1 = dp[1], d2 = dp[2], d3 = dp[3];
        const int s = (int)d0.w;
        const int64_t r = (int64_t)d1.x;
        const uint64_t qual16 = d1.w;
        GermList germ; germ.e = E.germ + (size_t)s * kGermStride + 4; germ.n = __ldg(E.germ + (size_t)s * kGermStride);
        const uint4* ap = reinterpret_cast<const uint4*>(O.out_qual + 32ull * qual16);   // write_many_aux
        const uint4 x0 = ap[0], x1 = ap[1];
        const uint8_t* qrec = nullptr;
        if (B.qual && !B.qual_reads) qrec = B.qual + 32ull * d0.x;
        else if (B.qual) {
            const int64_t qi = (int64_t)x1.w;
            if (qi < B.n_qual && __ldg(B.qual_reads + qi) == (int32_t)r) qrec = B.qual + 32ull * __ldg(B.qual_off16 + qi);
            else qrec = qual_record_in(B, r, (int64_t)d3.x, (int64_t)d3.y);
        }
        __syncwarp();                                                 // every lane has read the aux before it is overwritten
        emit_many_group<32>(B, O.totals, O, true, E.many + x0.x, (int)(x1.z & 0xffu), (int)((x1.z >> 8) & 0xffu), r, (int)d0.y, (int)(d0.z & 0xffffu), d0.x,
                            d2.x, d2.y, ((d0.z >> 20) & 1u) != 0u, (int)d2.z, qrec, (uint64_t)d1.z, qual16, (int)d1.y, lane, germ);
    }
}

}  // namespace ga
