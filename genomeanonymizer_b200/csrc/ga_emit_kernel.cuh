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
    const int tid = threadIdx.x, lane = tid & 31, glane = tid % kGroup, gw = (lane >> 3);
    const int64_t n = (int64_t)min((unsigned long long)*E.n_kind1, (unsigned long long)O.cap_records);   // 0: the resolve kernels wrote every body themselves (the usual case)
    const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (tid >> 5);
    const int64_t stride = (int64_t)gridDim.x * (kThreads / 32) * 32;
    for (int64_t k0 = warp_global * 32; k0 < n; k0 += stride) {
        // ---- lane = record: its index from the list, descriptor and output slot (two coalesced round trips)
        const int64_t k = k0 + lane;
        const uint32_t kind = k < n ? 1u : 0u;
        uint4 d = make_uint4(0u, 0u, 0u, 0u);
        uint32_t dst16 = 0u;
        if (kind == 1u) { const uint32_t rec = E.kind1_list[k]; d = E.edesc[rec]; dst16 = O.mod_seq_off16[rec]; }
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

// ------------------------------------------------------------------ staged records of a lane group
// A group of 8 lanes stages its record (SNV-masked below) and, for an indel-masked read, the quality record in shared
// memory.  Data starts at word 4 (16-byte aligned for 128-bit stores); word 3 is a zero pad so that source index -1 .. -8
// is addressable, and four words behind the data are addressable too: every piece of an output word is fetched without
// a branch and cut to size by masks.
constexpr int kStageW = 4 + kGroupStage + 4;
constexpr int kQStageW = 4 + 2 * kGroupStage + 4;

__device__ __forceinline__ uint32_t low_nibbles_bf(int cnt) {            // the low min(max(cnt, 0), 8) nibbles
    return __funnelshift_rc(0xffffffffu, 0u, (uint32_t)(32 - 4 * min(max(cnt, 0), 8)));
}
__device__ __forceinline__ uint32_t low_bytes_bf(int cnt) {              // the low min(max(cnt, 0), 4) bytes
    return __funnelshift_rc(0xffffffffu, 0u, (uint32_t)(32 - 8 * min(max(cnt, 0), 4)));
}
// 8 staged bases from source index sidx (any value: clamped to the addressable range, the caller masks what lies outside the read)
__device__ __forceinline__ uint32_t stage_at(const uint32_t* sg, int sidx) {
    const int s = min(max(sidx, -8), 8 * kGroupStage + 8);
    const int wi = (s >> 3) + 4;
    return __funnelshift_r(sg[wi], sg[wi + 1], (uint32_t)(s & 7) * 4u);
}
// 4 staged quality bytes from BAM byte bidx (same rules)
__device__ __forceinline__ uint32_t qual_at(const uint32_t* qs, int bidx) {
    const int s = min(max(bidx, -4), 8 * kGroupStage + 4);
    const int wi = (s >> 2) + 4;
    return __funnelshift_r(qs[wi], qs[wi + 1], (uint32_t)(s & 3) * 8u);
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

// What a group knows about its record (read from the warp's descriptor table, see emit_special_kernel).
struct SpecRec {
    uint32_t src_unit, c0, c1, germ_n, qunit;
    int pos, L, s, new_len, col_begin;
    int64_t r;
    uint64_t seq16, qual16;
    bool reverse;
};

// Steps shared by the staged kinds: (1) the record - and for an indel-masked read its quality record - goes to shared
// memory with 128-bit loads; (2) SNV masking (anonymizer_methods.py:170-176): every germline allele of the session is
// carried through the CIGAR to its query offset and, when the read shows it, replaced by the reference base.  Returns
// the sum of the read's qualities (kQual) to every lane of the group.
template <bool kQual>
__device__ __forceinline__ uint32_t stage_and_mask(const BatchView& B, const EmitScratch2& E, bool act, const SpecRec& R, uint32_t* sg, uint32_t* qs, int glane) {
    uint32_t part = 0u;
    if (act) {
        const int L = R.L;
        const uint4* rec = reinterpret_cast<const uint4*>(B.seq4 + 16ull * R.src_unit);
        uint4* sg4 = reinterpret_cast<uint4*>(sg) + 1;
        const int nu = (L + 31) >> 5;                                    // at most kGroupStage / 4 = 8 units
        if (glane < nu) {
            uint4 v = ldg128(rec + glane);
            v.x &= tail_mask(L, 4 * glane); v.y &= tail_mask(L, 4 * glane + 1); v.z &= tail_mask(L, 4 * glane + 2); v.w &= tail_mask(L, 4 * glane + 3);
            sg4[glane] = v;
        }
        if (glane == 7) { sg[3] = 0u; sg[4 + 4 * nu] = 0u; }            // pad in front; the word behind the last unit
        if (kQual) {
            const uint4* qrec = reinterpret_cast<const uint4*>(B.qual + 32ull * R.qunit);
            uint4* qs4 = reinterpret_cast<uint4*>(qs) + 1;
            const int nq = (L + 15) >> 4;                                // at most 16
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int u = glane + 8 * h;
                if (u < nq) {
                    uint4 v = ldg128(qrec + u);
                    v.x &= low_bytes_bf(L - 16 * u); v.y &= low_bytes_bf(L - 16 * u - 4); v.z &= low_bytes_bf(L - 16 * u - 8); v.w &= low_bytes_bf(L - 16 * u - 12);
                    part += __vsadu4(v.x, 0u) + __vsadu4(v.y, 0u) + __vsadu4(v.z, 0u) + __vsadu4(v.w, 0u);
                    qs4[u] = v;
                }
            }
            if (glane == 7) { qs[3] = 0u; qs[4 + 4 * nq] = 0u; }
        }
    }
    if (kQual) { part += __shfl_xor_sync(0xffffffffu, part, 1); part += __shfl_xor_sync(0xffffffffu, part, 2); part += __shfl_xor_sync(0xffffffffu, part, 4); }
    __syncwarp();
    if (act) {
        const uint32_t* ge = E.germ + (size_t)R.s * kGermStride + 4;
        for (uint32_t a = glane; a < R.germ_n; a += kGroup) {
            const uint32_t key = __ldg(ge + a), code = key & 15u;
            const int at = R.col_begin + (int)(key >> 4);
            int rc = R.pos, q = 0;
            for (uint32_t ci = R.c0; ci < R.c1; ++ci) {
                if (at < rc) break;                                    // the column lies before what is left of the read
                const uint32_t cw = __ldg(B.cigar + ci), op = cw & 15u;
                const int ln = (int)(cw >> 4);
                if (op == 0u || op == 7u || op == 8u) {
                    if (at < rc + ln) {
                        const int qq = q + (at - rc);
                        if (qq < R.L && ((sg[4 + (qq >> 3)] >> ((qq & 7) * 4)) & 15u) == code)
                            atomicXor(&sg[4 + (qq >> 3)], (code ^ ref_code(B.ref4, at)) << ((qq & 7) * 4));
                        break;
                    }
                    q += ln; rc += ln;
                } else if (op == 1u || op == 4u) q += ln;
                else if (op == 2u || op == 3u) { if (at < rc + ln) break; rc += ln; }
            }
        }
    }
    __syncwarp();
    return part;
}

// kind 2: another CIGAR, SNV-only - the staged, masked record is the output.
__device__ __forceinline__ void emit_snv_only_quad(const BatchView& B, const ResultView& O, const EmitScratch2& E, bool act, const SpecRec& R, uint32_t* sg, int glane) {
    stage_and_mask<false>(B, E, act, R, sg, nullptr, glane);
    if (!act) return;
    const int nu = (R.L + 31) >> 5;
    int units = (R.new_len + 31) >> 5; if (units < 1) units = 1;
    uint4* out = reinterpret_cast<uint4*>(O.out_seq4 + 16ull * R.seq16);
    const uint4* sg4 = reinterpret_cast<const uint4*>(sg) + 1;
    for (int u = glane; u < units; u += kGroup) out[u] = u < nu ? sg4[u] : make_uint4(0u, 0u, 0u, 0u);
}

// kind 3, one germline indel: DEL -> the deleted reference bases come back with quality floor(mean(qualities)), INS ->
// the inserted bases and qualities go (anonymizer_methods.py:178-203); edits index the forward-orientation quality
// array, printed order = BAM order (anonymizer_methods.py:95, 213).  The final array is three pieces: [0, p) = source as
// is, [p, ins_end) = re-inserted reference bases (DEL only), [ins_end, new_len) = source shifted by `shift`; every output
// word is merged from the pieces with nibble / byte masks, without a branch (the lanes of a warp sit in different pieces).
__device__ __forceinline__ void emit_one_edit_quad(const BatchView& B, const ResultView& O, const EmitScratch2& E, bool act, const SpecRec& R, const Ed2& Ed,
                                                   uint32_t* sg, uint32_t* qs, int glane) {
    const uint32_t sum = stage_and_mask<true>(B, E, act, R, sg, qs, glane);
    if (!act) return;
    const int L = R.L, new_len = R.new_len;
    const bool is_del = Ed.n_del == 1;
    const int p = Ed.p[0], len = Ed.len[0], shift = is_del ? -len : Ed.e[0] - Ed.p[0];       // source = final + shift behind the edit
    const int ins_end = is_del ? p + len : p;                                                 // [p, ins_end): re-inserted elements
    const uint32_t mean4 = (is_del && L ? sum / (uint32_t)L : 0u) * 0x01010101u;              // anonymizer_methods.py:193
    if (is_del && glane == 0 && (int64_t)Ed.pos[0] + len > B.ref_len) raise_error(O.totals, GA_ERR_LENGTH_MISMATCH, (uint32_t)R.r);
    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * R.seq16);
    for (int w = glane; w < units * 4; w += kGroup) {
        const int j0 = w << 3;
        const uint32_t mA = low_nibbles_bf(p - j0), mAB = low_nibbles_bf(ins_end - j0);
        uint32_t v = (stage_at(sg, j0) & mA) | (stage_at(sg, j0 + shift) & ~mAB);
        const uint32_t mR = mAB & ~mA;
        if (mR) v |= ref_word(B.ref4, (int64_t)Ed.pos[0] + (j0 - p)) & mR;
        oseq[w] = v & low_nibbles_bf(new_len - j0);
    }
    // qualities in printed (= BAM) order.  The edit indexes the forward-orientation array (quirk Q2), so for a reverse
    // read the pieces come in the opposite order: printed [0, b1) = BAM bytes as they are, [b1, b2) = the mean,
    // [b2, new_len) = BAM bytes shifted by d3.
    const int b1 = R.reverse ? new_len - ins_end : p, b2 = R.reverse ? new_len - p : ins_end, d3 = R.reverse ? L - new_len : shift;
    uint32_t* oq = reinterpret_cast<uint32_t*>(O.out_qual + 32ull * R.qual16);
    for (int w = glane; w < units * 8; w += kGroup) {
        const int p0 = w << 2;
        const uint32_t m1 = low_bytes_bf(b1 - p0), m12 = low_bytes_bf(b2 - p0);
        const uint32_t v = (qual_at(qs, p0) & m1) | (mean4 & m12 & ~m1) | (qual_at(qs, p0 + d3) & ~m12);
        oq[w] = v & low_bytes_bf(new_len - p0);
    }
}

// kind 3, two germline indels: the body from the staged (SNV-masked) words and quality bytes through the run list.
__device__ __forceinline__ void emit_two_edit_quad(const BatchView& B, const ResultView& O, const EmitScratch2& E, bool act, const SpecRec& R, const Ed2& Ed,
                                                   uint32_t* sg, uint32_t* qs, RunList* RL, int glane) {
    const uint32_t sum0 = stage_and_mask<true>(B, E, act, R, sg, qs, glane);
    if (act && glane == 0) {
        build_runs(Ed, R.L, RL);
        for (int q = 0; q < Ed.n_del; ++q)
            if ((int64_t)Ed.pos[q] + Ed.len[q] > B.ref_len) raise_error(O.totals, GA_ERR_LENGTH_MISMATCH, (uint32_t)R.r);
    }
    __syncwarp();                                                         // the run list is visible to the group
    if (!act) return;
    const int L = R.L, new_len = R.new_len;
    // quality of re-inserted bases: floor(mean(current qualities)), recomputed after each DEL (anonymizer_methods.py:193)
    uint32_t mean0 = 0u, mean1 = 0u;
    {
        uint32_t sum = sum0, n = (uint32_t)L;
        if (Ed.n_del >= 1) { mean0 = n ? sum / n : 0u; sum += mean0 * (uint32_t)Ed.len[0]; n += (uint32_t)Ed.len[0]; }
        if (Ed.n_del >= 2) mean1 = n ? sum / n : 0u;
    }
    const int n_runs = RL->n;
    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * R.seq16);
    // a lane's words ascend, so it walks the run list once: k = the first run that reaches beyond the word's start
    int k = 0;
    for (int w = glane; w < units * 4; w += kGroup) {
        const int j0 = w << 3;
        uint32_t v = 0u;
        while (k < n_runs - 1 && RL->f[k] + RL->len[k] <= j0) ++k;
#pragma unroll 1
        for (int kk = k;; ++kk) {
            const int f = RL->f[kk], e = f + RL->len[kk];
            const uint32_t m = low_nibbles_bf(e - j0) & ~low_nibbles_bf(f - j0);
            if (m) {
                const int at = RL->a[kk] + (j0 - f);                     // >= -7 where the mask is set
                v |= (RL->kind[kk] ? ref_word(B.ref4, (int64_t)max(at, -8)) : stage_at(sg, at)) & m;
            }
            if (e >= j0 + 8 || kk >= n_runs - 1) break;
        }
        oseq[w] = v & low_nibbles_bf(new_len - j0);
    }
    // qualities in printed (= BAM) order; the runs index the forward-orientation array (anonymizer_methods.py:95, 187,
    // 195: quirk Q2), so for a reverse read a run [f, f + len) is printed at [new_len - f - len, new_len - f) and its BAM
    // bytes ascend with the printed index
    uint32_t* oq = reinterpret_cast<uint32_t*>(O.out_qual + 32ull * R.qual16);
    int kq = 0;                                                           // position in printed order: run kq is run n_runs - 1 - kq of a reverse read
    for (int w = glane; w < units * 8; w += kGroup) {
        const int p0 = w << 2;
        uint32_t v = 0u;
        auto printed = [&](int kp, int* kf, int* ln) -> int {             // printed start, forward index and length of the kp-th run in printed order
            *kf = R.reverse ? n_runs - 1 - kp : kp;
            *ln = RL->len[*kf];
            return R.reverse ? new_len - RL->f[*kf] - *ln : RL->f[*kf];
        };
        int kf, ln;
        while (kq < n_runs - 1 && printed(kq, &kf, &ln) + ln <= p0) ++kq;
#pragma unroll 1
        for (int kk = kq;; ++kk) {
            const int f = printed(kk, &kf, &ln), e = f + ln;
            const uint32_t m = low_bytes_bf(e - p0) & ~low_bytes_bf(f - p0);
            if (m) {
                const int kind = RL->kind[kf];
                uint32_t val;
                if (kind) val = (kind == 1 ? mean0 : mean1) * 0x01010101u;
                else val = qual_at(qs, p0 + (R.reverse ? L - new_len - RL->a[kf] + RL->f[kf] : RL->a[kf] - RL->f[kf]));   // BAM byte of printed byte p0
                v |= val & m;
            }
            if (e >= p0 + 4 || kk >= n_runs - 1) break;
        }
        oq[w] = v & low_bytes_bf(new_len - p0);
    }
}

// ------------------------------------------------------------------ LANE = RECORD
// emit_records_kernel writes the common special records - reads up to 156 bases with another CIGAR and SNV hits only
// (kind 2) or with ONE germline indel (kind 3) - with ONE LANE PER RECORD.  Eight lanes per record (emit_special_kernel,
// round 1) repeat every scalar step of a record eight times and leave a third of the lanes idle in the word loops; a
// lane of its own runs the whole record as straight-line arithmetic, and the 32 records of a warp step differ only in
// data, never in control flow.  What a lane cannot do efficiently is touch global memory (32 lanes = 32 cache lines per
// instruction), so a warp step is:
//   1. lane = record: descriptor, edit description, quality-record slot (coalesced: the descriptors are dense); what
//      this kernel does not take (two edits, many hits, longer reads) is listed for emit_special_kernel;
//   2. the WARP copies the 32 source records (and quality records) into shared memory with asynchronous 4-byte copies
//      (cp.async, zero fill for the padding words), one record per step with consecutive lanes on consecutive words, all
//      in flight together - a lane's record lands in a private row of odd stride (no bank conflicts);
//   3. lane = record: SNV masking (anonymizer_methods.py:170-176) in the row; then the final array is three pieces -
//      [0, p) = source as is, [p, ins_end) = re-inserted reference bases with quality floor(mean) (DEL only),
//      [ins_end, new_len) = source shifted (anonymizer_methods.py:178-203; no edit: one piece) - and every output word is
//      merged from them by masks, four words per 128-bit store; qualities in printed (= BAM) order (quirks Q1, Q2).
constexpr int kRecMaxL = 156;            // bases of a read handled here (longer: emit_special_kernel)
constexpr int kRecSeqW = 23;             // row: [0] zero pad (source index -8 .. -1), [1, 21) record words, [21, 23) zero
constexpr int kRecQualW = 41;            // row: [0] zero pad, [1, 40) quality words, [40, 41) zero
constexpr int kRecWarps = 4;
constexpr int kRecMaxEdit = 32;          // longest germline indel handled here: bounds how far a shifted row pointer reaches outside its row
constexpr int kRecMaxNew = 188;          // longest record written here (six 32-base units)
// guards: a shifted row pointer reaches at most 4 (bases) / 8 (qualities) words in front of its row and, in the last unit of
// the longest record, 6 / 16 words behind it; the values read there are never used, the addresses must exist
struct RecWarp { uint32_t guard0[16]; uint32_t seq[32][kRecSeqW]; uint32_t guard1[16]; uint32_t qual[32][kRecQualW]; uint32_t guard2[24]; };
static_assert(kRecMaxEdit / 8 <= 16 && 1 + kRecMaxEdit / 8 + 4 * ((kRecMaxNew + 31) / 32) + 1 - kRecSeqW <= 16, "sequence rows: reach of a shifted pointer");
static_assert(kRecMaxEdit / 4 <= 16 && 1 + kRecMaxEdit / 4 + 8 * ((kRecMaxNew + 31) / 32) + 1 - kRecQualW <= 24, "quality rows: reach of a shifted pointer");
static_assert(kRecSeqW % 2 == 1 && kRecQualW % 2 == 1, "rows of odd word stride");
static_assert(kRecQualW == 32 + (32 - kRecSeqW) && 4 * (kRecQualW - 2) >= kRecMaxL && 8 * (kRecSeqW - 3) >= kRecMaxL, "two copies per lane stage a record");

__device__ __forceinline__ uint32_t rec_stage_at(const uint32_t* row, int sidx) {      // 8 staged bases from source index sidx (any value)
    const int s = min(max(sidx, -8), 8 * (kRecSeqW - 2) - 1);
    const int wi = (s >> 3) + 1;
    return __funnelshift_r(row[wi], row[wi + 1], (uint32_t)(s & 7) * 4u);
}
__device__ __forceinline__ uint32_t rec_qual_at(const uint32_t* row, int bidx) {       // 4 staged quality bytes from BAM byte bidx (any value)
    const int s = min(max(bidx, -4), 4 * (kRecQualW - 2) - 1);
    const int wi = (s >> 2) + 1;
    return __funnelshift_r(row[wi], row[wi + 1], (uint32_t)(s & 3) * 8u);
}
// asynchronous 4-byte copy global -> shared; n_src = 0 writes zeros instead (SASS: LDGSTS)
__device__ __forceinline__ void cp_async4(uint32_t* dst, const uint32_t* src, uint32_t n_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(n_src) : "memory");
}

__global__ void __launch_bounds__(32 * kRecWarps) emit_records_kernel(BatchView B, ResultView O, EmitScratch2 E) {
    extern __shared__ __align__(16) uint8_t emit_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    RecWarp& W = reinterpret_cast<RecWarp*>(emit_smem)[warp];
    uint32_t* sq = W.seq[lane];
    uint32_t* qq = W.qual[lane];
    const uint32_t n_x = (uint32_t)min((int64_t)*E.n_special, O.cap_records);   // slots past the capacity were never written
    const uint32_t stride = gridDim.x * kRecWarps * 32;
    for (uint32_t jb = (blockIdx.x * kRecWarps + warp) * 32; jb < n_x; jb += stride) {      // warp-uniform
        // ---- 1. lane = record
        const uint32_t j = jb + lane;
        int cls = -1;                                                    // 0: SNV-only, 1: one germline indel
        bool rare = false;
        uint32_t so = 0u, c0 = 0u, c1 = 0u, germ_n = 0u, qunit = 0u, seq16 = 0u, qual16 = 0u, rd = 0u;
        int pos = 0, L = 0, s = 0, new_len = 0, col_begin = 0;
        bool reverse = false;
        int ed_pos = 0, ed_len = 0, p = 0, ins_end = 0, shift = 0;        // the edit: [p, ins_end) re-inserted elements, source = final + shift behind them
        bool is_del = false;
        if (j < n_x) {
            const uint4* dp = E.sdesc + 4ull * j;
            const uint4 d0 = dp[0], d1 = dp[1], d2 = dp[2];
            const uint32_t kind = (d0.z >> 16) & 15u;                    // 0: the slot of a session that did not fit
            so = d0.x; pos = (int)d0.y; L = (int)(d0.z & 0xffffu); reverse = ((d0.z >> 20) & 1u) != 0u; s = (int)d0.w;
            rd = d1.x; new_len = (int)d1.y; seq16 = d1.z; qual16 = d1.w;
            c0 = d2.x; c1 = d2.y; col_begin = (int)d2.z;
            p = ins_end = new_len;
            if (kind == 4u || ((kind == 2u || kind == 3u) && L > kRecMaxL)) rare = true;
            else if (kind == 2u) cls = 0;
            else if (kind == 3u) {
                const uint4* ap = reinterpret_cast<const uint4*>(O.out_qual + 32ull * qual16);
                const uint4 x0 = ap[0], x1 = ap[1];                      // EditAux written by the resolve kernel
                if ((x1.z & 0xffu) != 1u || (x0.z & 0x7fffffffu) > (uint32_t)kRecMaxEdit || new_len > kRecMaxNew) rare = true;   // two edits, or an edit beyond the rows' guards: emit_special_kernel
                else {
                    Ed2 Ed; Ed.ne = 1; Ed.n_del = (int)((x1.z >> 8) & 0xffu);
                    Ed.irp[0] = (int)x0.x; Ed.pos[0] = (int)x0.y; Ed.len[0] = (int)(x0.z & 0x7fffffffu);
                    Ed.irp[1] = 0; Ed.pos[1] = 0; Ed.len[1] = 0; Ed.mean[0] = Ed.mean[1] = 0u;
                    clamp_edits2(Ed, L);
                    is_del = Ed.n_del == 1; ed_pos = Ed.pos[0]; ed_len = Ed.len[0];
                    p = Ed.p[0]; shift = is_del ? -ed_len : Ed.e[0] - Ed.p[0]; ins_end = is_del ? p + ed_len : p;
                    if (is_del && (int64_t)ed_pos + ed_len > B.ref_len) raise_error(O.totals, GA_ERR_LENGTH_MISMATCH, rd);
                    // the read's quality record: dense upload, or the slot the resolve kernel predicted in the sparse index
                    // (verified; a caller may list more reads than those with I/D ops), or a search of the slice
                    qunit = 0xffffffffu;
                    if (B.qual && !B.qual_reads) qunit = so;
                    else if (B.qual) {
                        const int64_t qi = (int64_t)x1.w;
                        if (qi < B.n_qual && __ldg(B.qual_reads + qi) == (int32_t)rd) qunit = __ldg(B.qual_off16 + qi);
                        else {
                            const uint4 d3 = dp[3];
                            const uint8_t* qp = qual_record_in(B, (int64_t)rd, (int64_t)d3.x, (int64_t)d3.y);
                            if (qp) qunit = (uint32_t)((qp - B.qual) >> 5);
                        }
                    }
                    if (qunit == 0xffffffffu) raise_error(O.totals, GA_ERR_BAD_ARGUMENT, rd);   // no quality record for an indel-masked read
                    else cls = 1;
                }
            }
            if (cls >= 0) germ_n = __ldg(E.germ + (size_t)s * kGermStride);
        }
        {   // what is left for emit_special_kernel
            const uint32_t rm = __ballot_sync(0xffffffffu, rare);
            if (rm) {
                uint32_t at = 0u;
                if (lane == 0) at = atomicAdd(E.n_rare, (uint32_t)__popc(rm));
                at = __shfl_sync(0xffffffffu, at, 0);
                if (rare) E.rare_list[at + __popc(rm & ((1u << lane) - 1u))] = j;
            }
        }
        const uint32_t live = __ballot_sync(0xffffffffu, cls >= 0);
        if (!live) continue;
        // ---- 2. the warp stages the records: one record per step, consecutive lanes on consecutive words, all copies in flight
#pragma unroll 4
        for (int t = 0; t < 32; ++t) {
            if (!((live >> t) & 1u)) continue;
            const uint32_t t_so = __shfl_sync(0xffffffffu, so, t), t_qu = __shfl_sync(0xffffffffu, qunit, t);
            const int t_L = __shfl_sync(0xffffffffu, L, t), t_cls = __shfl_sync(0xffffffffu, cls, t);
            const uint32_t* rec = reinterpret_cast<const uint32_t*>(B.seq4 + 16ull * t_so);
            const uint32_t* qrec = reinterpret_cast<const uint32_t*>(B.qual + 32ull * t_qu);
            // two copies per lane: lanes 0-22 the sequence row and lanes 23-31 the tail of the quality row, then its first 32 words
            if (lane < kRecSeqW) {
                const int w = lane - 1;
                const bool data = w >= 0 && 8 * w < t_L;
                cp_async4(&W.seq[t][lane], rec + (data ? w : 0), data ? 4u : 0u);
            } else if (t_cls == 1) {
                const int jq = lane + (32 - kRecSeqW), w = jq - 1;        // 32 .. 40
                const bool data = 4 * w < t_L;
                cp_async4(&W.qual[t][jq], qrec + (data ? w : 0), data ? 4u : 0u);
            }
            if (t_cls == 1) {
                const int w = lane - 1;
                const bool data = w >= 0 && 4 * w < t_L;
                cp_async4(&W.qual[t][lane], qrec + (data ? w : 0), data ? 4u : 0u);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        // ---- 3a. lane = record: SNV masking - every germline allele of the session is carried through the CIGAR to its
        // query offset and, when the read shows it, replaced by the reference base
        {
            const uint32_t n_ops = c1 - c0;
            uint32_t cg0 = 0u, cg1 = 0u, cg2 = 0u, cg3 = 0u;             // the first four ops travel in registers
            if (cls >= 0) {
                if (n_ops > 0u) cg0 = __ldg(B.cigar + c0);
                if (n_ops > 1u) cg1 = __ldg(B.cigar + c0 + 1);
                if (n_ops > 2u) cg2 = __ldg(B.cigar + c0 + 2);
                if (n_ops > 3u) cg3 = __ldg(B.cigar + c0 + 3);
                const uint32_t tm = tail_mask(L, (L - 1) >> 3);          // the padding nibbles of the last word are not part of the read
                if (L > 0) sq[1 + ((L - 1) >> 3)] &= tm;
            }
            const uint32_t g_max = __reduce_max_sync(0xffffffffu, cls >= 0 ? germ_n : 0u);
            const uint32_t* ge = E.germ + (size_t)s * kGermStride + 4;
#pragma unroll 1
            for (uint32_t a = 0; a < g_max; ++a) {
                if (cls < 0 || a >= germ_n) continue;
                const uint32_t key = __ldg(ge + a), code = key & 15u;
                const int at = col_begin + (int)(key >> 4);
                if (at < pos) continue;                                  // the column lies in front of the read
                int rc = pos, q = 0;
#pragma unroll 1
                for (uint32_t ci = 0; ci < n_ops; ++ci) {
                    if (at < rc) break;                                  // the column lies before what is left of the read
                    const uint32_t cw = ci == 0u ? cg0 : ci == 1u ? cg1 : ci == 2u ? cg2 : ci == 3u ? cg3 : __ldg(B.cigar + c0 + ci);
                    const uint32_t op = cw & 15u;
                    const int ln = (int)(cw >> 4);
                    if (op == 0u || op == 7u || op == 8u) {
                        if (at < rc + ln) {
                            const int x = q + (at - rc);
                            if (x < L) {
                                const uint32_t sh = (uint32_t)(x & 7) * 4u, w = sq[1 + (x >> 3)];
                                if (((w >> sh) & 15u) == code) sq[1 + (x >> 3)] = w ^ ((code ^ ref_code(B.ref4, at)) << sh);
                            }
                            break;
                        }
                        q += ln; rc += ln;
                    } else if (op == 1u || op == 4u) q += ln;
                    else if (op == 2u || op == 3u) { if (at < rc + ln) break; rc += ln; }
                }
            }
        }
        // ---- 3b. quality of re-inserted bases: floor(mean(qualities)) (anonymizer_methods.py:193)
        uint32_t mean4 = 0u;
        if (__any_sync(0xffffffffu, is_del)) {
            uint32_t sum = 0u;
            if (is_del) {
#pragma unroll 8
                for (int w = 0; w < kRecQualW - 2; ++w) sum += __vsadu4(qq[1 + w] & low_bytes_bf(L - 4 * w), 0u);
            }
            mean4 = (is_del && L ? sum / (uint32_t)L : 0u) * 0x01010101u;
        }
        // ---- 3c. bases.  A word that lies wholly in front of the edit is the staged word, a word wholly behind it one funnel
        // shift of two staged words (constant shift per record); only the words that straddle a piece boundary or hold
        // re-inserted reference bases need the masked merge - they are rewritten afterwards, a handful per record.
        const int units = cls >= 0 ? max(1, (new_len + 31) >> 5) : 0;
        const int u_max = __reduce_max_sync(0xffffffffu, units);
        {
            uint4* out = reinterpret_cast<uint4*>(O.out_seq4 + 16ull * seq16);
            const int wf = p >> 3, wz = (new_len + 7) >> 3;
            const uint32_t* sb = sq + 1 + (shift >> 3);                  // may point a few words outside the row: guard words around the rows
            const uint32_t sh = (uint32_t)(shift & 7) * 4u;
#pragma unroll 1
            for (int u = 0; u < u_max; ++u) {
                if (u >= units) continue;
                const uint32_t* a4 = sq + 1 + 4 * u;
                const uint32_t* b4 = sb + 4 * u;
                const uint32_t b0 = b4[0], b1w = b4[1], b2w = b4[2], b3w = b4[3], b4w = b4[4];
                uint32_t o[4] = {__funnelshift_r(b0, b1w, sh), __funnelshift_r(b1w, b2w, sh), __funnelshift_r(b2w, b3w, sh), __funnelshift_r(b3w, b4w, sh)};
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int w = 4 * u + x;
                    if (w < wf) o[x] = a4[x];
                    if (w >= wz) o[x] = 0u;
                }
                out[u] = make_uint4(o[0], o[1], o[2], o[3]);
            }
            // the words around the edit and the last word
            uint32_t* ow = reinterpret_cast<uint32_t*>(out);
            auto merged = [&](int w) {
                const int j0 = w << 3;
                const uint32_t mA = low_nibbles_bf(p - j0), mAB = low_nibbles_bf(ins_end - j0);
                uint32_t v = (rec_stage_at(sq, j0) & mA) | (rec_stage_at(sq, j0 + shift) & ~mAB);
                const uint32_t mR = mAB & ~mA;
                if (mR) v |= ref_word(B.ref4, (int64_t)ed_pos + (j0 - p)) & mR;
                ow[w] = v & low_nibbles_bf(new_len - j0);
            };
            const int w_lo = p >> 3, w_hi = cls >= 0 ? min((max(ins_end, p + 1) - 1) >> 3, 4 * units - 1) : -1;
            const int n_fix = __reduce_max_sync(0xffffffffu, cls >= 0 ? w_hi - w_lo + 1 : 0);
#pragma unroll 1
            for (int t = 0; t < n_fix; ++t) if (w_lo + t <= w_hi) merged(w_lo + t);
            if (cls >= 0 && (new_len & 7) && (new_len >> 3) < 4 * units) merged(new_len >> 3);
        }
        // ---- 3d. qualities in printed (= BAM) order.  The edit indexes the forward-orientation array (quirk Q2), so for a
        // reverse read the pieces come in the opposite order: printed [0, b1) = BAM bytes as they are, [b1, b2) = the mean,
        // [b2, new_len) = BAM bytes shifted by d3.  Same scheme: whole words first, the straddling words afterwards.
        if (__any_sync(0xffffffffu, cls == 1)) {
            const int b1 = reverse ? new_len - ins_end : p, b2 = reverse ? new_len - p : ins_end, d3 = reverse ? L - new_len : shift;
            const int qunits = cls == 1 ? 2 * units : 0;
            const int q_max = __reduce_max_sync(0xffffffffu, qunits);
            uint4* out = reinterpret_cast<uint4*>(O.out_qual + 32ull * qual16);
            const int wf = b1 >> 2, wb = (b2 + 3) >> 2, wz = (new_len + 3) >> 2;
            const uint32_t* qb = qq + 1 + (d3 >> 2);
            const uint32_t sh = (uint32_t)(d3 & 3) * 8u;
#pragma unroll 1
            for (int u = 0; u < q_max; ++u) {
                if (u >= qunits) continue;
                const uint32_t* a4 = qq + 1 + 4 * u;
                const uint32_t* b4 = qb + 4 * u;
                const uint32_t b0 = b4[0], b1w = b4[1], b2w = b4[2], b3w = b4[3], b4w = b4[4];
                uint32_t o[4] = {__funnelshift_r(b0, b1w, sh), __funnelshift_r(b1w, b2w, sh), __funnelshift_r(b2w, b3w, sh), __funnelshift_r(b3w, b4w, sh)};
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int w = 4 * u + x;
                    if (w < wb) o[x] = mean4;
                    if (w < wf) o[x] = a4[x];
                    if (w >= wz) o[x] = 0u;
                }
                out[u] = make_uint4(o[0], o[1], o[2], o[3]);
            }
            uint32_t* ow = reinterpret_cast<uint32_t*>(out);
            auto merged = [&](int w) {
                const int p0 = w << 2;
                const uint32_t m1 = low_bytes_bf(b1 - p0), m12 = low_bytes_bf(b2 - p0);
                const uint32_t v = (rec_qual_at(qq, p0) & m1) | (mean4 & m12 & ~m1) | (rec_qual_at(qq, p0 + d3) & ~m12);
                ow[w] = v & low_bytes_bf(new_len - p0);
            };
            if (cls == 1) {
                if ((b1 & 3) && (b1 >> 2) < 4 * qunits) merged(b1 >> 2);
                if ((b2 & 3) && (b2 >> 2) < 4 * qunits) merged(b2 >> 2);
                if ((new_len & 3) && (new_len >> 2) < 4 * qunits) merged(new_len >> 2);
            }
        }
        __syncwarp();                                                     // the rows are rewritten by the next 32 records
    }
}

// kind 4: a clean read with more than two germline hits - copy, looking every mismatch up in the session's germline list.
__device__ __forceinline__ void emit_many_hits_quad(const BatchView& B, const ResultView& O, const GermList& germ, const SpecRec& R, int glane) {
    const int L = R.new_len;
    const uint4* rec = reinterpret_cast<const uint4*>(B.seq4 + 16ull * R.src_unit);
    uint4* out = reinterpret_cast<uint4*>(O.out_seq4 + 16ull * R.seq16);
    const int units = (L + 31) >> 5;
    for (int u = glane; u < units; u += kGroup) {
        const uint4 vv = ldg128(rec + u);
        const int64_t ni = (int64_t)R.pos + 32 * u + 8;
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
                if (b != 15u && germ(R.pos + 8 * wi + nb - R.col_begin, b)) v = (v & ~(0xfu << (nb * 4))) | (((f[q] >> (nb * 4)) & 15u) << (nb * 4));
            }
            w[q] = v;
        }
        out[u] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Records of kind >= 2 (reads with other CIGARs, reads with many hits), taken densely from the list the resolve kernels
// packed.  A warp takes 32 records at a time, LANE = RECORD: descriptor, edit description (EditAux, parked by the
// resolve kernel in the record's quality slot), the germline-list length and the place of the quality record arrive
// with coalesced loads - one chain of dependent round trips per 32 records instead of one per record - and go to the
// warp's table in shared memory.  The records are then regrouped BY KIND and taken four at a time, 8 lanes each, so
// that a warp step runs one code path: SNV-only / one edit / two edits / many hits / (reads beyond 248 bases: the
// general forms of ga_record_ops.cuh).
constexpr int kSpecFields = 23;
enum { SF_SO, SF_POS, SF_LZ, SF_S, SF_R, SF_NEWLEN, SF_SEQ16, SF_QUAL16, SF_C0, SF_C1, SF_COLB, SF_QLO, SF_QHI, SF_AUX0, SF_QUNIT = SF_AUX0 + 8, SF_GERMN };
static_assert(SF_GERMN + 1 == kSpecFields, "field count");
constexpr int kSpecClasses = 5;         // 0 SNV-only, 1 one edit, 2 two edits, 3 many hits, 4 long reads
struct SpecWarp { uint32_t f[kSpecFields][32]; uint8_t order[kSpecClasses][32]; };

__global__ void __launch_bounds__(kThreads, 4) emit_special_kernel(BatchView B, const SessionDesc* __restrict__ descs, ResultView O, EmitScratch2 E) {
    extern __shared__ __align__(16) uint8_t emit_smem[];
    SpecWarp* tabs = reinterpret_cast<SpecWarp*>(emit_smem);
    uint32_t* stage_all = reinterpret_cast<uint32_t*>(tabs + kThreads / 32);
    uint32_t* qstage_all = stage_all + (kThreads / kGroup) * kStageW;
    RunList* runs_all = reinterpret_cast<RunList*>(qstage_all + (kThreads / kGroup) * kQStageW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, group = tid / kGroup, glane = tid % kGroup, gw = lane >> 3;
    SpecWarp& T = tabs[warp];
    uint32_t* sg = stage_all + group * kStageW;
    uint32_t* qs = qstage_all + group * kQStageW;
    RunList* RL = runs_all + group;
    const uint32_t n_x = *E.n_rare;                                     // the records emit_records_kernel left (two edits, many hits, long reads)
    // records per warp step: 32, or fewer when there are only a few records (then every warp takes one short step instead
    // of a few warps walking eight groups of four one after the other)
    const uint32_t warps_total = gridDim.x * (kThreads / 32);
    const uint32_t per_warp = min(32u, max(4u, ((n_x + warps_total - 1u) / warps_total + 3u) & ~3u));
    const uint32_t stride = warps_total * per_warp;
    for (uint32_t jb = (blockIdx.x * (kThreads / 32) + warp) * per_warp; jb < n_x; jb += stride) {      // warp-uniform
        // ---- lane = record
        int cls = -1;
        if ((uint32_t)lane < per_warp && jb + lane < n_x) {
            const uint4* dp = E.sdesc + 4ull * E.rare_list[jb + lane];
            const uint4 d0 = dp[0], d1 = dp[1], d2 = dp[2], d3 = dp[3];
            const uint32_t kind = (d0.z >> 16) & 15u;                    // 0: the slot of a session that did not fit
            const int L = (int)(d0.z & 0xffffu);
            T.f[SF_SO][lane] = d0.x; T.f[SF_POS][lane] = d0.y; T.f[SF_LZ][lane] = d0.z; T.f[SF_S][lane] = d0.w;
            T.f[SF_R][lane] = d1.x; T.f[SF_NEWLEN][lane] = d1.y; T.f[SF_SEQ16][lane] = d1.z; T.f[SF_QUAL16][lane] = d1.w;
            T.f[SF_C0][lane] = d2.x; T.f[SF_C1][lane] = d2.y; T.f[SF_COLB][lane] = d2.z; T.f[SF_QLO][lane] = d3.x; T.f[SF_QHI][lane] = d3.y;
            const bool is_long = ((L + 7) >> 3) > kGroupStage - 1;
            if (kind >= 2u && kind <= 4u) T.f[SF_GERMN][lane] = __ldg(E.germ + (size_t)d0.w * kGermStride);
            if (kind == 2u) cls = is_long ? 4 : 0;
            else if (kind == 4u) cls = 3;
            else if (kind == 3u) {
                const uint4* ap = reinterpret_cast<const uint4*>(O.out_qual + 32ull * d1.w);
                const uint4 x0 = ap[0], x1 = ap[1];                      // EditAux written by the resolve kernel
                T.f[SF_AUX0 + 0][lane] = x0.x; T.f[SF_AUX0 + 1][lane] = x0.y; T.f[SF_AUX0 + 2][lane] = x0.z; T.f[SF_AUX0 + 3][lane] = x0.w;
                T.f[SF_AUX0 + 4][lane] = x1.x; T.f[SF_AUX0 + 5][lane] = x1.y; T.f[SF_AUX0 + 6][lane] = x1.z; T.f[SF_AUX0 + 7][lane] = x1.w;
                // the read's quality record: dense upload, or the slot the resolve kernel predicted in the sparse index
                // (verified; a caller may list more reads than those with I/D ops), or a search of the slice
                uint32_t qunit = 0xffffffffu;
                if (B.qual && !B.qual_reads) qunit = d0.x;
                else if (B.qual) {
                    const int64_t qi = (int64_t)x1.w;
                    if (qi < B.n_qual && __ldg(B.qual_reads + qi) == (int32_t)d1.x) qunit = __ldg(B.qual_off16 + qi);
                    else {
                        const uint8_t* qp = qual_record_in(B, (int64_t)d1.x, (int64_t)d3.x, (int64_t)d3.y);
                        if (qp) qunit = (uint32_t)((qp - B.qual) >> 5);
                    }
                }
                T.f[SF_QUNIT][lane] = qunit;
                if (qunit == 0xffffffffu) raise_error(O.totals, GA_ERR_BAD_ARGUMENT, d1.x);   // no quality record for an indel-masked read
                else cls = is_long ? 4 : ((x1.z & 0xffu) == 1u ? 1 : 2);
            }
        }
        unsigned long long n_cls = 0ull;                                  // records per kind, 8 bits each
#pragma unroll
        for (int k = 0; k < kSpecClasses; ++k) {
            const uint32_t m = __ballot_sync(0xffffffffu, cls == k);
            n_cls |= (unsigned long long)__popc(m) << (8 * k);
            if (cls == k) T.order[k][__popc(m & ((1u << lane) - 1u))] = (uint8_t)lane;
        }
        __syncwarp();
        // ---- four records of one kind per step, 8 lanes each
#pragma unroll 1
        for (int k = 0; k < kSpecClasses; ++k) {
            const uint32_t n_k = (uint32_t)(n_cls >> (8 * k)) & 0xffu;
#pragma unroll 1
            for (uint32_t q0 = 0; q0 < n_k; q0 += 4) {
                const bool act = q0 + gw < n_k;
                const int t = act ? (int)T.order[k][q0 + gw] : 0;
                SpecRec R;
                const uint32_t lz = T.f[SF_LZ][t];
                R.src_unit = T.f[SF_SO][t]; R.pos = (int)T.f[SF_POS][t]; R.L = (int)(lz & 0xffffu); R.reverse = ((lz >> 20) & 1u) != 0u; R.s = (int)T.f[SF_S][t];
                R.r = (int64_t)T.f[SF_R][t]; R.new_len = (int)T.f[SF_NEWLEN][t]; R.seq16 = T.f[SF_SEQ16][t]; R.qual16 = T.f[SF_QUAL16][t];
                R.c0 = T.f[SF_C0][t]; R.c1 = T.f[SF_C1][t]; R.col_begin = (int)T.f[SF_COLB][t]; R.germ_n = T.f[SF_GERMN][t]; R.qunit = T.f[SF_QUNIT][t];
                Ed2 Ed; Ed.ne = 0; Ed.n_del = 0;
#pragma unroll
                for (int q = 0; q < 2; ++q) { Ed.irp[q] = 0; Ed.len[q] = 0; Ed.pos[q] = 0; Ed.mean[q] = 0u; Ed.p[q] = 0; Ed.e[q] = 0; }
                if (k == 1 || k == 2 || k == 4) {
                    const uint32_t kind = (lz >> 16) & 15u;
                    if (act && kind == 3u) {
                        Ed.irp[0] = (int)T.f[SF_AUX0][t]; Ed.pos[0] = (int)T.f[SF_AUX0 + 1][t]; Ed.len[0] = (int)(T.f[SF_AUX0 + 2][t] & 0x7fffffffu);
                        Ed.irp[1] = (int)T.f[SF_AUX0 + 3][t]; Ed.pos[1] = (int)T.f[SF_AUX0 + 4][t]; Ed.len[1] = (int)(T.f[SF_AUX0 + 5][t] & 0x7fffffffu);
                        const uint32_t nn = T.f[SF_AUX0 + 6][t];
                        Ed.ne = (int)(nn & 0xffu); Ed.n_del = (int)((nn >> 8) & 0xffu);
                        clamp_edits2(Ed, R.L);
                    }
                }
                if (k == 0) emit_snv_only_quad(B, O, E, act, R, sg, glane);
                else if (k == 1) emit_one_edit_quad(B, O, E, act, R, Ed, sg, qs, glane);
                else if (k == 2) emit_two_edit_quad(B, O, E, act, R, Ed, sg, qs, RL, glane);
                else {
                    GermList germ; germ.e = E.germ + (size_t)R.s * kGermStride + 4; germ.n = act ? R.germ_n : 0u;
                    if (k == 3) { if (act) emit_many_hits_quad(B, O, germ, R, glane); }
                    else {                                                // reads beyond the staging area: the general forms
                        const bool indel = act && ((lz >> 16) & 15u) == 3u;
                        if (act && !indel) {
                            int units = (R.new_len + 31) >> 5; if (units < 1) units = 1;
                            uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * R.seq16);
                            masked_words_g(B, R.r, R.pos, R.new_len, R.c0, R.c1, R.col_begin, units * 4, glane, kGroup, germ, [&](int wd, uint32_t v) { oseq[wd] = v; });
                        }
                        if (__any_sync(0xffffffffu, indel))
                            emit_indel_group_t(B, O.totals, O, indel, Ed, R.r, R.col_begin, (int64_t)T.f[SF_QLO][t], (int64_t)T.f[SF_QHI][t], sg, R.seq16, R.qual16, R.new_len, glane, germ);
                    }
                }
                __syncwarp();                                             // the staging areas are reused by the next step
            }
        }
        __syncwarp();                                                     // the table is rewritten by the next 32 records
    }
}
constexpr size_t kEmitSpecialSmem = sizeof(SpecWarp) * (kThreads / 32) + sizeof(uint32_t) * (kThreads / kGroup) * (kStageW + kQStageW) + sizeof(RunList) * (kThreads / kGroup);

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
