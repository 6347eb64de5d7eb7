// ga_device.cuh - device-side building blocks shared by the session kernels.
//
// Reference semantics being implemented (citations into /root/reference/src/GenomeAnonymizer):
//   SNV discovery      variation_classifier.py:144-150   (base != ref, base != 'N', ref in ACGT)
//   indel discovery    variation_classifier.py:52-107    (pos, in_read_pos quirk, allele slices)
//   key equality       variants.py:83-96
//   masking            anonymizer_methods.py:170-203, 254-270, 537-556
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ga_b200.h"

namespace ga {

constexpr int kThreads = 256;          // threads per session CTA
constexpr int kColsCap = 4096;         // smem allele-table columns per session (small path)
constexpr int kReadsCap = 4096;        // candidate reads per session (small path)
constexpr int kObsCap = 1024;          // indel observations per session (small path)

// observation meta word
constexpr uint32_t kMetaIns = 1u << 31;
constexpr uint32_t kMetaDs = 1u << 30;
constexpr uint32_t kMetaGerm = 1u << 29;
constexpr uint32_t kMetaRep = 1u << 28;
constexpr uint32_t kMetaSeenT = 1u << 24;   // on a key's representative observation: a tumor / normal read showed the key
constexpr uint32_t kMetaSeenN = 1u << 25;
constexpr uint32_t kMetaCol = 1u << 26;     // on a key's representative: a normal read that shows the key also covers its position
constexpr uint32_t kMetaTrail = 1u << 27;   // an insertion with no reference-consuming op behind it: it sits at its read's reference_end,
                                            // a position the read itself does not cover (see normal_covers, ga_session_kernel.cuh)
constexpr uint32_t kMetaLenMask = (1u << 24) - 1;   // op lengths are below 65,536 (the read length is 16 bits)

// msize word (per candidate read, after analysis)
constexpr uint32_t kModFlag = 1u << 30;
constexpr uint32_t kQualFlag = 1u << 31;
constexpr uint32_t kLenMask = (1u << 30) - 1;

struct SessionDesc {
    int32_t t_begin, t_end, n_begin, n_end;   // candidate read ranges (pos < last, pos > first - maxspan)
    int32_t col_begin, n_cols;                // allele-table span (conservative)
    int32_t obs_bound;                        // upper bound on I/D ops among the candidates
    int32_t big;                              // 1: tables live in the CTA's global scratch
    int32_t qt_begin, qt_end, qn_begin, qn_end; // sparse quality records (ga_reads.qual_reads) of the candidate ranges
    uint32_t t_seq_lo, t_seq_n, n_seq_lo, n_seq_n;   // seq4 units [lo, lo+n) holding the candidate records (prefetch hint)
    uint32_t t_cig_lo, t_cig_n, n_cig_lo, n_cig_n;   // CIGAR words of the candidate reads (prefetch hint)
};

// View of one batch + reference, passed by value to kernels.
struct BatchView {
    const int32_t* pos;
    const uint32_t* len_flag;
    const uint32_t* seq_off16;
    const uint32_t* cigar_off;
    const uint32_t* cigar;
    const uint8_t* seq4;
    const uint8_t* qual;
    const int32_t* qual_reads;
    const uint32_t* qual_off16;
    int64_t n_reads, n_tumor, n_qual;
    const uint32_t* ref4;      // 4-bit reference, word 0 is padding: base p lives in nibble (p + 8)
    int64_t ref_len;
};

struct SessView {
    const int32_t* first;
    const int32_t* last;
    const int32_t* keep_type;
    const int32_t* keep_pos;
    const int32_t* keep_end;
    const int32_t* keep_len;
    const uint32_t* keep_allele_off;
    const uint8_t* keep_alleles;
    int32_t n_sessions;
};

struct ResultView {
    int64_t cap_records, cap_seq16, cap_qual16;
    int32_t* mod_session;
    int32_t* mod_read;
    uint32_t* mod_len;
    uint32_t* mod_seq_off16;
    uint32_t* mod_qual_off16;
    uint8_t* out_seq4;
    uint8_t* out_qual;
    uint32_t* sess_counts;
    ga_totals* totals;
};

__device__ __forceinline__ bool is_acgt(uint32_t c) { return c == 1u || c == 2u || c == 4u || c == 8u; }

__device__ __forceinline__ void raise_error(ga_totals* t, uint32_t code, uint32_t detail) {
    if (atomicCAS(&t->error, 0u, code) == 0u) t->error_detail = detail;
}

// 8 reference nibbles starting at reference position p (p >= -8).
__device__ __forceinline__ uint32_t ref_word(const uint32_t* __restrict__ ref4, int64_t p) {
    const int64_t ni = p + 8;
    const uint32_t lo = __ldg(ref4 + (ni >> 3));
    const uint32_t sh = (uint32_t)(ni & 7) * 4u;
    if (sh == 0) return lo;
    const uint32_t hi = __ldg(ref4 + (ni >> 3) + 1);
    return __funnelshift_r(lo, hi, sh);
}
__device__ __forceinline__ uint32_t ref_code(const uint32_t* __restrict__ ref4, int64_t p) {
    const int64_t ni = p + 8;
    return (__ldg(ref4 + (ni >> 3)) >> ((uint32_t)(ni & 7) * 4u)) & 15u;
}
__device__ __forceinline__ uint32_t read_code(const uint32_t* __restrict__ rec, int k) {
    return (__ldg(rec + (k >> 3)) >> ((k & 7) * 4)) & 15u;
}

// Reference span of a read (htslib bam_endpos: M, D, N, =, X consume the reference).
__device__ __forceinline__ int ref_span_of(const uint32_t* __restrict__ cigar, uint32_t c0, uint32_t c1) {
    int span = 0;
    for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t w = __ldg(cigar + c), op = w & 15u;
        if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) span += (int)(w >> 4);
    }
    return span;
}

// Compare one aligned (M/=/X) segment of a read with the reference, 8 bases per 32-bit word.
// f(q, refpos, base_code, ref_code) is called for every base that is an SNV candidate:
// base != ref, base != N, ref in {A,C,G,T}  (variation_classifier.py:147-150).
template <class F>
__device__ __forceinline__ void scan_segment(const uint32_t* __restrict__ rec, const uint32_t* __restrict__ ref4,
                                             int qs, int qe, int rs, F&& f) {
    if (qe <= qs) return;
    const int w0 = qs >> 3, w1 = (qe - 1) >> 3;
    for (int w = w0; w <= w1; ++w) {
        const int qb = w << 3;                               // query index of nibble 0 of this word
        const uint32_t rw = __ldg(rec + w);
        const uint32_t fw = ref_word(ref4, (int64_t)rs + (qb - qs));
        uint32_t mask = 0xffffffffu;
        if (qb < qs) mask &= 0xffffffffu << ((qs - qb) * 4);
        if (qb + 8 > qe) mask &= 0xffffffffu >> ((qb + 8 - qe) * 4);
        uint32_t x = (rw ^ fw) & mask;
        while (x) {
            const int k = (__ffs(x) - 1) >> 2;
            x &= ~(0xfu << (k * 4));
            const uint32_t b = (rw >> (k * 4)) & 15u, rf = (fw >> (k * 4)) & 15u;
            if (b != 15u && is_acgt(rf)) f(qb + k, rs + (qb + k - qs), b, rf);
        }
    }
}

// Block-wide exclusive scan of one value per thread (kThreads threads).  `tmp` holds kThreads/32 + 1 words.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* tmp, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    __syncthreads();                      // tmp may still be read from a previous call
    if (lane == 31) tmp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = lane < kThreads / 32 ? tmp[lane] : 0u;
        uint32_t si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, si, d);
            if (lane >= d) si += n;
        }
        if (lane < kThreads / 32) tmp[lane] = si - s;
        if (lane == kThreads / 32 - 1) tmp[kThreads / 32] = si;
    }
    __syncthreads();
    if (total) *total = tmp[kThreads / 32];
    return tmp[warp] + inc - v;
}

}  // namespace ga
