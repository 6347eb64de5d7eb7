// ga_emit_kernel.cuh - stage 3 of the streaming pipeline: grid-wide emission of the compacted modified records
// (north_star jobs (3) + (4)).
//
// The resolve kernel (ga_resolve_kernel.cuh) decides WHAT changes: it allocates the output slots, writes the record
// headers and hands over, per record, a 16-byte descriptor and, per session, the list of germline SNV alleles.
// This kernel writes the record bodies with full-chip parallelism and no block-level synchronisation.  A warp takes
// 32 records at a time: their descriptors and germline lists arrive with two coalesced round trips, then a group
// of 8 lanes copies each record with 128-bit loads / stores (four records per warp step).
//   kind 1  clean read, SNV-only   copy, replacing every base whose allele is germline by the reference base
//                                  (anonymizer_methods.py:170-176); qualities untouched
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
    const unsigned long long n_all = O.totals->n_modified;
    const int64_t n = (int64_t)(n_all < (unsigned long long)O.cap_records ? n_all : (unsigned long long)O.cap_records);
    const int64_t warp_global = (int64_t)blockIdx.x * (kThreads / 32) + (tid >> 5);
    const int64_t stride = (int64_t)gridDim.x * (kThreads / 32) * 32;
    for (int64_t k0 = warp_global * 32; k0 < n; k0 += stride) {
        // ---- lane = record: descriptor, output slot and the session's germline list (two round trips)
        const int64_t k = k0 + lane;
        const uint32_t kind = k < n ? E.kind[k] : 0u;
        uint4 d = make_uint4(0u, 0u, 0u, 0u), gh = d, ga = d;
        uint32_t dst16 = 0u;
        if (kind) { d = E.edesc[k]; dst16 = O.mod_seq_off16[k]; }
        if (kind) {
            const uint4* gp = reinterpret_cast<const uint4*>(E.germ + (size_t)d.w * kGermStride);
            gh = __ldg(gp); ga = __ldg(gp + 1);
        }
        const uint32_t m1 = __ballot_sync(0xffffffffu, kind == 1u), mx = __ballot_sync(0xffffffffu, kind >= 2u);
        // ---- kind 1: four records per step, 8 lanes each
#pragma unroll 2
        for (int step = 0; step < 8; ++step) {
            if (!((m1 >> (4 * step)) & 0xfu)) continue;
            const int rr = 4 * step + gw;
            const uint32_t r_kind = __shfl_sync(0xffffffffu, kind, rr);
            const uint32_t r_src = __shfl_sync(0xffffffffu, d.x, rr);
            const int r_pos = (int)__shfl_sync(0xffffffffu, d.y, rr);
            const int L = (int)__shfl_sync(0xffffffffu, d.z, rr);
            const uint32_t r_dst = __shfl_sync(0xffffffffu, dst16, rr);
            const uint32_t g_n = __shfl_sync(0xffffffffu, gh.x, rr);
            const int col_begin = (int)__shfl_sync(0xffffffffu, gh.y, rr);
            const uint32_t a0 = __shfl_sync(0xffffffffu, ga.x, rr), a1 = __shfl_sync(0xffffffffu, ga.y, rr);
            const uint32_t a2 = __shfl_sync(0xffffffffu, ga.z, rr), a3 = __shfl_sync(0xffffffffu, ga.w, rr);
            const uint32_t r_s = __shfl_sync(0xffffffffu, d.w, rr);
            if (r_kind != 1u) continue;
            const uint4* rec = reinterpret_cast<const uint4*>(B.seq4 + 16ull * r_src);
            uint4* out = reinterpret_cast<uint4*>(O.out_seq4 + 16ull * r_dst);
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
                        if (b == 15u) continue;
                        const uint32_t key = ((uint32_t)(r_pos + 8 * wi + nb - col_begin) << 4) | b;
                        bool hit = (g_n > 0u && key == a0) || (g_n > 1u && key == a1) || (g_n > 2u && key == a2) || (g_n > 3u && key == a3);
                        for (uint32_t e = 4; e < g_n && !hit; ++e) hit = __ldg(E.germ + (size_t)r_s * kGermStride + 4 + e) == key;
                        if (hit) v = (v & ~(0xfu << (nb * 4))) | (((f[q] >> (nb * 4)) & 15u) << (nb * 4));
                    }
                    w[q] = v;
                }
                out[u] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
        // ---- kinds 2 and 3 (reads with other CIGARs): one group of 8 lanes each, the read's arrays re-read
        if (mx) {
            int32_t my_r = 0; uint32_t my_len = 0u, my_q16 = 0u;
            if (kind >= 2u) { my_r = O.mod_read[k]; my_len = O.mod_len[k]; my_q16 = O.mod_qual_off16[k]; }
            for (int step = 0; step < 8; ++step) {
                if (!((mx >> (4 * step)) & 0xfu)) continue;
                const int rr = 4 * step + gw;
                const uint32_t r_kind = __shfl_sync(0xffffffffu, kind, rr);
                const int64_t r = (int64_t)__shfl_sync(0xffffffffu, my_r, rr);
                const int new_len = (int)__shfl_sync(0xffffffffu, my_len, rr);
                const uint64_t seq16 = __shfl_sync(0xffffffffu, dst16, rr), qual16 = __shfl_sync(0xffffffffu, my_q16, rr);
                const int s = (int)__shfl_sync(0xffffffffu, d.w, rr);
                const int col_begin = (int)__shfl_sync(0xffffffffu, gh.y, rr);
                GermList germ; germ.e = E.germ + (size_t)s * kGermStride + 4; germ.n = __shfl_sync(0xffffffffu, gh.x, rr);
                if (r_kind == 2u) {
                    const uint32_t c0 = __ldg(B.cigar_off + r), c1 = __ldg(B.cigar_off + r + 1);
                    int units = (new_len + 31) >> 5; if (units < 1) units = 1;
                    uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
                    masked_words_g(B, r, __ldg(B.pos + r), new_len, c0, c1, col_begin, units * 4, glane, kGroup, germ, [&](int wd, uint32_t v) { oseq[wd] = v; });
                }
                const bool indel = r_kind == 3u;
                if (__any_sync(0xffffffffu, indel)) {
                    Ed2 Ed; Ed.ne = 0; Ed.n_del = 0;
#pragma unroll
                    for (int q = 0; q < 2; ++q) { Ed.irp[q] = 0; Ed.len[q] = 0; Ed.pos[q] = 0; Ed.mean[q] = 0u; }
                    int64_t q_lo = 0, q_hi = 0;
                    if (indel) {
                        const uint4* ap = reinterpret_cast<const uint4*>(O.out_qual + 32ull * qual16);
                        const uint4 x0 = ap[0], x1 = ap[1];              // EditAux written by the resolve kernel
                        Ed.irp[0] = (int)x0.x; Ed.pos[0] = (int)x0.y; Ed.len[0] = (int)(x0.z & 0x7fffffffu);
                        Ed.irp[1] = (int)x0.w; Ed.pos[1] = (int)x1.x; Ed.len[1] = (int)(x1.y & 0x7fffffffu);
                        Ed.ne = (int)x1.z; Ed.n_del = (int)x1.w;
                        clamp_edits2(Ed, (int)(__ldg(B.len_flag + r) & 0xffffu));
                        const bool tumor = r < B.n_tumor;
                        q_lo = tumor ? descs[s].qt_begin : descs[s].qn_begin; q_hi = tumor ? descs[s].qt_end : descs[s].qn_end;
                    }
                    __syncwarp();                                         // every lane of the group has read the aux before it is overwritten
                    emit_indel_group_t(B, O.totals, O, indel, Ed, r, col_begin, q_lo, q_hi, stage[group], seq16, qual16, new_len, glane, germ);
                }
            }
        }
    }
}

}  // namespace ga
