// ga_emit_kernel.cuh - grid-wide emission of the compacted modified records (north_star jobs 3 + 4).
//
// The session kernel (ga_session_v2.cuh) decides WHAT changes: it allocates the output slots, writes the record
// headers and hands over, per session, the list of germline SNV alleles and, per indel-masked record, its edits.
// This kernel writes the record bodies with full-chip parallelism and no per-session barrier: a group of 8 lanes
// per record (four records per warp), coalesced 128-bit loads / stores.
//   kind 1  clean read, SNV-only   copy, replacing every base whose allele is germline by the reference base
//                                  (anonymizer_methods.py:170-176); qualities untouched
//   kind 2  other CIGAR, SNV-only  same, walking the CIGAR per 8-base word
//   kind 3  indel-masked           all DELs then all INSs at original offsets with the quality rules of
//                                  anonymizer_methods.py:178-203, 254-270 (emit_indel_group_t)
#pragma once
#include "ga_session_v2.cuh"

namespace ga {

struct GermList {
    const uint32_t* e; uint32_t n;
    __device__ __forceinline__ bool operator()(int col, uint32_t b) const {
        const uint32_t key = ((uint32_t)col << 4) | b;
        for (uint32_t k = 0; k < n; ++k) if (__ldg(e + k) == key) return true;
        return false;
    }
};

__global__ void __launch_bounds__(kThreads) emit_kernel(BatchView B, const SessionDesc* __restrict__ descs, ResultView O, EmitScratch X) {
    __shared__ uint32_t stage[kThreads / kGroup][kGroupStage];
    const int tid = threadIdx.x, group = tid / kGroup, glane = tid % kGroup, gw = group & 3;
    const unsigned long long n_all = O.totals->n_modified;
    const int64_t n = (int64_t)(n_all < (unsigned long long)O.cap_records ? n_all : (unsigned long long)O.cap_records);
    const int64_t per_pass = (int64_t)gridDim.x * (kThreads / kGroup);
    for (int64_t kb = (int64_t)blockIdx.x * (kThreads / kGroup) + (tid >> 5) * 4; kb < n; kb += per_pass) {   // warp-uniform
        const int64_t k = kb + gw;
        const bool have = k < n;
        const uint32_t kind = have ? X.kind[k] : 0u;
        int s = 0; int64_t r = 0; int new_len = 0; uint64_t seq16 = 0, qual16 = 0;
        if (kind) { s = O.mod_session[k]; r = O.mod_read[k]; new_len = (int)O.mod_len[k]; seq16 = O.mod_seq_off16[k]; qual16 = O.mod_qual_off16[k]; }
        const int col_begin = kind ? __ldg(&descs[s].col_begin) : 0;
        GermList germ; germ.e = X.germ + (size_t)s * kGermCap; germ.n = kind ? X.germ_n[s] : 0u;
        if (kind == 1u) {
            const int pos = __ldg(B.pos + r), L = new_len;
            const uint4* rec = reinterpret_cast<const uint4*>(B.seq4 + 16ull * __ldg(B.seq_off16 + r));
            uint4* out = reinterpret_cast<uint4*>(O.out_seq4 + 16ull * seq16);
            const int units = (L + 31) >> 5;
            for (int u = glane; u < units; u += kGroup) {
                const uint4 vv = ldg128(rec + u);
                uint32_t w[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int wi = 4 * u + q;
                    const uint32_t tm = tail_mask(L, wi);
                    uint32_t v = w[q] & tm;
                    if (tm) {
                        const int p0 = pos + 8 * wi;
                        const uint32_t fw = ref_word(B.ref4, (int64_t)p0);
                        uint32_t x = (v ^ fw) & tm;
                        while (x) {
                            const int nb = (__ffs(x) - 1) >> 2;
                            x &= ~(0xfu << (nb * 4));
                            const uint32_t b = (v >> (nb * 4)) & 15u;
                            if (b != 15u && germ(p0 + nb - col_begin, b)) v = (v & ~(0xfu << (nb * 4))) | (((fw >> (nb * 4)) & 15u) << (nb * 4));
                        }
                    }
                    w[q] = v;
                }
                out[u] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else if (kind == 2u) {
            const uint32_t c0 = __ldg(B.cigar_off + r), c1 = __ldg(B.cigar_off + r + 1);
            int units = (new_len + 31) >> 5; if (units < 1) units = 1;
            uint32_t* oseq = reinterpret_cast<uint32_t*>(O.out_seq4 + 16ull * seq16);
            masked_words_g(B, r, __ldg(B.pos + r), new_len, c0, c1, col_begin, units * 4, glane, kGroup, germ, [&](int wd, uint32_t v) { oseq[wd] = v; });
        }
        const bool indel = kind == 3u;
        if (__any_sync(0xffffffffu, indel)) {
            Ed2 E; E.ne = 0; E.n_del = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) { E.irp[q] = 0; E.len[q] = 0; E.pos[q] = 0; E.mean[q] = 0u; }
            int64_t q_lo = 0, q_hi = 0;
            if (indel) {
                const uint4* ap = reinterpret_cast<const uint4*>(O.out_qual + 32ull * qual16);
                const uint4 a0 = ap[0], a1 = ap[1];                    // EditAux written by the session kernel
                E.irp[0] = (int)a0.x; E.pos[0] = (int)a0.y; E.len[0] = (int)(a0.z & 0x7fffffffu);
                E.irp[1] = (int)a0.w; E.pos[1] = (int)a1.x; E.len[1] = (int)(a1.y & 0x7fffffffu);
                E.ne = (int)a1.z; E.n_del = (int)a1.w;
                clamp_edits2(E, (int)(__ldg(B.len_flag + r) & 0xffffu));
                const bool tumor = r < B.n_tumor;
                q_lo = tumor ? descs[s].qt_begin : descs[s].qn_begin; q_hi = tumor ? descs[s].qt_end : descs[s].qn_end;
            }
            __syncwarp();                                             // every lane of the group has read the aux before it is overwritten
            emit_indel_group_t(B, O.totals, O, indel, E, r, col_begin, q_lo, q_hi, stage[group], seq16, qual16, new_len, glane, germ);
        }
    }
}

}  // namespace ga
