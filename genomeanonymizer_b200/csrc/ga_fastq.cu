// ga_fastq.cu - N1 of SURVEY.md 8(f): FASTQ records rendered on the device.
//
// Reference: AnonymizedRead.get_anonymized_fastq_record (anonymizer_methods.py:205-243) + generate_anonymized_read
// (anonymizer_methods.py:57-58) + the writer's trailing newline (short_read_tumor_normal_anonymizer.py:157-158):
//     "@" query_name "/" (1 if READ1 else 2) "\n" SEQ "\n+\n" QUAL "\n"
// Reverse-strand reads are reverse-complemented for output (anonymizer_methods.py:205-214, table at :22); their
// qualities are kept in original-read orientation internally and reversed again at print time
// (anonymizer_methods.py:95, 213), i.e. they are printed in BAM order next to the reverse-complemented sequence
// (quirk Q1) - which is the order ga_reads.qual and ga_result.out_qual already hold.
// The reference only knows A, C, G, T, N (anything else raises TypeError); IUPAC codes are
// complemented here by reversing the 4 bits of the BAM code, which is their IUPAC complement.
//
// Two calls: ga_fastq_layout sizes every record and scans the sizes into byte offsets; ga_fastq_render writes the
// text, one warp per record: characters are produced into shared memory (8 bases per lane through byte permutes, 4
// qualities per lane) with the alignment of their destination and leave with aligned 128-bit stores.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include "ga_engine_internal.h"

namespace ga {

constexpr int kScanBlock = 1024;

struct FastqView {
    const int32_t* read; const int32_t* record; const uint8_t* names; const int64_t* name_off; int64_t n_items;
    // reads
    const uint32_t* len_flag; const uint32_t* seq_off16; const uint8_t* seq4; const uint8_t* qual;
    const int32_t* qual_reads; const uint32_t* qual_off16; int64_t n_qual; int64_t n_reads;
    // records
    const uint32_t* mod_len; const uint32_t* mod_seq_off16; const uint32_t* mod_qual_off16; const uint8_t* out_seq4; const uint8_t* out_qual;
    int64_t n_records;
};

__device__ __forceinline__ int64_t item_bytes(const FastqView& V, int64_t k) {
    const int32_t r = V.read[k], rec = V.record[k];
    const int64_t nl = V.name_off[r + 1] - V.name_off[r];
    const int64_t L = rec >= 0 ? (int64_t)V.mod_len[rec] : (int64_t)(V.len_flag[r] & 0xffffu);
    return nl + 2 * L + 8;                                               // '@' '/' mate '\n' | '\n' '+' '\n' | '\n'
}

// exclusive scan of the record sizes, three small kernels: block sums, scan of the block sums, offsets
__global__ void fastq_block_sums_kernel(FastqView V, int64_t* __restrict__ block_sums) {
    __shared__ long long s_part[kScanBlock / 32];
    const int64_t k = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    long long v = k < V.n_items ? item_bytes(V, k) : 0;
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        long long t = threadIdx.x < kScanBlock / 32 ? s_part[threadIdx.x] : 0;
        for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kScanBlock) fastq_scan_sums_kernel(int64_t* __restrict__ block_sums, int64_t n_blocks, int64_t* __restrict__ total_out) {
    // one CTA: every thread takes a run of consecutive block sums (independent loads, all in flight together), the CTA
    // scans the per-thread totals, and every thread writes the exclusive prefix of its run
    __shared__ long long s_part[kScanBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (n_blocks + kScanBlock - 1) / kScanBlock;
    const int64_t b0 = (int64_t)threadIdx.x * per, b1 = min(n_blocks, b0 + per);
    long long mine = 0;
    for (int64_t b = b0; b < b1; ++b) mine += block_sums[b];
    long long inc = mine;
    for (int d = 1; d < 32; d <<= 1) { const long long n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += n; }
    if (lane == 31) s_part[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const long long v = s_part[lane];
        long long t = v;
        for (int d = 1; d < 32; d <<= 1) { const long long n = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += n; }
        s_part[lane] = t - v;                                            // exclusive prefix of the warps
        if (lane == 31) *total_out = t;
    }
    __syncthreads();
    long long run = s_part[warp] + inc - mine;
    for (int64_t b = b0; b < b1; ++b) { const long long v = block_sums[b]; block_sums[b] = run; run += v; }
}

__global__ void fastq_offsets_kernel(FastqView V, const int64_t* __restrict__ block_sums, int64_t* __restrict__ text_off) {
    __shared__ long long s_part[kScanBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t k = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const long long v = k < V.n_items ? item_bytes(V, k) : 0;
    long long inc = v;
    for (int d = 1; d < 32; d <<= 1) { const long long n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += n; }
    if (lane == 31) s_part[warp] = inc;
    __syncthreads();
    long long before = block_sums[blockIdx.x];
    for (int w = 0; w < warp; ++w) before += s_part[w];
    if (k < V.n_items) text_off[k] = before + inc - v;
}

__device__ __forceinline__ uint8_t base_char(uint32_t code) {
    // "=ACMGRSVTWYHKDBN" as two little-endian 64-bit words
    const unsigned long long w0 = ((unsigned long long)'=') | ((unsigned long long)'A' << 8) | ((unsigned long long)'C' << 16) | ((unsigned long long)'M' << 24) |
                                  ((unsigned long long)'G' << 32) | ((unsigned long long)'R' << 40) | ((unsigned long long)'S' << 48) | ((unsigned long long)'V' << 56);
    const unsigned long long w1 = ((unsigned long long)'T') | ((unsigned long long)'W' << 8) | ((unsigned long long)'Y' << 16) | ((unsigned long long)'H' << 24) |
                                  ((unsigned long long)'K' << 32) | ((unsigned long long)'D' << 40) | ((unsigned long long)'B' << 48) | ((unsigned long long)'N' << 56);
    return (uint8_t)(((code & 8u) ? w1 : w0) >> (8u * (code & 7u)));
}

constexpr int kStageBytes = 544;         // staged text per group of 8 lanes (records of up to ~250 bases)
constexpr int kRenderGroup = 8;          // lanes that render one record

struct FastqItem {                       // everything one record needs, gathered lane-per-record and broadcast by shuffle
    const uint32_t* seqw; const uint8_t* q; const uint8_t* name; int64_t off; int nl, L; uint32_t flag; int ok;
};

__device__ __forceinline__ FastqItem shfl_item(const FastqItem& it, int src) {
    FastqItem o;
    o.seqw = reinterpret_cast<const uint32_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(it.seqw), src));
    o.q = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(it.q), src));
    o.name = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(it.name), src));
    o.off = (int64_t)__shfl_sync(0xffffffffu, (unsigned long long)it.off, src);
    o.nl = __shfl_sync(0xffffffffu, it.nl, src); o.L = __shfl_sync(0xffffffffu, it.L, src);
    o.flag = __shfl_sync(0xffffffffu, it.flag, src); o.ok = __shfl_sync(0xffffffffu, it.ok, src);
    return o;
}

// eight base codes (one 32-bit word of seq4) -> eight characters, six byte permutes
__device__ __forceinline__ void base_chars8(uint32_t cw, uint32_t* ch) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t n16 = (cw >> (16 * h)) & 0xffffu;
        const uint32_t sel = n16 & 0x7777u;
        const uint32_t lo = __byte_perm(0x4d43413du, 0x56535247u, sel);      // "=ACM" "GRSV"
        const uint32_t hi = __byte_perm(0x48595754u, 0x4e42444bu, sel);      // "TWYH" "KDBN"
        ch[h] = __byte_perm(lo, hi, 0x3210u + ((n16 & 0x8888u) >> 1));
    }
}

// A warp takes 32 records at a time: their indices, lengths, offsets and source pointers are gathered lane-per-record
// (three dependent round trips for the whole batch).  Then FOUR records are rendered at a time, a group of 8 lanes
// each: every lane first issues all of its loads (at most four sequence words, eight quality words), so that the
// loads of four records are in flight together, then the characters are produced into the group's shared-memory
// slice with the alignment of their destination and leave with aligned 128-bit stores.
__global__ void __launch_bounds__(256, 4) fastq_render_kernel(FastqView V, const int64_t* __restrict__ text_off, uint8_t* __restrict__ text,
                                                           int64_t text_cap, ga_totals* totals) {
    __shared__ __align__(16) uint8_t stage[8][32 / kRenderGroup][kStageBytes];
    const int lane = threadIdx.x & 31, g = lane / kRenderGroup, gl = lane % kRenderGroup;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t k0 = warp0 * 32; k0 < V.n_items; k0 += n_warps * 32) {
        // ---- lane = record
        FastqItem me; me.seqw = nullptr; me.q = nullptr; me.name = nullptr; me.off = 0; me.nl = 0; me.L = 0; me.flag = 0u; me.ok = 0;
        const int64_t k = k0 + lane;
        if (k < V.n_items) {
            const int32_t r = V.read[k], rec = V.record[k];
            if (r < 0 || r >= V.n_reads || rec >= V.n_records) raise_error(totals, GA_ERR_BAD_ARGUMENT, (uint32_t)k);
            else {
                const int64_t nb = V.name_off[r];
                me.name = V.names + nb; me.nl = (int)(V.name_off[r + 1] - nb);
                const uint32_t lf = V.len_flag[r];
                me.flag = lf >> 16;
                me.L = rec >= 0 ? (int)V.mod_len[rec] : (int)(lf & 0xffffu);
                me.seqw = rec >= 0 ? reinterpret_cast<const uint32_t*>(V.out_seq4 + 16ull * V.mod_seq_off16[rec])
                                   : reinterpret_cast<const uint32_t*>(V.seq4 + 16ull * V.seq_off16[r]);
                // qualities in printed (= BAM) order: the record's own when it carries them, else the read's
                if (rec >= 0 && V.mod_qual_off16[rec] != 0xffffffffu) me.q = V.out_qual + 32ull * V.mod_qual_off16[rec];
                else if (V.qual && !V.qual_reads) me.q = V.qual + 32ull * V.seq_off16[r];
                else if (V.qual) {
                    int64_t b = 0, e = V.n_qual;
                    while (b < e) { const int64_t m = (b + e) >> 1; if (V.qual_reads[m] < r) b = m + 1; else e = m; }
                    if (b < V.n_qual && V.qual_reads[b] == r) me.q = V.qual + 32ull * V.qual_off16[b];
                }
                me.off = text_off[k];
                if (!me.q) raise_error(totals, GA_ERR_BAD_ARGUMENT, (uint32_t)k);
                else if (me.off + me.nl + 2ll * me.L + 8 > text_cap) raise_error(totals, GA_ERR_CAPACITY, (uint32_t)k);
                else me.ok = 1;
            }
        }
        const int n_here = (int)min((int64_t)32, V.n_items - k0);
        // does the record fit a group's slice?  (destination alignment pad + text; at most 32 words of bases, 64 of qualities)
        const int my_pad = (int)(reinterpret_cast<uintptr_t>(text + me.off) & 15u);
        const bool my_fit = me.ok && my_pad + me.nl + 2ll * me.L + 8 + 8 <= (int64_t)kStageBytes && me.L <= 248 && me.nl <= 200;   // + 8: aligned words reach past the record's end
        uint32_t slow_mask = __ballot_sync(0xffffffffu, me.ok && !my_fit);
        const uint32_t fit_mask = __ballot_sync(0xffffffffu, my_fit);
        // ---- four records per step, 8 lanes each
        for (int step = 0; step < 32 / (32 / kRenderGroup); ++step) {
            if (!((fit_mask >> (4 * step)) & 0xfu)) continue;           // warp-uniform
            const int src = 4 * step + g;
            const FastqItem it = shfl_item(me, src);
            const bool act = src < n_here && ((fit_mask >> src) & 1u);
            const int nl = it.nl, L = it.L;
            const bool reverse = (it.flag & 0x10u) != 0u;
            const uint8_t mate = (it.flag & 0x40u) ? '1' : '2';           // anonymizer_methods.py:218
            const int total = nl + 2 * L + 8;
            uint8_t* out = text + it.off;
            const int pad = (int)(reinterpret_cast<uintptr_t>(out) & 15u);   // the staged copy has the alignment of its destination
            // The records of a step usually lie back to back in the text (the layout is a running sum): then the four of
            // them are staged as ONE span with the alignment of its destination, and only the span's first and last
            // 16-byte chunk are shared with neighbours - not those of every record.
            const int64_t off0 = __shfl_sync(0xffffffffu, (unsigned long long)it.off, 0);
            const int64_t prev_end = __shfl_up_sync(0xffffffffu, (unsigned long long)(it.off + total), kRenderGroup);
            const bool chained = act && (g == 0 || prev_end == it.off);
            const bool span = __all_sync(0xffffffffu, chained);
            const int pad0 = (int)(reinterpret_cast<uintptr_t>(text + off0) & 15u);
            const int span_len = (int)(__shfl_sync(0xffffffffu, (unsigned long long)(it.off + total), 31) - off0);
            uint8_t* wb = &stage[threadIdx.x >> 5][0][0];                // the warp's staging area, 16-byte aligned
            uint8_t* sb = wb + g * kStageBytes;
            const int o_sg = span ? pad0 + (int)(it.off - off0) : g * kStageBytes + pad;   // where this record's text begins in it
            uint8_t* sg = wb + o_sg;
            // every load first
            uint8_t nm[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int c = gl + kRenderGroup * t; nm[t] = (act && c >= 1 && c <= nl) ? it.name[c - 1] : (uint8_t)0; }
            uint32_t sw[4], qv[8];
            const int nw = act ? (L + 7) >> 3 : 0, nq = act ? (L + 3) >> 2 : 0;
            const uint32_t* qw = reinterpret_cast<const uint32_t*>(it.q);
#pragma unroll
            for (int t = 0; t < 4; ++t) { const int w = gl + kRenderGroup * t; sw[t] = w < nw ? __ldg(it.seqw + w) : 0u; }
#pragma unroll
            for (int t = 0; t < 8; ++t) { const int w = gl + kRenderGroup * t; qv[t] = w < nq ? __ldg(qw + w) : 0u; }
            // The text is produced as ALIGNED 32-bit words of the staging area: the characters of a section start at an
            // arbitrary byte, so a lane combines its source word with the one before it (held by the lane to its left, or by
            // lane 7 one round earlier) and shifts.  The first and last word of a section also cover up to three bytes of its
            // neighbours (up to seven for the bases): sections are therefore written in an order in which the neighbour comes
            // later - bases, then qualities, then header, separator and final newline - with a warp barrier in between.
            const int o_ss = o_sg + nl + 4, s_seq = o_ss & 3;             // bases
            const int o_sq = o_ss + L + 3, s_q = o_sq & 3;                // qualities
            const int from_left = (lane & ~(kRenderGroup - 1)) | ((gl + kRenderGroup - 1) & (kRenderGroup - 1));
            const int max_L = __reduce_max_sync(0xffffffffu, act ? L : 0);
            {
                // output word pair w covers the characters [8w - s_seq, 8w - s_seq + 8) of the printed sequence R.  Forward reads:
                // R = the codes, the pair is the 32-bit window of the code words (w-1, w) at 4 * (8 - s_seq) bits.  Reverse reads:
                // R[i] = complement(code[L-1-i]) = the bit-reversed window of the words (w-1, w) at 4 * ((L + s_seq) & 7) bits, and
                // it is output pair ((L + s_seq) >> 3) - w.
                uint32_t* dst = reinterpret_cast<uint32_t*>(wb + (o_ss - s_seq));
                const int n_pairs = act ? (L + s_seq + 7) >> 3 : 0;
                const int shift = reverse ? 4 * ((L + s_seq) & 7) : 32 - 4 * s_seq;
                const int top = (L + s_seq) >> 3;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (8 * kRenderGroup * t >= max_L + 11) break;       // warp-uniform
                    const int w = gl + kRenderGroup * t;
                    const uint32_t give = gl == kRenderGroup - 1 ? (t ? sw[t ? t - 1 : 0] : 0u) : sw[t];
                    const uint32_t left = __shfl_sync(0xffffffffu, give, from_left);
                    uint32_t win = reverse ? __funnelshift_r(left, sw[t], shift) : __funnelshift_rc(left, sw[t], shift);
                    if (reverse) win = __brev(win);
                    const int w_out = reverse ? top - w : w;
                    if (w_out >= 0 && w_out < n_pairs) {
                        uint32_t ch[2];
                        base_chars8(win, ch);
                        dst[2 * w_out] = ch[0]; dst[2 * w_out + 1] = ch[1];
                    }
                }
            }
            __syncwarp();
            {
                // output word m covers the qualities [4m - s_q, 4m - s_q + 4): source words (m-1, m) shifted by 8 * (4 - s_q) bits
                uint32_t* dst = reinterpret_cast<uint32_t*>(wb + (o_sq - s_q));
                const int n_words = act ? (L + s_q + 3) >> 2 : 0;
                const int shift = 32 - 8 * s_q;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    if (4 * kRenderGroup * t >= max_L + 7) break;        // warp-uniform
                    const int m = gl + kRenderGroup * t;
                    const uint32_t give = gl == kRenderGroup - 1 ? (t ? qv[t ? t - 1 : 0] : 0u) : qv[t];
                    const uint32_t left = __shfl_sync(0xffffffffu, give, from_left);
                    if (m < n_words) dst[m] = __funnelshift_rc(left, qv[t], shift) + 0x21212121u;   // anonymizer_methods.py:232 (phred <= 93: no carry between bytes)
                }
            }
            __syncwarp();
            if (act) {
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int c = gl + kRenderGroup * t;
                    if (c >= 1 && c <= nl) sg[c] = nm[t];
                }
                for (int c = gl + 4 * kRenderGroup; c <= nl; c += kRenderGroup) sg[c] = it.name[c - 1];
                // the eight fixed characters, one per lane: '@' | '/' mate '\n' | '\n' '+' '\n' | '\n'
                const int at = gl == 0 ? 0 : gl < 4 ? nl + gl : gl < 7 ? nl + 4 + L + (gl - 4) : nl + 4 + 2 * L + 3;
                const uint32_t fixed_lo = (uint32_t)'@' | ((uint32_t)'/' << 8) | ((uint32_t)mate << 16) | ((uint32_t)'\n' << 24);
                const uint32_t fixed_hi = (uint32_t)'\n' | ((uint32_t)'+' << 8) | ((uint32_t)'\n' << 16) | ((uint32_t)'\n' << 24);
                sg[at] = (uint8_t)((gl < 4 ? fixed_lo : fixed_hi) >> (8 * (gl & 3)));
            }
            __syncwarp();
            if (span) {                                                   // warp-uniform: the whole warp copies the span
                const uint8_t* ws = stage[threadIdx.x >> 5][0];
                uint8_t* gb = text + off0 - pad0;                         // 16-byte aligned
                const int end = pad0 + span_len;
                for (int c16 = lane * 16; c16 + 16 <= end; c16 += 32 * 16)
                    if (c16 >= pad0) *reinterpret_cast<uint4*>(gb + c16) = *reinterpret_cast<const uint4*>(ws + c16);
                // the span's first and last 16-byte chunk are shared with the neighbours' bytes: one byte per lane
                // (lanes 0-15: what is left of the first chunk, lanes 16-31: what is left of the last; a byte may be written twice)
                const int tb = lane < 16 ? lane : (end & ~15) + (lane - 16);
                if (tb >= pad0 && tb < end && (lane >= 16 || pad0 != 0)) gb[tb] = ws[tb];
            } else if (act) {
                uint8_t* gb = out - pad;                                  // 16-byte aligned
                const int end = pad + total;
                for (int c16 = gl * 16; c16 < end; c16 += kRenderGroup * 16) {
                    if (c16 >= pad && c16 + 16 <= end) *reinterpret_cast<uint4*>(gb + c16) = *reinterpret_cast<const uint4*>(sb + c16);
                    else for (int t = max(c16, pad); t < min(c16 + 16, end); ++t) gb[t] = sb[t];   // the neighbours' bytes share these chunks
                }
            }
            __syncwarp();
        }
        // ---- general path (records longer than a slice; rare): one record after the other, one character per lane and step
        while (slow_mask) {
            const int j = __ffs(slow_mask) - 1; slow_mask &= slow_mask - 1u;
            const FastqItem it = shfl_item(me, j);
            const int nl = it.nl, L = it.L;
            const bool reverse = (it.flag & 0x10u) != 0u;
            const uint8_t mate = (it.flag & 0x40u) ? '1' : '2';
            uint8_t* out = text + it.off;
            for (int c = lane; c < nl + 4; c += 32)
                out[c] = c == 0 ? (uint8_t)'@' : c <= nl ? it.name[c - 1] : c == nl + 1 ? (uint8_t)'/' : c == nl + 2 ? mate : (uint8_t)'\n';
            uint8_t* os = out + nl + 4;
            for (int jj = lane; jj < L; jj += 32) {
                const int bj = reverse ? L - 1 - jj : jj;
                uint32_t code = (it.seqw[bj >> 3] >> ((bj & 7) * 4)) & 15u;
                if (reverse) code = __brev(code) >> 28;
                os[jj] = base_char(code);
            }
            if (lane < 3) os[L + lane] = lane == 1 ? (uint8_t)'+' : (uint8_t)'\n';
            uint8_t* oq = os + L + 3;
            for (int jj = lane; jj < L; jj += 32) oq[jj] = (uint8_t)(it.q[jj] + 33u);    // anonymizer_methods.py:232
            if (lane == 0) oq[L] = (uint8_t)'\n';
        }
    }
}

}  // namespace ga

static ga::FastqView make_view(const ga_reads* R, const ga_result* O, const ga_fastq_items* I, int64_t n_records) {
    ga::FastqView V;
    V.read = I->read; V.record = I->record; V.names = I->names; V.name_off = I->name_off; V.n_items = I->n_items;
    V.len_flag = R->len_flag; V.seq_off16 = R->seq_off16; V.seq4 = R->seq4; V.qual = R->qual;
    V.qual_reads = R->qual_reads; V.qual_off16 = R->qual_off16; V.n_qual = R->qual_reads ? R->n_qual : 0; V.n_reads = R->n_reads;
    V.mod_len = O ? O->mod_len : nullptr; V.mod_seq_off16 = O ? O->mod_seq_off16 : nullptr; V.mod_qual_off16 = O ? O->mod_qual_off16 : nullptr;
    V.out_seq4 = O ? O->out_seq4 : nullptr; V.out_qual = O ? O->out_qual : nullptr;
    V.n_records = O ? n_records : 0;
    return V;
}

extern "C" {

int ga_fastq_layout(ga_engine* e, const ga_reads* R, const ga_result* O, int64_t n_records, const ga_fastq_items* I,
                    int64_t* text_off, void* stream_) {
    if (!e || !R || !I || !text_off || I->n_items < 0) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_fastq_layout: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    GA_CUDA(cudaSetDevice(e->device));
    if (I->n_items == 0) { GA_CUDA(cudaMemsetAsync(text_off, 0, sizeof(int64_t), st)); return GA_OK; }
    const int64_t n_blocks = (I->n_items + ga::kScanBlock - 1) / ga::kScanBlock;
    if (n_blocks > e->cap_fastq_blocks) {
        cudaFree(e->d_fastq_sums); e->d_fastq_sums = nullptr; e->cap_fastq_blocks = 0;
        GA_CUDA(cudaMalloc(&e->d_fastq_sums, (size_t)(n_blocks + 1024) * sizeof(int64_t)));
        e->cap_fastq_blocks = n_blocks + 1024;
    }
    const ga::FastqView V = make_view(R, O, I, n_records);
    ga::fastq_block_sums_kernel<<<(unsigned)n_blocks, ga::kScanBlock, 0, st>>>(V, e->d_fastq_sums);
    ga::fastq_scan_sums_kernel<<<1, ga::kScanBlock, 0, st>>>(e->d_fastq_sums, n_blocks, text_off + I->n_items);
    ga::fastq_offsets_kernel<<<(unsigned)n_blocks, ga::kScanBlock, 0, st>>>(V, e->d_fastq_sums, text_off);
    e->launches += 3;
    GA_CUDA(cudaGetLastError());
    return GA_OK;
}

int ga_fastq_render(ga_engine* e, const ga_reads* R, const ga_result* O, int64_t n_records, const ga_fastq_items* I,
                    const int64_t* text_off, uint8_t* text, int64_t text_cap, ga_totals* status, void* stream_) {
    if (!e || !R || !I || !text_off || !text || !status || I->n_items < 0) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_fastq_render: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    GA_CUDA(cudaSetDevice(e->device));
    if (I->n_items == 0) return GA_OK;
    const ga::FastqView V = make_view(R, O, I, n_records);
    const int64_t ctas = (I->n_items + 255) / 256;                    // 32 records per warp, 8 warps per CTA
    const unsigned grid = (unsigned)std::min<int64_t>(ctas, (int64_t)e->n_sm * 32);
    ga::fastq_render_kernel<<<grid, 256, 0, st>>>(V, text_off, text, text_cap, status);
    e->launches += 1;
    GA_CUDA(cudaGetLastError());
    return GA_OK;
}

}  // extern "C"
