// ga_wire.cu - what sits next to the masking kernels on the way in and out of the engine:
//   * ga_result_digest: key + 128-bit hash of every modified record of a device-resident result and their
//     order-independent sum (arithmetic of include/ga_digest.h), so that two results - engine vs oracle, one GPU vs
//     eight - are compared record by record without moving the records.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include "ga_engine_internal.h"
#include "../../include/ga_digest.h"

namespace ga {

__device__ __forceinline__ uint32_t clear_nibbles_from(uint32_t w, int keep) {       // keep in 1..8
    return keep >= 8 ? w : (w & ((1u << (4 * keep)) - 1u));
}
__device__ __forceinline__ uint32_t clear_bytes_from(uint32_t w, int keep) {         // keep in 1..4
    return keep >= 4 ? w : (w & ((1u << (8 * keep)) - 1u));
}

// One thread per record: the records are short (20 + 38 words at 150 bp) and this kernel is not on the timed path.
__global__ void digest_kernel(ResultView O, int64_t n_records, ga_digest_ids ids, uint64_t* __restrict__ rec_keys,
                              uint64_t* __restrict__ rec_hash, unsigned long long* __restrict__ digest) {
    unsigned long long s_lo = 0ull, s_hi = 0ull, s_n = 0ull, s_len = 0ull;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_records; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = (int64_t)O.mod_session[k] + ids.session_base;
        const int64_t r = O.mod_read[k];
        const uint64_t gid = r < ids.n_tumor ? (uint64_t)(ids.tumor_base + r) : ((1ull << 40) | (uint64_t)(ids.normal_base + r - ids.n_tumor));
        const uint32_t len = O.mod_len[k];
        const uint32_t qo = O.mod_qual_off16[k];
        const bool has_q = qo != 0xffffffffu;
        uint64_t lo = GA_DIGEST_SEED_LO, hi = GA_DIGEST_SEED_HI;
        auto mix = [&](uint64_t w) { lo = ga_digest_mix(lo, w, GA_DIGEST_MUL_LO); hi = ga_digest_mix(hi, w, GA_DIGEST_MUL_HI); };
        mix((uint64_t)ids.contig); mix((uint64_t)s); mix(gid); mix((uint64_t)len | ((uint64_t)(has_q ? 1u : 0u) << 32));
        const uint32_t* sq = reinterpret_cast<const uint32_t*>(O.out_seq4 + 16ull * O.mod_seq_off16[k]);
        const int nw = ((int)len + 7) >> 3;
        for (int w = 0; w < nw; ++w) mix((uint64_t)clear_nibbles_from(__ldg(sq + w), (int)len - 8 * w));
        if (has_q) {
            const uint32_t* qq = reinterpret_cast<const uint32_t*>(O.out_qual + 32ull * qo);
            const int nq = ((int)len + 3) >> 2;
            for (int w = 0; w < nq; ++w) mix((uint64_t)clear_bytes_from(__ldg(qq + w), (int)len - 4 * w));
        }
        lo = ga_digest_fin(lo); hi = ga_digest_fin(hi);
        if (rec_keys) { rec_keys[2 * k] = (uint64_t)s; rec_keys[2 * k + 1] = gid; }
        if (rec_hash) { rec_hash[2 * k] = lo; rec_hash[2 * k + 1] = hi; }
        s_lo += lo; s_hi += hi; s_n += 1ull; s_len += len;
    }
    for (int d = 16; d; d >>= 1) {
        s_lo += __shfl_xor_sync(0xffffffffu, s_lo, d); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, d);
        s_n += __shfl_xor_sync(0xffffffffu, s_n, d);   s_len += __shfl_xor_sync(0xffffffffu, s_len, d);
    }
    if ((threadIdx.x & 31) == 0 && s_n) {
        atomicAdd(digest + 0, s_lo); atomicAdd(digest + 1, s_hi); atomicAdd(digest + 2, s_n); atomicAdd(digest + 3, s_len);
    }
}

}  // namespace ga

extern "C" int ga_result_digest(ga_engine* e, const ga_result* out, int64_t n_records, const ga_digest_ids* ids,
                                uint64_t* rec_keys, uint64_t* rec_hash, uint64_t* digest, void* stream_) {
    if (!e || !out || !ids || !digest || n_records < 0) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_result_digest: null argument");
    if (n_records > out->cap_records) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_result_digest: more records than the result holds");
    if (n_records == 0) return GA_OK;
    GA_CUDA(cudaSetDevice(e->device));
    ga::ResultView O;
    O.cap_records = out->cap_records; O.cap_seq16 = out->cap_seq16; O.cap_qual16 = out->cap_qual16;
    O.mod_session = out->mod_session; O.mod_read = out->mod_read; O.mod_len = out->mod_len;
    O.mod_seq_off16 = out->mod_seq_off16; O.mod_qual_off16 = out->mod_qual_off16;
    O.out_seq4 = out->out_seq4; O.out_qual = out->out_qual; O.sess_counts = out->sess_counts; O.totals = out->totals;
    const int threads = 128;
    const int grid = (int)std::min<int64_t>((n_records + threads - 1) / threads, (int64_t)e->n_sm * 32);
    ga::digest_kernel<<<grid, threads, 0, (cudaStream_t)stream_>>>(O, n_records, *ids, rec_keys, rec_hash,
                                                                  reinterpret_cast<unsigned long long*>(digest));
    e->launches++;
    GA_CUDA(cudaGetLastError());
    return GA_OK;
}

// =====================================================================================================================
// Wire form (include/ga_wire.h): host packer and device expansion.
#include <string.h>
#include <thread>
#include <vector>
#include "../../include/ga_wire.h"

namespace {

inline size_t align4(size_t n) { return (n + 3) & ~(size_t)3; }

struct BlockPlan { int64_t r0; uint32_t n, n_words, n_gen, n_gen_ops, n_exc, units, ops, len_common, n_lenx, n_posx, n_flags; uint16_t fdict[16]; uint64_t bytes; };

inline bool is_generic(const ga_reads* R, int64_t r, uint32_t L) {
    const uint32_t c0 = R->cigar_off[r], c1 = R->cigar_off[r + 1];
    return !(c1 - c0 == 1u && (R->cigar[c0] & 15u) == 0u && (R->cigar[c0] >> 4) == L);
}
inline uint32_t units_of_len(uint32_t L) { return L ? (L + 31u) / 32u : 1u; }
inline bool plain_code(uint32_t c) { return c == 1u || c == 2u || c == 4u || c == 8u; }

// Sixteen base codes at a time (one 64-bit word of seq4, low nibble first).  not_plain: bit 4k set when nibble k is not one
// of A C G T (1 2 4 8); two_bits: the sixteen 2-bit codes (A C G T = 0 1 2 3, anything else 0) packed into 32 bits.
constexpr uint64_t kNib1 = 0x1111111111111111ull;
inline uint64_t load_nibbles16(const uint8_t* rec, uint32_t q, uint32_t L) {   // bases [q, q + 16) of a record of L bases, zero behind the read
    uint64_t w;
    memcpy(&w, rec + (q >> 1), 8);                                      // records are whole 16-byte units: the load stays inside
    const uint32_t left = L - q;
    return left >= 16u ? w : (w & ((1ull << (4u * left)) - 1ull));
}
inline uint64_t not_plain16(uint64_t w) {
    const uint64_t sum = (w & kNib1) + ((w >> 1) & kNib1) + ((w >> 2) & kNib1) + ((w >> 3) & kNib1);   // set bits per nibble, 0 .. 4
    const uint64_t t = sum ^ kNib1;                                      // zero exactly where one bit is set
    return (t | (t >> 1) | (t >> 2)) & kNib1;
}
inline uint32_t two_bits16(uint64_t w, uint64_t not_plain) {
    const uint64_t ok = ~not_plain & kNib1;
    const uint64_t lo = ((w >> 1) | (w >> 3)) & ok, hi = ((w >> 2) | (w >> 3)) & ok;
    uint64_t x = lo | (hi << 1);                                         // a 2-bit code in the low half of every nibble
    x = (x | (x >> 2)) & 0x0f0f0f0f0f0f0f0full;
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)x;
}

// Block boundaries of one dataset: at most GA_WIRE_BLOCK_READS reads, position differences below 65,536.
bool cut_blocks(const ga_reads* R, int64_t lo, int64_t hi, std::vector<BlockPlan>& out) {
    int64_t r = lo;
    while (r < hi) {
        BlockPlan b; memset(&b, 0, sizeof b);
        b.r0 = r;
        int64_t k = r + 1;
        while (k < hi && k - r < GA_WIRE_BLOCK_READS) {
            const int64_t d = (int64_t)R->pos[k] - R->pos[k - 1];
            if (d < 0) return false;
            if (d > 65535) break;
            ++k;
        }
        b.n = (uint32_t)(k - r);
        out.push_back(b);
        r = k;
    }
    return true;
}

// The length most reads of the block have (ties: the smaller one); the others are listed as exceptions.
uint32_t common_length(const ga_reads* R, const BlockPlan& b) {
    std::vector<uint16_t> ls(b.n);
    for (uint32_t i = 0; i < b.n; ++i) ls[i] = (uint16_t)(R->len_flag[b.r0 + i] & 0xffffu);
    std::sort(ls.begin(), ls.end());
    uint32_t best = ls.empty() ? 0u : ls[0], best_n = 0;
    for (size_t i = 0; i < ls.size();) {
        size_t j = i;
        while (j < ls.size() && ls[j] == ls[i]) ++j;
        if (j - i > best_n) { best_n = (uint32_t)(j - i); best = ls[i]; }
        i = j;
    }
    return best;
}

void size_block(const ga_reads* R, BlockPlan& b) {
    b.n_words = b.n_gen = b.n_gen_ops = b.n_exc = b.units = b.ops = 0;
    b.len_common = common_length(R, b);
    b.n_lenx = 0;
    uint64_t n_bases = 0;
    for (int64_t r = b.r0; r < b.r0 + b.n; ++r) {
        const uint32_t L = R->len_flag[r] & 0xffffu;
        n_bases += L;
        if (L != b.len_common) ++b.n_lenx;
        b.units += units_of_len(L);
        const uint32_t nops = R->cigar_off[r + 1] - R->cigar_off[r];
        b.ops += nops;
        if (is_generic(R, r, L)) { ++b.n_gen; b.n_gen_ops += nops; }
        const uint8_t* rec = R->seq4 + 16ull * R->seq_off16[r];
        for (uint32_t q = 0; q < L; q += 16) {
            const uint32_t left = L - q;
            uint64_t np = not_plain16(load_nibbles16(rec, q, L));
            if (left < 16u) np &= (1ull << (4u * left)) - 1ull;            // the zero nibbles behind the read are not bases
            b.n_exc += (uint32_t)__builtin_popcountll(np);
        }
    }
    b.n_words = (uint32_t)((n_bases + 15) / 16) + 3;                      // three zero words behind: the expansion reads two words ahead of any base
    // flags: a dictionary of up to 16 values and 4-bit indices when the block has no more distinct flags (the usual case), else 16 bits each
    b.n_flags = 0; b.n_posx = 0;
    memset(b.fdict, 0, sizeof b.fdict);
    bool dict_ok = true;
    for (int64_t r = b.r0; r < b.r0 + b.n; ++r) {
        const uint16_t f = (uint16_t)(R->len_flag[r] >> 16);
        uint32_t k = 0;
        while (k < b.n_flags && b.fdict[k] != f) ++k;
        if (k == b.n_flags && dict_ok) { if (b.n_flags < 16) b.fdict[b.n_flags++] = f; else dict_ok = false; }
        if (r > b.r0 && (int64_t)R->pos[r] - R->pos[r - 1] >= 255) ++b.n_posx;
    }
    if (!dict_ok) b.n_flags = 0;
    const size_t flag_bytes = b.n_flags ? 32 + align4(((size_t)b.n + 1) / 2) : align4(2 * (size_t)b.n);
    size_t bytes = 32 + flag_bytes + align4((size_t)b.n) + 4 * (size_t)b.n_posx + 4 * (size_t)b.n_lenx + align4(2 * (size_t)b.n_gen) + 4 * ((size_t)b.n_gen + 1) +
                   4 * (size_t)b.n_gen_ops + 4 * (size_t)b.n_exc + 4 * (size_t)b.n_words;
    b.bytes = (bytes + 15) & ~(size_t)15;
}

void write_block(const ga_reads* R, const BlockPlan& b, uint8_t* dst) {
    memset(dst, 0, b.bytes);
    uint32_t* hd = reinterpret_cast<uint32_t*>(dst);
    hd[0] = b.n; hd[1] = b.n_words; hd[2] = b.n_gen; hd[3] = b.n_gen_ops; hd[4] = b.n_exc; hd[5] = b.len_common | (b.n_flags << 16); hd[6] = b.n_lenx; hd[7] = b.n_posx;
    uint8_t* at_p = dst + 32;
    uint16_t* fdict = nullptr; uint8_t* fidx = nullptr; uint16_t* flag = nullptr;
    if (b.n_flags) { fdict = reinterpret_cast<uint16_t*>(at_p); at_p += 32; fidx = at_p; at_p += align4(((size_t)b.n + 1) / 2); memcpy(fdict, b.fdict, 32); }
    else { flag = reinterpret_cast<uint16_t*>(at_p); at_p += align4(2 * (size_t)b.n); }
    uint8_t* dpos8 = at_p; at_p += align4((size_t)b.n);
    uint32_t* posx = reinterpret_cast<uint32_t*>(at_p);
    uint32_t* lenx = posx + b.n_posx;
    uint16_t* gen_idx = reinterpret_cast<uint16_t*>(lenx + b.n_lenx);
    uint32_t* gen_off = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(gen_idx) + align4(2 * (size_t)b.n_gen));
    uint32_t* gen_cig = gen_off + b.n_gen + 1;
    uint32_t* exc = gen_cig + b.n_gen_ops;
    uint32_t* bases = exc + b.n_exc;
    uint32_t npx = 0;
    uint32_t ng = 0, ngo = 0, ne = 0, nx = 0;
    uint64_t at = 0;                                                      // bases of the block so far: the reads lie back to back
    for (uint32_t i = 0; i < b.n; ++i) {
        const int64_t r = b.r0 + i;
        const uint32_t L = R->len_flag[r] & 0xffffu;
        const uint16_t f = (uint16_t)(R->len_flag[r] >> 16);
        if (b.n_flags) { uint32_t k = 0; while (b.fdict[k] != f) ++k; fidx[i >> 1] |= (uint8_t)(k << (4 * (i & 1))); }
        else flag[i] = f;
        if (L != b.len_common) lenx[nx++] = (i << 16) | L;
        const uint32_t dlt = i ? (uint32_t)(R->pos[r] - R->pos[r - 1]) : 0u;
        dpos8[i] = (uint8_t)(dlt < 255u ? dlt : 255u);
        if (dlt >= 255u) posx[npx++] = (i << 16) | dlt;
        if (is_generic(R, r, L)) {
            gen_idx[ng] = (uint16_t)i; gen_off[ng] = ngo; ++ng;
            for (uint32_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; ++c) gen_cig[ngo++] = R->cigar[c];
        }
        const uint8_t* rec = R->seq4 + 16ull * R->seq_off16[r];
        for (uint32_t q = 0; q < L; q += 16) {                             // sixteen bases per step into the block's continuous 2-bit stream
            const uint32_t left = L - q, take = left < 16u ? left : 16u;
            const uint64_t w = load_nibbles16(rec, q, L);
            uint64_t np = not_plain16(w);
            if (left < 16u) np &= (1ull << (4u * left)) - 1ull;
            const uint64_t two = two_bits16(w, np);                        // zero behind the read (the nibbles there are zero: not plain)
            const uint32_t word = (uint32_t)(at >> 4), sh = 2u * (uint32_t)(at & 15);
            bases[word] |= (uint32_t)(two << sh);
            if (sh) bases[word + 1] |= (uint32_t)(two >> (32u - sh));     // n_words has three spare words behind the last base
            while (np) {                                                 // the rare other codes, in read order
                const uint32_t k = (uint32_t)__builtin_ctzll(np) >> 2;
                np &= np - 1ull;
                exc[ne++] = (i << 20) | ((q + k) << 4) | (uint32_t)((w >> (4u * k)) & 15u);
            }
            at += take;
        }
    }
    gen_off[ng] = ngo;
}

template <class F> void parallel_blocks(size_t n, int n_threads, F&& f) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = n_threads > 0 ? (size_t)n_threads : (hw ? hw : 1);
    nt = std::min(nt, std::max<size_t>(1, n));
    if (nt <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; ++t) th.emplace_back([&, t]() { for (size_t i = t; i < n; i += nt) f(i); });
    for (auto& x : th) x.join();
}

int plan_wire(const ga_reads* R, int n_threads, std::vector<BlockPlan>& blocks, int64_t* n_tumor_blocks, int32_t* maxspan) {
    if (!R || R->n_reads < 0 || R->n_tumor < 0 || R->n_tumor > R->n_reads || R->n_reads > 0x7fffffffll) return GA_ERR_BAD_ARGUMENT;
    if (R->n_reads && (!R->pos || !R->len_flag || !R->seq_off16 || !R->cigar_off || !R->cigar || !R->seq4)) return GA_ERR_BAD_ARGUMENT;
    if (!cut_blocks(R, 0, R->n_tumor, blocks)) return GA_ERR_BAD_ARGUMENT;
    *n_tumor_blocks = (int64_t)blocks.size();
    if (!cut_blocks(R, R->n_tumor, R->n_reads, blocks)) return GA_ERR_BAD_ARGUMENT;
    parallel_blocks(blocks.size(), n_threads, [&](size_t i) { size_block(R, blocks[i]); });
    if (maxspan) {
        int32_t m = R->max_ref_span;
        if (m <= 0) {
            m = 1;
            for (int64_t r = 0; r < R->n_reads; ++r) {
                int32_t sp = 0;
                for (uint32_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; ++c) {
                    const uint32_t op = R->cigar[c] & 15u;
                    if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) sp += (int32_t)(R->cigar[c] >> 4);
                }
                m = std::max(m, sp);
            }
        }
        *maxspan = m;
    }
    return GA_OK;
}

}  // namespace

extern "C" int ga_wire_pack_sizes(const ga_reads* R, int64_t* n_blocks, int64_t* n_tumor_blocks, int64_t* blob_bytes, int32_t* max_ref_span) {
    if (!n_blocks || !n_tumor_blocks || !blob_bytes) return GA_ERR_BAD_ARGUMENT;
    std::vector<BlockPlan> blocks;
    const int rc = plan_wire(R, 0, blocks, n_tumor_blocks, max_ref_span);
    if (rc) return rc;
    uint64_t bytes = 0;
    for (const BlockPlan& b : blocks) bytes += b.bytes;
    *n_blocks = (int64_t)blocks.size(); *blob_bytes = (int64_t)bytes;
    return GA_OK;
}

extern "C" int ga_wire_pack(const ga_reads* R, uint8_t* blob, int64_t blob_bytes, ga_wire_dir* dir, int64_t n_blocks, int n_threads) {
    if (!blob && blob_bytes > 0) return GA_ERR_BAD_ARGUMENT;
    if (!dir || (reinterpret_cast<uintptr_t>(blob) & 15u)) return GA_ERR_BAD_ARGUMENT;
    std::vector<BlockPlan> blocks;
    int64_t ntb = 0;
    const int rc = plan_wire(R, n_threads, blocks, &ntb, nullptr);
    if (rc) return rc;
    if ((int64_t)blocks.size() != n_blocks) return GA_ERR_BAD_ARGUMENT;
    uint64_t byte = 0, unit = 0, ops = 0;
    for (size_t i = 0; i < blocks.size(); ++i) {
        memset(&dir[i], 0, sizeof dir[i]);
        dir[i].byte = byte; dir[i].read = (uint32_t)blocks[i].r0; dir[i].unit = (uint32_t)unit; dir[i].ops = (uint32_t)ops; dir[i].pos = R->pos[blocks[i].r0];
        byte += blocks[i].bytes; unit += blocks[i].units; ops += blocks[i].ops;
    }
    if ((int64_t)byte != blob_bytes || unit > 0xffffffffull || ops > 0xffffffffull) return GA_ERR_BAD_ARGUMENT;
    memset(&dir[n_blocks], 0, sizeof dir[0]);
    dir[n_blocks].byte = byte; dir[n_blocks].read = (uint32_t)R->n_reads; dir[n_blocks].unit = (uint32_t)unit; dir[n_blocks].ops = (uint32_t)ops;
    dir[n_blocks].pos = 0x7fffffff;
    parallel_blocks(blocks.size(), n_threads, [&](size_t i) { write_block(R, blocks[i], blob + dir[i].byte); });
    return GA_OK;
}

// ---------------------------------------------------------------------------------------------------- device expansion
namespace ga {

__device__ __forceinline__ uint32_t wire_nibbles(uint32_t h16) {         // 8 two-bit codes -> 8 one-hot BAM nibbles
    uint32_t t = h16 & 0xffffu;
    t = (t | (t << 8)) & 0x00ff00ffu; t = (t | (t << 4)) & 0x0f0f0f0fu; t = (t | (t << 2)) & 0x33333333u;
    const uint32_t b0 = t & 0x11111111u, b1 = (t >> 1) & 0x11111111u;
    return (~(b0 | b1) & 0x11111111u) | ((b0 & ~b1) << 1) | ((b1 & ~b0) << 2) | ((b0 & b1) << 3);
}
__device__ __forceinline__ uint32_t keep_nibbles(uint32_t w, int n) {   // the first n nibbles of w (n may be <= 0 or >= 8)
    return n >= 8 ? w : (n <= 0 ? 0u : (w & (0xffffffffu >> ((8 - n) * 4))));
}

constexpr int kWireThreads = 256;
constexpr int kWirePer = GA_WIRE_BLOCK_READS / kWireThreads;             // reads per thread

// One CTA expands one block into the chunk's ga_reads arrays (chunk-local offsets; tumor slice then normal slice).
__global__ void __launch_bounds__(kWireThreads) wire_expand_kernel(const uint8_t* __restrict__ blob, const ga_wire_dir* __restrict__ dir, int nb_t, int nb_all,
        uint64_t byte0_t, uint64_t byte0_n, uint64_t place_n, uint32_t read0_t, uint32_t read0_n, uint32_t n_t, uint32_t unit0_t, uint32_t unit0_n, uint32_t units_t,
        uint32_t ops0_t, uint32_t ops0_n, uint32_t ops_t, uint32_t n_all, uint32_t ops_all,
        int32_t* __restrict__ pos, uint32_t* __restrict__ len_flag, uint32_t* __restrict__ seq_off16, uint32_t* __restrict__ cigar_off,
        uint32_t* __restrict__ cigar, uint8_t* __restrict__ seq4) {
    __shared__ int16_t s_gen[GA_WIRE_BLOCK_READS];
    __shared__ uint16_t s_len[GA_WIRE_BLOCK_READS];
    __shared__ uint16_t s_dpos[GA_WIRE_BLOCK_READS];
    __shared__ uint32_t s_seqoff[GA_WIRE_BLOCK_READS];
    __shared__ uint32_t s_part[4][kWireThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ds = b >= nb_t;
    const ga_wire_dir d = dir[b];
    const uint8_t* blk = blob + (ds ? place_n + (d.byte - byte0_n) : d.byte - byte0_t);
    const uint32_t r_base = ds ? n_t + (d.read - read0_n) : d.read - read0_t;
    const uint32_t u_base = ds ? units_t + (d.unit - unit0_n) : d.unit - unit0_t;
    const uint32_t o_base = ds ? ops_t + (d.ops - ops0_n) : d.ops - ops0_t;
    const uint32_t* hd = reinterpret_cast<const uint32_t*>(blk);
    const uint32_t n = hd[0], n_words = hd[1], n_gen = hd[2], n_gen_ops = hd[3], n_exc = hd[4], len_common = hd[5] & 0xffffu, n_flags = hd[5] >> 16, n_lenx = hd[6], n_posx = hd[7];
    const uint8_t* at_p = blk + 32;
    const uint16_t* fdict = nullptr; const uint8_t* fidx = nullptr; const uint16_t* flag = nullptr;
    if (n_flags) { fdict = reinterpret_cast<const uint16_t*>(at_p); at_p += 32; fidx = at_p; at_p += (((size_t)n + 1) / 2 + 3) & ~(size_t)3; }
    else { flag = reinterpret_cast<const uint16_t*>(at_p); at_p += (2ull * n + 3) & ~3ull; }
    const uint8_t* dpos8 = at_p; at_p += ((size_t)n + 3) & ~(size_t)3;
    const uint32_t* posx = reinterpret_cast<const uint32_t*>(at_p);
    const uint32_t* lenx = posx + n_posx;
    const uint16_t* gen_idx = reinterpret_cast<const uint16_t*>(lenx + n_lenx);
    const uint32_t* gen_off = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(gen_idx) + ((2ull * n_gen + 3) & ~3ull));
    const uint32_t* gen_cig = gen_off + n_gen + 1;
    const uint32_t* exc = gen_cig + n_gen_ops;
    const uint32_t* bases = exc + n_exc;

    for (uint32_t i = tid; i < n; i += kWireThreads) { s_gen[i] = (int16_t)-1; s_len[i] = (uint16_t)len_common; s_dpos[i] = (uint16_t)dpos8[i]; }
    __syncthreads();
    for (uint32_t j = tid; j < n_gen; j += kWireThreads) s_gen[gen_idx[j]] = (int16_t)j;
    for (uint32_t j = tid; j < n_lenx; j += kWireThreads) { const uint32_t x = lenx[j]; s_len[x >> 16] = (uint16_t)(x & 0xffffu); }
    for (uint32_t j = tid; j < n_posx; j += kWireThreads) { const uint32_t x = posx[j]; s_dpos[x >> 16] = (uint16_t)(x & 0xffffu); }
    __syncthreads();

    // per-thread: kWirePer consecutive reads
    uint32_t Lk[kWirePer], lfk[kWirePer], nops[kWirePer], dp[kWirePer];
    uint32_t su = 0, so = 0, sw = 0, sp = 0;
#pragma unroll
    for (int k = 0; k < kWirePer; ++k) {
        const uint32_t i = (uint32_t)tid * kWirePer + k;
        Lk[k] = i < n ? (uint32_t)s_len[i] : 0u; dp[k] = i < n ? (uint32_t)s_dpos[i] : 0u;
        lfk[k] = 0u;
        if (i < n) lfk[k] = ((uint32_t)(n_flags ? fdict[(fidx[i >> 1] >> (4 * (i & 1))) & 15u] : flag[i]) << 16) | Lk[k];
        nops[k] = 0u;
        if (i < n) { const int j = s_gen[i]; nops[k] = j >= 0 ? gen_off[j + 1] - gen_off[j] : 1u; }
        su += i < n ? (Lk[k] ? (Lk[k] + 31u) >> 5 : 1u) : 0u; so += nops[k]; sw += Lk[k]; sp += dp[k];     // sw: bases (the reads lie back to back in bases2)
    }
    // block-wide exclusive scans of the four per-thread sums
    uint32_t v[4] = {su, so, sw, sp}, ex[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t inc = v[q];
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, dd); if (lane >= dd) inc += t; }
        if (lane == 31) s_part[q][warp] = inc;
        ex[q] = inc - v[q];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) { uint32_t before = 0u; for (int w = 0; w < warp; ++w) before += s_part[q][w]; ex[q] += before; }

    uint32_t u = u_base + ex[0], o = o_base + ex[1], w = ex[2], p = (uint32_t)d.pos + ex[3];
#pragma unroll
    for (int k = 0; k < kWirePer; ++k) {
        const uint32_t i = (uint32_t)tid * kWirePer + k;
        if (i >= n) break;
        const uint32_t L = Lk[k], r = r_base + i;
        p += dp[k];
        pos[r] = (int32_t)p; len_flag[r] = lfk[k]; seq_off16[r] = u; cigar_off[r] = o;
        s_seqoff[i] = u;
        const int j = s_gen[i];
        if (j < 0) cigar[o] = L << 4;                                    // one M op spanning the read
        else { const uint32_t g0 = gen_off[j]; for (uint32_t c = 0; c < nops[k]; ++c) cigar[o + c] = gen_cig[g0 + c]; }
        const uint32_t nu = L ? (L + 31u) >> 5 : 1u;
        uint4* dst = reinterpret_cast<uint4*>(seq4 + 16ull * u);
        const uint32_t sh = (w & 15u) * 2u;                              // the read starts anywhere in a word of 16 bases
        for (uint32_t q = 0; q < nu; ++q) {
            const uint32_t i0 = min((w + 32u * q) >> 4, n_words - 3u);    // (zero words close the stream)
            const uint32_t a = bases[i0], b1 = bases[i0 + 1], c = bases[i0 + 2];
            const uint32_t w0 = __funnelshift_r(a, b1, sh), w1 = __funnelshift_r(b1, c, sh);
            const int left = (int)L - 32 * (int)q;                       // bases behind the read belong to the next one: cut
            dst[q] = make_uint4(keep_nibbles(wire_nibbles(w0), left), keep_nibbles(wire_nibbles(w0 >> 16), left - 8),
                                keep_nibbles(wire_nibbles(w1), left - 16), keep_nibbles(wire_nibbles(w1 >> 16), left - 24));
        }
        u += nu; o += nops[k]; w += L;
    }
    if (b == nb_all - 1 && tid == 0) cigar_off[n_all] = ops_all;
    __syncthreads();                                                      // the records of this block are written
    // base codes other than A C G T
    for (uint32_t x = tid; x < n_exc; x += kWireThreads) {
        const uint32_t e = exc[x], i = e >> 20, q = (e >> 4) & 0xffffu, code = e & 15u;
        uint32_t* word = reinterpret_cast<uint32_t*>(seq4 + 16ull * s_seqoff[i]) + (q >> 3);
        const uint32_t sh = (q & 7u) * 4u;
        atomicAnd(word, ~(0xfu << sh));
        atomicOr(word, code << sh);
    }
}

}  // namespace ga

int ga_wire_expand(ga_engine* e, cudaStream_t st, const uint8_t* d_blob, const ga_wire_dir* d_dir, int n_blocks_t, int n_blocks_all,
                   uint64_t byte0_t, uint64_t byte0_n, uint64_t place_n, uint32_t read0_t, uint32_t read0_n, uint32_t n_t,
                   uint32_t unit0_t, uint32_t unit0_n, uint32_t units_t, uint32_t ops0_t, uint32_t ops0_n, uint32_t ops_t, uint32_t n_all, uint32_t ops_all,
                   int32_t* pos, uint32_t* len_flag, uint32_t* seq_off16, uint32_t* cigar_off, uint32_t* cigar, uint8_t* seq4) {
    if (n_blocks_all <= 0) return GA_OK;
    ga::wire_expand_kernel<<<n_blocks_all, ga::kWireThreads, 0, st>>>(d_blob, d_dir, n_blocks_t, n_blocks_all, byte0_t, byte0_n, place_n, read0_t, read0_n, n_t,
        unit0_t, unit0_n, units_t, ops0_t, ops0_n, ops_t, n_all, ops_all, pos, len_flag, seq_off16, cigar_off, cigar, seq4);
    e->launches++;
    GA_CUDA(cudaGetLastError());
    return GA_OK;
}
