// ga_synth.cu - counter-based synthetic tumor/normal session generator (benchmark and test input).
//
// Not part of the masking path: it only manufactures inputs of the shapes BASELINE.json names
// (SURVEY.md 8(d) "Synthetic generator").  Every value is a pure function of (seed, coordinates), so a
// shard of windows can be generated on any GPU - or, for small sizes, on the host through the *_host
// twins used by the CPU tests - and comes out bit-identical.
//
// Model: one session per somatic SNV at the window centre (get_windows, SR.py:71-131: SNV window =
// [pos-1000, pos+1001) in 0-based coordinates of the 1-based VCF pos); per window and dataset a fixed
// number of reads with stratified sorted starts over [first-L+1, last); germline SNPs / insertions /
// deletions shared by tumor and normal on two haplotypes (het or hom); tumor-only somatic allele at VAF;
// substitution errors, N bases, soft clips.  No H/N/P ops, no unmapped / secondary / supplementary
// records, only ACGTN.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/ga_synth.h"

namespace gs {

#define GS_HD __host__ __device__ __forceinline__

GS_HD uint64_t mix(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
GS_HD uint64_t h2(uint64_t seed, uint64_t a) { return mix(seed ^ mix(a)); }
GS_HD uint64_t h3(uint64_t seed, uint64_t a, uint64_t b) { return mix(mix(seed ^ mix(a)) ^ (b * 0xd6e8feb86659fd93ull)); }
GS_HD float unif(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

constexpr uint64_t kRef = 0x1001, kGerm = 0x1002, kSom = 0x1003, kWin = 0x1004, kRead = 0x2001, kIns = 0x1005, kDepth = 0x1006;

// base index 0..3 = A C G T
GS_HD int ref_idx(const ga_synth_params& P, int64_t p) { return (int)(h2(P.seed ^ kRef, (uint64_t)p) & 3u); }
GS_HD uint32_t idx_code(int i) { return 1u << i; }                       // BAM nibble: A=1 C=2 G=4 T=8

struct Geo {
    int32_t stride, jitter_span, margin;
    int32_t n_per[2];
    int32_t W;             // width of the start-position range
    int32_t units;
};

GS_HD int32_t round_pos(float x) { return (int32_t)(x + 0.5f); }

__host__ __device__ inline Geo geometry(const ga_synth_params& P) {
    Geo g;
    g.margin = 2 * P.window_half + P.read_len + 64;
    const int64_t usable = P.contig_len - 2ll * g.margin;
    g.stride = (int32_t)(usable / (P.total_windows > 0 ? P.total_windows : 1));
    const int32_t min_spacing = 2 * P.window_half + 2 * P.read_len + 64;   // windows never share a read
    g.jitter_span = g.stride - min_spacing;
    if (g.jitter_span < 1) g.jitter_span = 1;
    g.W = 2 * P.window_half + 1 + P.read_len - 1;
    g.n_per[0] = round_pos(P.cov_tumor * (float)g.W / (float)P.read_len);
    g.n_per[1] = round_pos(P.cov_normal * (float)g.W / (float)P.read_len);
    g.units = (P.read_len + 31) / 32;
    if (g.units < 1) g.units = 1;
    return g;
}

// 0-based position of the somatic SNV of global window w
GS_HD int64_t window_pos(const ga_synth_params& P, const Geo& g, int64_t w) {
    return (int64_t)g.margin + w * (int64_t)g.stride + (int64_t)(h2(P.seed ^ kWin, (uint64_t)w) % (uint64_t)g.jitter_span);
}
GS_HD int som_alt_idx(const ga_synth_params& P, int64_t pw) {
    return (ref_idx(P, pw) + 1 + (int)(h2(P.seed ^ kSom, (uint64_t)pw) % 3u)) & 3;
}

struct Var { int kind, len; uint32_t hap; uint64_t h; };   // kind 0 none, 1 SNP, 2 INS, 3 DEL

GS_HD Var germline(const ga_synth_params& P, int64_t p) {
    Var v; v.kind = 0; v.len = 0; v.hap = 0;
    const uint64_t h = h2(P.seed ^ kGerm, (uint64_t)p);
    v.h = h;
    const float u = unif(h);
    if (u >= P.snp_rate + P.indel_rate) return v;
    const uint32_t hv = (uint32_t)(h & 0xffu);
    v.hap = (hv < 154u) ? (1u + ((uint32_t)(h >> 8) & 1u)) : 3u;        // 60 % heterozygous
    if (u < P.snp_rate) { v.kind = 1; v.len = 1; return v; }
    v.len = 1 + (int)((h >> 12) % (uint64_t)(P.max_indel > 0 ? P.max_indel : 1));
    v.kind = ((h >> 9) & 1u) ? 2 : 3;
    return v;
}

// Sinks: Count only tallies, Fill writes the record.
struct Sink {
    // outputs (Fill)
    uint32_t* seq_words;   // record start (null = count only)
    uint32_t* cigar;       // first op slot
    uint8_t* qual;         // quality record or null
    // state
    uint32_t acc; int n_bases; int n_ops; uint32_t cur_op; int cur_len; int has_indel; int ref_span;
    uint64_t rid; float err_rate, n_rate;
};

GS_HD void sink_flush_op(Sink& s) {
    if (s.cur_len > 0) {
        if (s.cigar) s.cigar[s.n_ops] = ((uint32_t)s.cur_len << 4) | s.cur_op;
        ++s.n_ops;
        if (s.cur_op == 1u || s.cur_op == 2u) s.has_indel = 1;
        if (s.cur_op == 0u || s.cur_op == 2u) s.ref_span += s.cur_len;
    }
    s.cur_len = 0;
}
GS_HD void sink_op(Sink& s, uint32_t op, int len) {
    if (len <= 0) return;
    if (s.cur_len > 0 && s.cur_op == op) { s.cur_len += len; return; }
    sink_flush_op(s);
    s.cur_op = op; s.cur_len = len;
}
// push one base (index 0..3) with sequencing error applied
GS_HD void sink_base(Sink& s, int idx) {
    const uint64_t h = h3(s.rid, 0xE44, (uint64_t)s.n_bases);
    const float u = unif(h);
    uint32_t code;
    if (u < s.n_rate) code = 15u;
    else if (u < s.n_rate + s.err_rate) code = idx_code((idx + 1 + (int)((h >> 8) % 3u)) & 3);
    else code = idx_code(idx);
    if (s.seq_words) {
        s.acc |= code << ((s.n_bases & 7) * 4);
        if ((s.n_bases & 7) == 7) { s.seq_words[s.n_bases >> 3] = s.acc; s.acc = 0; }
    }
    ++s.n_bases;
}

// Builds read i of (dataset ds, global window w).  Returns the 0-based reference_start.
template <bool FILL>
__host__ __device__ inline int32_t build_read(const ga_synth_params& P, const Geo& g, int ds, int64_t w, int i, Sink& s, uint32_t* flag_out) {
    const int L = P.read_len;
    const uint64_t rid = h3(P.seed ^ kRead, (uint64_t)w * 2u + (uint64_t)ds, (uint64_t)i);
    s.rid = rid; s.err_rate = P.err_rate; s.n_rate = P.n_rate;
    s.acc = 0; s.n_bases = 0; s.n_ops = 0; s.cur_op = 0; s.cur_len = 0; s.has_indel = 0; s.ref_span = 0;
    const int64_t pw = window_pos(P, g, w);
    const int64_t first = pw + 1 - P.window_half;
    const int64_t lo = first - L + 1;
    const int n = g.n_per[ds];
    const float ju = unif(h2(rid, 1));
    int64_t off = (int64_t)(((double)i + (double)ju) * (double)g.W / (double)n);
    if (off >= g.W) off = g.W - 1;
    int64_t p = lo + off;
    const int32_t start = (int32_t)p;
    const uint64_t hf = h2(rid, 2);
    const int hap = (int)(hf & 1u);
    const bool reverse = ((hf >> 1) & 1u) != 0;
    const bool is_r1 = ((hf >> 2) & 1u) != 0;
    *flag_out = 0x1u | 0x2u | (reverse ? 0x10u : 0x20u) | (is_r1 ? 0x40u : 0x80u);
    if (P.depth_var_pct > 0) {                                           // depth varies from window to window: the reads a window loses stay in the
        const float keep = 1.0f - 0.01f * (float)P.depth_var_pct * unif(h2(P.seed ^ kDepth, (uint64_t)w * 2u + (uint64_t)ds));   // batch as placed-unmapped records
        if (unif(h2(rid, 7)) >= keep) *flag_out |= 0x4u;
    }
    int clip = 0; bool clip_head = false;
    if (unif(h2(rid, 3)) < P.clip_frac) {
        int mc = P.max_clip < L / 3 ? P.max_clip : L / 3;
        if (mc < 1) mc = 1;
        clip = 1 + (int)((hf >> 8) % (uint64_t)mc);
        clip_head = ((hf >> 3) & 1u) != 0;
    }
    const int target = L - ((clip && !clip_head) ? clip : 0);
    if (clip && clip_head) {
        for (int k = 0; k < clip; ++k) sink_base(s, (int)(h3(rid, 4, (uint64_t)k) & 3u));
        sink_op(s, 4u, clip);
    }
    bool any_aligned = false;
    const bool tumor = (ds == 0);
    const bool carries_somatic = tumor && (unif(h2(rid, 5)) < P.somatic_vaf);
    while (s.n_bases < target) {
        if (p != pw) {
            const Var v = germline(P, p);
            if (v.kind && ((v.hap >> hap) & 1u) && s.n_ops < 12) {
                if (v.kind == 1) {
                    sink_base(s, (ref_idx(P, p) + 1 + (int)((v.h >> 20) % 3u)) & 3);
                    sink_op(s, 0u, 1);
                    ++p; any_aligned = true;
                    continue;
                }
                if (v.kind == 2 && any_aligned) {
                    const int room = target - s.n_bases;
                    if (room <= v.len) {           // the read ends inside the insertion: an aligner soft-clips it
                        for (int k = 0; k < room; ++k) sink_base(s, (int)(h3(P.seed ^ kIns, (uint64_t)p, (uint64_t)k) & 3u));
                        sink_op(s, 4u, room);
                        break;
                    }
                    for (int k = 0; k < v.len; ++k) sink_base(s, (int)(h3(P.seed ^ kIns, (uint64_t)p, (uint64_t)k) & 3u));
                    sink_op(s, 1u, v.len);
                    // fall through: the reference base at p follows the insertion
                } else if (v.kind == 3 && any_aligned) {
                    sink_op(s, 2u, v.len);
                    p += v.len;
                    continue;
                }
            }
        }
        int b = ref_idx(P, p);
        if (p == pw && carries_somatic) b = som_alt_idx(P, pw);
        sink_base(s, b);
        sink_op(s, 0u, 1);
        ++p; any_aligned = true;
    }
    if (clip && !clip_head) {
        for (int k = 0; k < clip; ++k) sink_base(s, (int)(h3(rid, 4, (uint64_t)k) & 3u));
        sink_op(s, 4u, clip);
    }
    sink_flush_op(s);
    if (FILL) {
        if (s.seq_words) {
            if (s.n_bases & 7) s.seq_words[s.n_bases >> 3] = s.acc;
            for (int wd = (s.n_bases + 7) >> 3; wd < g.units * 4; ++wd) s.seq_words[wd] = 0u;
        }
        if (s.qual) {
            for (int k = 0; k < g.units * 32; ++k)
                s.qual[k] = k < L ? (uint8_t)(2u + (uint32_t)(h3(rid, 6, (uint64_t)k) % 39u)) : (uint8_t)0;
        }
    }
    return start;
}

struct ReadCoord { int ds; int64_t w; int i; };
GS_HD ReadCoord coord_of(const ga_synth_params& P, const Geo& g, int64_t r) {
    ReadCoord c;
    const int64_t n_t = (int64_t)P.n_windows * g.n_per[0];
    c.ds = r < n_t ? 0 : 1;
    const int64_t rr = c.ds ? r - n_t : r;
    const int per = g.n_per[c.ds] > 0 ? g.n_per[c.ds] : 1;
    c.w = (int64_t)P.window_begin + rr / per;
    c.i = (int)(rr % per);
    return c;
}

__host__ __device__ inline void count_one(const ga_synth_params& P, const Geo& g, int64_t r, uint32_t* n_ops, uint8_t* has_indel, int32_t* span_out) {
    const ReadCoord c = coord_of(P, g, r);
    Sink s; s.seq_words = nullptr; s.cigar = nullptr; s.qual = nullptr;
    uint32_t flag;
    build_read<false>(P, g, c.ds, c.w, c.i, s, &flag);
    n_ops[r] = (uint32_t)s.n_ops;
    has_indel[r] = (uint8_t)s.has_indel;
    *span_out = s.ref_span;
}

__host__ __device__ inline void fill_one(const ga_synth_params& P, const Geo& g, int64_t r, const ga_reads& D, const int32_t* qual_slot) {
    const ReadCoord c = coord_of(P, g, r);
    Sink s;
    s.seq_words = reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(D.seq4) + 16ull * (uint64_t)g.units * (uint64_t)r);
    s.cigar = const_cast<uint32_t*>(D.cigar) + D.cigar_off[r];
    const int32_t qs = qual_slot ? qual_slot[r] : (int32_t)r;
    s.qual = (D.qual && qs >= 0) ? const_cast<uint8_t*>(D.qual) + 32ull * (uint64_t)g.units * (uint64_t)qs : nullptr;
    uint32_t flag;
    const int32_t start = build_read<true>(P, g, c.ds, c.w, c.i, s, &flag);
    const_cast<int32_t*>(D.pos)[r] = start;
    const_cast<uint32_t*>(D.len_flag)[r] = (flag << 16) | (uint32_t)P.read_len;
    const_cast<uint32_t*>(D.seq_off16)[r] = (uint32_t)((uint64_t)g.units * (uint64_t)r);
}

__global__ void count_kernel(ga_synth_params P, Geo g, int64_t n, uint32_t* n_ops, uint8_t* has_indel, int32_t* max_span) {
    int m = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        int32_t sp; count_one(P, g, r, n_ops, has_indel, &sp);
        m = max(m, sp);
    }
    for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_span, m);
}
__global__ void fill_kernel(ga_synth_params P, Geo g, int64_t n, ga_reads D, const int32_t* qual_slot) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        fill_one(P, g, r, D, qual_slot);
}
__global__ void reference_kernel(ga_synth_params P, uint8_t* out, int64_t begin, int64_t n) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        out[k] = (uint8_t)"ACGT"[ref_idx(P, begin + k)];
}

__host__ __device__ inline void session_one(const ga_synth_params& P, const Geo& g, int k, int32_t* first, int32_t* last, int32_t* kt, int32_t* kp,
                                            int32_t* ke, int32_t* kl, uint32_t* koff, uint8_t* kall) {
    const int64_t pw = window_pos(P, g, (int64_t)P.window_begin + k);
    first[k] = (int32_t)(pw + 1 - P.window_half);             // VCF pos (1-based) - window_size/2
    last[k] = (int32_t)(pw + 1 + P.window_half + 1);
    kt[k] = GA_VT_SNV; kp[k] = (int32_t)pw; ke[k] = (int32_t)pw; kl[k] = 1;
    koff[k] = (uint32_t)k;
    if (k == P.n_windows - 1) koff[k + 1] = (uint32_t)(k + 1);
    kall[k] = (uint8_t)"ACGT"[som_alt_idx(P, pw)];
}
__global__ void sessions_kernel(ga_synth_params P, Geo g, int32_t* first, int32_t* last, int32_t* kt, int32_t* kp, int32_t* ke, int32_t* kl,
                                uint32_t* koff, uint8_t* kall) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < P.n_windows) session_one(P, g, k, first, last, kt, kp, ke, kl, koff, kall);
}

static bool valid(const ga_synth_params* p) {
    if (!p || p->read_len < 8 || p->read_len > 0xffff || p->n_windows < 0 || p->window_begin < 0 || p->total_windows < 1) return false;
    if (p->window_begin + (int64_t)p->n_windows > p->total_windows) return false;
    const Geo g = geometry(*p);
    const int32_t min_spacing = 2 * p->window_half + 2 * p->read_len + 64;
    return g.stride >= min_spacing + 1 && p->contig_len > 4ll * g.margin;
}

}  // namespace gs

extern "C" {

int ga_synth_plan_sizes(const ga_synth_params* p, ga_synth_plan* out) {
    if (!out || !gs::valid(p)) return GA_ERR_BAD_ARGUMENT;
    const gs::Geo g = gs::geometry(*p);
    out->reads_per_window[0] = g.n_per[0]; out->reads_per_window[1] = g.n_per[1];
    out->n_tumor = (int64_t)p->n_windows * g.n_per[0];
    out->n_reads = out->n_tumor + (int64_t)p->n_windows * g.n_per[1];
    out->units_per_read = g.units;
    out->seq4_bytes = out->n_reads * 16ll * g.units;
    out->window_stride = g.stride;
    return GA_OK;
}

#define GS_LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? GA_OK : GA_ERR_CUDA)

int ga_synth_reference(const ga_synth_params* p, uint8_t* d_ascii, int64_t begin, int64_t n, void* stream) {
    if (!gs::valid(p) || !d_ascii || begin < 0 || n < 0) return GA_ERR_BAD_ARGUMENT;
    if (n == 0) return GA_OK;
    gs::reference_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(*p, d_ascii, begin, n);
    return GS_LAUNCH_OK();
}

int ga_synth_sessions(const ga_synth_params* p, int32_t* first, int32_t* last, int32_t* kt, int32_t* kp, int32_t* ke, int32_t* kl,
                      uint32_t* koff, uint8_t* kall, void* stream) {
    if (!gs::valid(p)) return GA_ERR_BAD_ARGUMENT;
    if (p->n_windows == 0) return GA_OK;
    gs::sessions_kernel<<<(p->n_windows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*p, gs::geometry(*p), first, last, kt, kp, ke, kl, koff, kall);
    return GS_LAUNCH_OK();
}

int ga_synth_reads_count(const ga_synth_params* p, uint32_t* n_ops, uint8_t* has_indel, int32_t* max_ref_span, void* stream) {
    ga_synth_plan pl;
    if (ga_synth_plan_sizes(p, &pl) != GA_OK) return GA_ERR_BAD_ARGUMENT;
    if (pl.n_reads == 0) return GA_OK;
    gs::count_kernel<<<148 * 8, 128, 0, (cudaStream_t)stream>>>(*p, gs::geometry(*p), pl.n_reads, n_ops, has_indel, max_ref_span);
    return GS_LAUNCH_OK();
}

int ga_synth_reads_fill(const ga_synth_params* p, const ga_reads* dst, const int32_t* qual_slot, void* stream) {
    ga_synth_plan pl;
    if (!dst || ga_synth_plan_sizes(p, &pl) != GA_OK) return GA_ERR_BAD_ARGUMENT;
    if (pl.n_reads == 0) return GA_OK;
    gs::fill_kernel<<<148 * 8, 128, 0, (cudaStream_t)stream>>>(*p, gs::geometry(*p), pl.n_reads, *dst, qual_slot);
    return GS_LAUNCH_OK();
}

// Host twins (same arithmetic, host pointers) for CPU-side tests of the sharding / merge logic.
int ga_synth_reference_host(const ga_synth_params* p, uint8_t* ascii, int64_t begin, int64_t n) {
    if (!gs::valid(p) || !ascii) return GA_ERR_BAD_ARGUMENT;
    for (int64_t k = 0; k < n; ++k) ascii[k] = (uint8_t)"ACGT"[gs::ref_idx(*p, begin + k)];
    return GA_OK;
}
int ga_synth_sessions_host(const ga_synth_params* p, int32_t* first, int32_t* last, int32_t* kt, int32_t* kp, int32_t* ke, int32_t* kl,
                           uint32_t* koff, uint8_t* kall) {
    if (!gs::valid(p)) return GA_ERR_BAD_ARGUMENT;
    const gs::Geo g = gs::geometry(*p);
    for (int k = 0; k < p->n_windows; ++k) gs::session_one(*p, g, k, first, last, kt, kp, ke, kl, koff, kall);
    return GA_OK;
}
int ga_synth_reads_count_host(const ga_synth_params* p, uint32_t* n_ops, uint8_t* has_indel, int32_t* max_ref_span) {
    ga_synth_plan pl;
    if (ga_synth_plan_sizes(p, &pl) != GA_OK) return GA_ERR_BAD_ARGUMENT;
    const gs::Geo g = gs::geometry(*p);
    int32_t m = *max_ref_span;
    for (int64_t r = 0; r < pl.n_reads; ++r) { int32_t sp; gs::count_one(*p, g, r, n_ops, has_indel, &sp); if (sp > m) m = sp; }
    *max_ref_span = m;
    return GA_OK;
}
int ga_synth_reads_fill_host(const ga_synth_params* p, const ga_reads* dst, const int32_t* qual_slot) {
    ga_synth_plan pl;
    if (!dst || ga_synth_plan_sizes(p, &pl) != GA_OK) return GA_ERR_BAD_ARGUMENT;
    const gs::Geo g = gs::geometry(*p);
    for (int64_t r = 0; r < pl.n_reads; ++r) gs::fill_one(*p, g, r, *dst, qual_slot);
    return GA_OK;
}

}  // extern "C"
