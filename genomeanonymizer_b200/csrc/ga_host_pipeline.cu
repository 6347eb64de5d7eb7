// ga_host_pipeline.cu - ga_run_host(): host SoA batch -> chunked H2D -> session kernels -> D2H, double buffered.
//
// This is the call the reference-facing plugin makes (bench.py's `e2e` line): the caller hands over HOST
// buffers (pinned for full PCIe speed) holding the reads of one contig and the session table; the engine
// cuts the table into chunks of consecutive sessions, uploads only the read slices each chunk needs,
// runs the chunk on one of three lanes (streams + scratch) so the uploads of the next chunks overlap chunk k's kernels
// and download, and appends the compacted modified records to the caller's host result.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

#include "ga_engine_internal.h"
#include "../../include/ga_wire.h"

namespace ga {

// Device buffer that only grows.
struct DevBuf {
    uint8_t* p = nullptr; size_t cap = 0;
    cudaError_t need(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = n + n / 4 + 4096;
        cudaError_t ce = cudaMalloc(&p, want);
        if (ce == cudaSuccess) cap = want;
        return ce;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// ga_run_host's precondition (include/ga_b200.h): inside each dataset the seq4 records - and the sparse quality records -
// lie in read order without overlap, so a chunk's records are one contiguous slice.  Checked on the uploaded slice.
__global__ void check_layout_kernel(const uint32_t* __restrict__ seq_off16, const uint32_t* __restrict__ len_flag, int64_t n_t, int64_t n_all,
                                    const int32_t* __restrict__ qual_reads, const uint32_t* __restrict__ qual_off16, int64_t nq_t, int64_t nq_all,
                                    int32_t qr_sub_t, int32_t qr_sub_n, int32_t qr_add_n, uint32_t* __restrict__ bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_all; r += stride) {
        if (r > 0 && r != n_t) {
            const uint32_t L = len_flag[r - 1] & 0xffffu, units = L ? (L + 31u) / 32u : 1u;
            if (seq_off16[r] < seq_off16[r - 1] + units) atomicOr(bad, 1u);
        }
        if (r < nq_all && r > 0 && r != nq_t) {
            const int32_t prev = r - 1 < nq_t ? qual_reads[r - 1] - qr_sub_t : qual_reads[r - 1] - qr_sub_n + qr_add_n;   // chunk-local read index
            const uint32_t L = (prev >= 0 && prev < n_all) ? (len_flag[prev] & 0xffffu) : 0u, units = L ? (L + 31u) / 32u : 1u;
            if (qual_off16[r] < qual_off16[r - 1] + units) atomicOr(bad, 2u);
        }
    }
}

// chunk-local offsets -> consistent offsets of the concatenated (tumor slice | normal slice) mini batch
__global__ void rebase_kernel(uint32_t* seq_off16, uint32_t* cigar_off, int32_t* qual_reads, uint32_t* qual_off16,
                              int64_t n_t, int64_t n_all, uint32_t seq_sub_t, uint32_t seq_sub_n, uint32_t seq_add_n,
                              uint32_t cig_sub_t, uint32_t cig_sub_n, uint32_t cig_add_n,
                              int64_t nq_t, int64_t nq_all, int32_t qr_sub_t, int32_t qr_sub_n, int32_t qr_add_n,
                              uint32_t qo_sub_t, uint32_t qo_sub_n, uint32_t qo_add_n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= n_all; r += stride) {
        if (seq_off16) {                                      // null: the wire expansion already wrote chunk-local offsets
            if (r < n_all) seq_off16[r] = r < n_t ? seq_off16[r] - seq_sub_t : seq_off16[r] - seq_sub_n + seq_add_n;
            cigar_off[r] = r < n_t ? cigar_off[r] - cig_sub_t : cigar_off[r] - cig_sub_n + cig_add_n;
        }
        if (r < nq_all) {
            qual_reads[r] = r < nq_t ? qual_reads[r] - qr_sub_t : qual_reads[r] - qr_sub_n + qr_add_n;
            qual_off16[r] = r < nq_t ? qual_off16[r] - qo_sub_t : qual_off16[r] - qo_sub_n + qo_add_n;
        }
    }
}

// chunk-local record indices -> table / batch indices and global output offsets
__global__ void finalize_kernel(int32_t* mod_session, int32_t* mod_read, uint32_t* mod_seq_off16, uint32_t* mod_qual_off16, int64_t n,
                                int32_t sess_base, int32_t n_t, int32_t t_lo, int32_t n_lo, uint32_t seq_base, uint32_t qual_base) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    mod_session[k] += sess_base;
    const int32_t i = mod_read[k];
    mod_read[k] = i < n_t ? t_lo + i : n_lo + (i - n_t);
    mod_seq_off16[k] += seq_base;
    if (mod_qual_off16[k] != 0xffffffffu) mod_qual_off16[k] += qual_base;
}

}  // namespace ga

struct HostSlot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    ga::DevBuf pos, len_flag, seq_off16, cigar_off, cigar, seq4, qual, qual_reads, qual_off16;
    ga::DevBuf s_first, s_last, s_kt, s_kp, s_ke, s_kl, s_koff, s_kall;
    ga::DevBuf w_blob, w_dir;            // ga_run_wire: the chunk's blocks as they came over PCIe, and their directory entries
    ga::DevBuf o_sess, o_read, o_len, o_soff, o_qoff, o_seq, o_qual, o_counts, o_totals;
    ga_totals* h_totals = nullptr;       // pinned
    uint32_t* d_bad = nullptr; uint32_t* h_bad = nullptr;   // layout check of the uploaded slice (device flag, pinned copy)
    uint32_t* h_koff = nullptr; size_t cap_koff = 0;   // pinned staging of the rebased keep_allele_off slice
    // description of the chunk in flight
    bool busy = false;
    int32_t s0 = 0, s1 = 0; int64_t t_lo = 0, t_hi = 0, n_lo = 0, n_hi = 0;
    uint32_t soa_seq_sub_t = 0, soa_seq_sub_n = 0, soa_cig_sub_t = 0, soa_cig_sub_n = 0;   // first seq4 unit / CIGAR word of the uploaded slices (SoA source)
    int64_t cap_rec = 0, cap_seq = 0, cap_qual = 0;
    ga_reads R; ga_sessions S; ga_result O;
};

void ga_host_slots_destroy(ga_engine* e) {
    if (!e->slots) return;
    for (int k = 0; k < kLanes; ++k) {
        HostSlot& h = e->slots[k];
        ga::DevBuf* all[] = {&h.w_blob, &h.w_dir, &h.pos, &h.len_flag, &h.seq_off16, &h.cigar_off, &h.cigar, &h.seq4, &h.qual, &h.qual_reads, &h.qual_off16,
                             &h.s_first, &h.s_last, &h.s_kt, &h.s_kp, &h.s_ke, &h.s_kl, &h.s_koff, &h.s_kall,
                             &h.o_sess, &h.o_read, &h.o_len, &h.o_soff, &h.o_qoff, &h.o_seq, &h.o_qual, &h.o_counts, &h.o_totals};
        for (ga::DevBuf* b : all) b->release();
        if (h.h_totals) cudaFreeHost(h.h_totals);
        if (h.h_bad) cudaFreeHost(h.h_bad);
        if (h.d_bad) cudaFree(h.d_bad);
        if (h.h_koff) cudaFreeHost(h.h_koff);
        if (h.st) cudaStreamDestroy(h.st);
        if (h.done) cudaEventDestroy(h.done);
    }
    delete[] e->slots;
    e->slots = nullptr;
}

static int ensure_slots(ga_engine* e) {
    if (e->slots) return GA_OK;
    e->slots = new HostSlot[kLanes];
    for (int k = 0; k < kLanes; ++k) {
        GA_CUDA(cudaStreamCreateWithFlags(&e->slots[k].st, cudaStreamNonBlocking));
        GA_CUDA(cudaEventCreateWithFlags(&e->slots[k].done, cudaEventDisableTiming));
        GA_CUDA(cudaHostAlloc(&e->slots[k].h_totals, sizeof(ga_totals), cudaHostAllocDefault));
        GA_CUDA(cudaHostAlloc(&e->slots[k].h_bad, sizeof(uint32_t), cudaHostAllocDefault));
        GA_CUDA(cudaMalloc(&e->slots[k].d_bad, sizeof(uint32_t)));
    }
    return GA_OK;
}

static int64_t lower_bound_i32(const int32_t* a, int64_t b, int64_t e, int64_t v) {
    while (b < e) { const int64_t m = (b + e) >> 1; if ((int64_t)a[m] < v) b = m + 1; else e = m; }
    return b;
}

static inline uint32_t units_of(uint32_t len_flag) { const uint32_t L = len_flag & 0xffffu; return L ? (L + 31u) / 32u : 1u; }

#define H2D(dst, src, bytes) do { if ((bytes) > 0) { GA_CUDA(cudaMemcpyAsync((dst), (src), (size_t)(bytes), cudaMemcpyHostToDevice, h.st)); e->last_h2d += (int64_t)(bytes); } } while (0)
#define D2H(dst, src, bytes) do { if ((bytes) > 0) { GA_CUDA(cudaMemcpyAsync((dst), (src), (size_t)(bytes), cudaMemcpyDeviceToHost, h.st)); e->last_d2h += (int64_t)(bytes); } } while (0)
#define NEED(buf, bytes) GA_CUDA((buf).need((size_t)(bytes) + 256))

// What ga_run_host / ga_run_wire were handed: exactly one of R (structure of arrays) and W (wire form) is set.
struct HostSource {
    const ga_reads* R = nullptr;
    const ga_reads_wire* W = nullptr;
    int64_t n_reads() const { return R ? R->n_reads : W->n_reads; }
    int64_t n_tumor() const { return R ? R->n_tumor : W->n_tumor; }
    int32_t contig_id() const { return R ? R->contig_id : W->contig_id; }
    const uint8_t* qual() const { return R ? R->qual : W->qual; }
    const int32_t* qual_reads() const { return R ? R->qual_reads : W->qual_reads; }
    const uint32_t* qual_off16() const { return R ? R->qual_off16 : W->qual_off16; }
    int64_t n_qual() const { return R ? R->n_qual : W->n_qual; }
};

// First block of dataset `ds` whose first read starts at or after v (wire form).
static int64_t wire_lower_block(const ga_reads_wire* W, int ds, int64_t v) {
    int64_t b = ds ? W->n_tumor_blocks : 0, e = ds ? W->n_blocks : W->n_tumor_blocks;
    while (b < e) { const int64_t m = (b + e) >> 1; if ((int64_t)W->dir[m].pos < v) b = m + 1; else e = m; }
    return b;
}

// Read ranges [t_lo, t_hi) / [n_lo, n_hi) a chunk with reads starting in [lo_v, max_last) needs.  The wire form is cut
// at block boundaries (the extra reads lie outside every session of the chunk; the assignment kernel skips them).
static void chunk_read_ranges(const HostSource& src, int64_t lo_v, int64_t max_last, int64_t* t_lo, int64_t* t_hi, int64_t* n_lo, int64_t* n_hi,
                              int64_t* blocks /* [4] tumor begin, end, normal begin, end; wire only */) {
    if (src.R) {
        const ga_reads* R = src.R;
        *t_lo = lower_bound_i32(R->pos, 0, R->n_tumor, lo_v);       *t_hi = lower_bound_i32(R->pos, 0, R->n_tumor, max_last);
        *n_lo = lower_bound_i32(R->pos, R->n_tumor, R->n_reads, lo_v); *n_hi = lower_bound_i32(R->pos, R->n_tumor, R->n_reads, max_last);
    } else {
        const ga_reads_wire* W = src.W;
        for (int ds = 0; ds < 2; ++ds) {
            const int64_t b_first = ds ? W->n_tumor_blocks : 0, b_last = ds ? W->n_blocks : W->n_tumor_blocks;
            int64_t b0 = wire_lower_block(W, ds, lo_v), b1 = wire_lower_block(W, ds, max_last);
            if (b0 > b_first) --b0;                                  // the block before may still hold reads >= lo_v
            if (b1 < b0) b1 = b0;
            (void)b_last;
            blocks[2 * ds] = b0; blocks[2 * ds + 1] = b1;
            *(ds ? n_lo : t_lo) = W->dir[b0].read; *(ds ? n_hi : t_hi) = W->dir[b1].read;
        }
    }
    if (*t_hi < *t_lo) *t_hi = *t_lo;
    if (*n_hi < *n_lo) *n_hi = *n_lo;
}

static int launch_chunk(ga_engine* e, int lane, const HostSource& src, const ga_sessions* S, int32_t s0, int32_t s1, int32_t maxspan,
                        int64_t cap_rec, int64_t cap_seq, int64_t cap_qual, int64_t units_t, int64_t units_all, int64_t ops_t, bool check_layout);
int ga_wire_expand(ga_engine* e, cudaStream_t st, const uint8_t* d_blob, const ga_wire_dir* d_dir, int n_blocks_t, int n_blocks_all,
                   uint64_t byte0_t, uint64_t byte0_n, uint64_t place_n, uint32_t read0_t, uint32_t read0_n, uint32_t n_t,
                   uint32_t unit0_t, uint32_t unit0_n, uint32_t units_t, uint32_t ops0_t, uint32_t ops0_n, uint32_t ops_t, uint32_t n_all, uint32_t ops_all,
                   int32_t* pos, uint32_t* len_flag, uint32_t* seq_off16, uint32_t* cigar_off, uint32_t* cigar, uint8_t* seq4);

// Uploads the slices of chunk [s0,s1) and launches it on lane `lane`.
static int submit_chunk(ga_engine* e, int lane, const HostSource& src, const ga_sessions* S, int32_t s0, int32_t s1, int32_t maxspan,
                        int64_t cap_rec, int64_t cap_seq, int64_t cap_qual) {
    HostSlot& h = e->slots[lane];
    // read ranges: sessions are sorted by first; last is not necessarily monotone
    int64_t max_last = S->last[s0];
    for (int32_t s = s0 + 1; s < s1; ++s) max_last = std::max<int64_t>(max_last, S->last[s]);
    const int64_t lo_v = (int64_t)S->first[s0] - maxspan + 1;
    int64_t blocks[4] = {0, 0, 0, 0};
    chunk_read_ranges(src, lo_v, max_last, &h.t_lo, &h.t_hi, &h.n_lo, &h.n_hi, blocks);
    h.s0 = s0; h.s1 = s1;
    const int64_t n_t = h.t_hi - h.t_lo, n_n = h.n_hi - h.n_lo, n_all = n_t + n_n;
    NEED(h.pos, 4 * n_all); NEED(h.len_flag, 4 * n_all); NEED(h.seq_off16, 4 * n_all); NEED(h.cigar_off, 4 * (n_all + 1));
    if (src.W) {
        // ---- wire form: one slice of the blob and one of the directory per dataset, expanded on the device
        const ga_reads_wire* W = src.W;
        const ga_wire_dir &dt0 = W->dir[blocks[0]], &dt1 = W->dir[blocks[1]], &dn0 = W->dir[blocks[2]], &dn1 = W->dir[blocks[3]];
        const int64_t bytes_t = (int64_t)(dt1.byte - dt0.byte), bytes_n = (int64_t)(dn1.byte - dn0.byte);
        const int64_t nb_t = blocks[1] - blocks[0], nb_n = blocks[3] - blocks[2];
        const int64_t units_t = (int64_t)dt1.unit - dt0.unit, units_n = (int64_t)dn1.unit - dn0.unit;
        const int64_t ops_t = (int64_t)dt1.ops - dt0.ops, ops_n = (int64_t)dn1.ops - dn0.ops;
        NEED(h.w_blob, bytes_t + bytes_n); NEED(h.w_dir, sizeof(ga_wire_dir) * (nb_t + nb_n + 1));
        NEED(h.cigar, 4ll * (ops_t + ops_n)); NEED(h.seq4, 16ll * (units_t + units_n));
        H2D(h.w_blob.p, W->blob + dt0.byte, bytes_t);
        H2D(h.w_blob.p + bytes_t, W->blob + dn0.byte, bytes_n);
        H2D(h.w_dir.p, W->dir + blocks[0], sizeof(ga_wire_dir) * nb_t);
        H2D(h.w_dir.p + sizeof(ga_wire_dir) * nb_t, W->dir + blocks[2], sizeof(ga_wire_dir) * nb_n);
        if (n_all > 0) {
            int rc = ga_wire_expand(e, h.st, h.w_blob.p, h.w_dir.as<ga_wire_dir>(), (int)nb_t, (int)(nb_t + nb_n), dt0.byte, dn0.byte, (uint64_t)bytes_t,
                                    dt0.read, dn0.read, (uint32_t)n_t, dt0.unit, dn0.unit, (uint32_t)units_t, dt0.ops, dn0.ops, (uint32_t)ops_t,
                                    (uint32_t)n_all, (uint32_t)(ops_t + ops_n),
                                    h.pos.as<int32_t>(), h.len_flag.as<uint32_t>(), h.seq_off16.as<uint32_t>(), h.cigar_off.as<uint32_t>(), h.cigar.as<uint32_t>(), h.seq4.p);
            if (rc) return rc;
        }
        return launch_chunk(e, lane, src, S, s0, s1, maxspan, cap_rec, cap_seq, cap_qual, units_t, units_t + units_n, ops_t, false);
    }
    const ga_reads* R = src.R;

    auto seq_end = [&](int64_t r_end, int64_t r_begin) -> uint32_t {     // first unit after read r_end-1
        if (r_end <= r_begin) return r_begin < R->n_reads ? R->seq_off16[r_begin] : 0u;
        return R->seq_off16[r_end - 1] + units_of(R->len_flag[r_end - 1]);
    };
    const uint32_t su_t0 = n_t ? R->seq_off16[h.t_lo] : 0u, su_t1 = n_t ? seq_end(h.t_hi, h.t_lo) : 0u;
    const uint32_t su_n0 = n_n ? R->seq_off16[h.n_lo] : 0u, su_n1 = n_n ? seq_end(h.n_hi, h.n_lo) : 0u;
    if (su_t1 < su_t0 || su_n1 < su_n0 || 16ll * su_t1 > R->seq4_bytes || 16ll * su_n1 > R->seq4_bytes)
        return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_host: seq_off16 must ascend with the read index inside each dataset (records of a chunk are one slice)");
    const uint32_t units_t = su_t1 - su_t0, units_n = su_n1 - su_n0;
    const uint32_t cg_t0 = R->cigar_off[h.t_lo], cg_t1 = R->cigar_off[h.t_hi];
    const uint32_t cg_n0 = R->cigar_off[h.n_lo], cg_n1 = R->cigar_off[h.n_hi];
    const uint32_t ops_t = cg_t1 - cg_t0, ops_n = cg_n1 - cg_n0;

    NEED(h.cigar, 4ll * (ops_t + ops_n)); NEED(h.seq4, 16ll * (units_t + units_n));
    H2D(h.pos.as<int32_t>(), R->pos + h.t_lo, 4 * n_t);                 H2D(h.pos.as<int32_t>() + n_t, R->pos + h.n_lo, 4 * n_n);
    H2D(h.len_flag.as<uint32_t>(), R->len_flag + h.t_lo, 4 * n_t);      H2D(h.len_flag.as<uint32_t>() + n_t, R->len_flag + h.n_lo, 4 * n_n);
    H2D(h.seq_off16.as<uint32_t>(), R->seq_off16 + h.t_lo, 4 * n_t);    H2D(h.seq_off16.as<uint32_t>() + n_t, R->seq_off16 + h.n_lo, 4 * n_n);
    H2D(h.cigar_off.as<uint32_t>(), R->cigar_off + h.t_lo, 4 * n_t);    H2D(h.cigar_off.as<uint32_t>() + n_t, R->cigar_off + h.n_lo, 4 * (n_n + 1));
    H2D(h.cigar.as<uint32_t>(), R->cigar + cg_t0, 4ll * ops_t);         H2D(h.cigar.as<uint32_t>() + ops_t, R->cigar + cg_n0, 4ll * ops_n);
    H2D(h.seq4.p, R->seq4 + 16ull * su_t0, 16ll * units_t);             H2D(h.seq4.p + 16ull * units_t, R->seq4 + 16ull * su_n0, 16ll * units_n);
    h.soa_seq_sub_t = su_t0; h.soa_seq_sub_n = su_n0; h.soa_cig_sub_t = cg_t0; h.soa_cig_sub_n = cg_n0;
    return launch_chunk(e, lane, src, S, s0, s1, maxspan, cap_rec, cap_seq, cap_qual, units_t, (int64_t)units_t + units_n, ops_t, true);
}

// Second half of a chunk submission: qualities, session table slice, output buffers, the masking pass.
static int launch_chunk(ga_engine* e, int lane, const HostSource& src, const ga_sessions* S, int32_t s0, int32_t s1, int32_t maxspan,
                        int64_t cap_rec, int64_t cap_seq, int64_t cap_qual, int64_t units_t, int64_t units_all, int64_t ops_t, bool soa) {
    HostSlot& h = e->slots[lane];
    const int32_t ns = s1 - s0;
    const int64_t n_t = h.t_hi - h.t_lo, n_n = h.n_hi - h.n_lo, n_all = n_t + n_n;
    const int64_t units_n = units_all - units_t;
    const uint8_t* Q = src.qual(); const int32_t* QR = src.qual_reads(); const uint32_t* QO = src.qual_off16(); const int64_t NQ = src.n_qual();

    // qualities: dense (record at 32*seq_off16) or sparse (qual_reads / qual_off16)
    int64_t nq_t = 0, nq_n = 0;
    uint32_t qo_t0 = 0, qo_n0 = 0, qu_t = 0;
    int64_t qa_t = 0, qa_n = 0;
    const bool sparse = Q && QR;
    if (Q && !sparse) {
        const ga_reads* R = src.R;                             // the wire form is always sparse
        NEED(h.qual, 32ll * units_all);
        H2D(h.qual.p, R->qual + 32ull * h.soa_seq_sub_t, 32ll * units_t);
        H2D(h.qual.p + 32ull * units_t, R->qual + 32ull * h.soa_seq_sub_n, 32ll * units_n);
    } else if (sparse) {
        qa_t = lower_bound_i32(QR, 0, NQ, h.t_lo);
        const int64_t qb_t = lower_bound_i32(QR, 0, NQ, h.t_hi);
        qa_n = lower_bound_i32(QR, 0, NQ, h.n_lo);
        const int64_t qb_n = lower_bound_i32(QR, 0, NQ, h.n_hi);
        nq_t = qb_t - qa_t; nq_n = qb_n - qa_n;
        auto q_end = [&](int64_t qb) -> uint32_t {            // first unit behind quality record qb - 1
            if (src.R) return QO[qb - 1] + units_of(src.R->len_flag[QR[qb - 1]]);
            return qb < NQ ? QO[qb] : (uint32_t)src.W->qual_units;   // wire: records are contiguous
        };
        qo_t0 = nq_t ? QO[qa_t] : 0u; qu_t = nq_t ? q_end(qb_t) - qo_t0 : 0u;
        qo_n0 = nq_n ? QO[qa_n] : 0u; const uint32_t qu_n = nq_n ? q_end(qb_n) - qo_n0 : 0u;
        if ((nq_t && q_end(qb_t) < qo_t0) || (nq_n && q_end(qb_n) < qo_n0))
            return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_host: qual_off16 must ascend with qual_reads inside each dataset");
        NEED(h.qual, 32ll * (qu_t + qu_n)); NEED(h.qual_reads, 4 * (nq_t + nq_n)); NEED(h.qual_off16, 4 * (nq_t + nq_n));
        H2D(h.qual_reads.as<int32_t>(), QR + qa_t, 4 * nq_t);     H2D(h.qual_reads.as<int32_t>() + nq_t, QR + qa_n, 4 * nq_n);
        H2D(h.qual_off16.as<uint32_t>(), QO + qa_t, 4 * nq_t);    H2D(h.qual_off16.as<uint32_t>() + nq_t, QO + qa_n, 4 * nq_n);
        H2D(h.qual.p, Q + 32ull * qo_t0, 32ll * qu_t);            H2D(h.qual.p + 32ull * qu_t, Q + 32ull * qo_n0, 32ll * qu_n);
    }
    GA_CUDA(cudaMemsetAsync(h.d_bad, 0, sizeof(uint32_t), h.st));
    if (soa) {
        ga::check_layout_kernel<<<e->n_sm * 2, 256, 0, h.st>>>(h.seq_off16.as<uint32_t>(), h.len_flag.as<uint32_t>(), n_t, n_all,
            sparse ? h.qual_reads.as<int32_t>() : nullptr, sparse ? h.qual_off16.as<uint32_t>() : nullptr, nq_t, sparse ? nq_t + nq_n : 0,
            (int32_t)h.t_lo, (int32_t)h.n_lo, (int32_t)n_t, h.d_bad);
        e->launches++;
    }
    if (soa || sparse) {
        ga::rebase_kernel<<<e->n_sm * 2, 256, 0, h.st>>>(soa ? h.seq_off16.as<uint32_t>() : nullptr, soa ? h.cigar_off.as<uint32_t>() : nullptr,
            sparse ? h.qual_reads.as<int32_t>() : nullptr, sparse ? h.qual_off16.as<uint32_t>() : nullptr,
            n_t, n_all, h.soa_seq_sub_t, h.soa_seq_sub_n, (uint32_t)units_t, h.soa_cig_sub_t, h.soa_cig_sub_n, (uint32_t)ops_t,
            nq_t, sparse ? nq_t + nq_n : 0, (int32_t)h.t_lo, (int32_t)h.n_lo, (int32_t)n_t, qo_t0, qo_n0, qu_t);
        e->launches++;
    }

    // session table slice
    NEED(h.s_first, 4 * ns); NEED(h.s_last, 4 * ns); NEED(h.s_kt, 4 * ns); NEED(h.s_kp, 4 * ns); NEED(h.s_ke, 4 * ns); NEED(h.s_kl, 4 * ns);
    NEED(h.s_koff, 4 * (ns + 1));
    const uint32_t ka0 = S->keep_allele_off[s0], ka1 = S->keep_allele_off[s1];
    NEED(h.s_kall, (ka1 - ka0) + 16);
    H2D(h.s_first.p, S->first + s0, 4 * ns); H2D(h.s_last.p, S->last + s0, 4 * ns);
    H2D(h.s_kt.p, S->keep_type + s0, 4 * ns); H2D(h.s_kp.p, S->keep_pos + s0, 4 * ns);
    H2D(h.s_ke.p, S->keep_end + s0, 4 * ns); H2D(h.s_kl.p, S->keep_len + s0, 4 * ns);
    H2D(h.s_kall.p, S->keep_alleles + ka0, ka1 - ka0);
    // keep_allele_off rebased to the slice, staged in the slot's own pinned buffer (free again once the
    // previous chunk of this lane was collected, which happens before the lane is reused)
    if ((size_t)ns + 1 > h.cap_koff) {
        if (h.h_koff) cudaFreeHost(h.h_koff);
        h.h_koff = nullptr; h.cap_koff = 0;
        GA_CUDA(cudaHostAlloc(&h.h_koff, 4 * ((size_t)ns + 1 + 1024), cudaHostAllocDefault));
        h.cap_koff = (size_t)ns + 1 + 1024;
    }
    for (int32_t k = 0; k <= ns; ++k) h.h_koff[k] = S->keep_allele_off[s0 + k] - ka0;
    H2D(h.s_koff.p, h.h_koff, 4 * (int64_t)(ns + 1));

    // outputs
    h.cap_rec = cap_rec; h.cap_seq = cap_seq; h.cap_qual = cap_qual;
    NEED(h.o_sess, 4 * cap_rec); NEED(h.o_read, 4 * cap_rec); NEED(h.o_len, 4 * cap_rec); NEED(h.o_soff, 4 * cap_rec); NEED(h.o_qoff, 4 * cap_rec);
    NEED(h.o_seq, 16 * cap_seq); NEED(h.o_qual, 32 * cap_qual); NEED(h.o_counts, 16ll * ns); NEED(h.o_totals, sizeof(ga_totals));

    memset(&h.R, 0, sizeof h.R);
    h.R.n_reads = n_all; h.R.n_tumor = n_t;
    h.R.pos = h.pos.as<int32_t>(); h.R.len_flag = h.len_flag.as<uint32_t>(); h.R.seq_off16 = h.seq_off16.as<uint32_t>();
    h.R.cigar_off = h.cigar_off.as<uint32_t>(); h.R.cigar = h.cigar.as<uint32_t>(); h.R.seq4 = h.seq4.p;
    h.R.qual = Q ? h.qual.p : nullptr;
    h.R.seq4_bytes = 16ll * units_all;
    h.R.n_qual = sparse ? nq_t + nq_n : 0;
    h.R.qual_reads = sparse ? h.qual_reads.as<int32_t>() : nullptr;
    h.R.qual_off16 = sparse ? h.qual_off16.as<uint32_t>() : nullptr;
    h.R.max_ref_span = maxspan; h.R.contig_id = src.contig_id();
    h.S.n_sessions = ns;
    h.S.first = h.s_first.as<int32_t>(); h.S.last = h.s_last.as<int32_t>(); h.S.keep_type = h.s_kt.as<int32_t>();
    h.S.keep_pos = h.s_kp.as<int32_t>(); h.S.keep_end = h.s_ke.as<int32_t>(); h.S.keep_len = h.s_kl.as<int32_t>();
    h.S.keep_allele_off = h.s_koff.as<uint32_t>(); h.S.keep_alleles = h.s_kall.p;
    h.O.cap_records = cap_rec; h.O.cap_seq16 = cap_seq; h.O.cap_qual16 = cap_qual;
    h.O.mod_session = h.o_sess.as<int32_t>(); h.O.mod_read = h.o_read.as<int32_t>(); h.O.mod_len = h.o_len.as<uint32_t>();
    h.O.mod_seq_off16 = h.o_soff.as<uint32_t>(); h.O.mod_qual_off16 = h.o_qoff.as<uint32_t>();
    h.O.out_seq4 = h.o_seq.p; h.O.out_qual = h.o_qual.p; h.O.sess_counts = h.o_counts.as<uint32_t>();
    h.O.totals = h.o_totals.as<ga_totals>();
    int rc = ga_run_lane(e, lane, &h.R, &h.S, &h.O, h.st);
    if (rc) return rc;
    D2H(h.h_totals, h.o_totals.p, sizeof(ga_totals));
    D2H(h.h_bad, h.d_bad, sizeof(uint32_t));
    GA_CUDA(cudaEventRecord(h.done, h.st));
    h.busy = true;
    return GA_OK;
}

// Waits for the chunk on `lane`, appends its records to the caller's result.  *retry is set when the chunk
// overflowed its device output buffers (the totals say how much it needs).
static int collect_chunk(ga_engine* e, int lane, ga_result* out, ga_totals* acc, bool* retry) {
    HostSlot& h = e->slots[lane];
    *retry = false;
    if (!h.busy) return GA_OK;
    GA_CUDA(cudaEventSynchronize(h.done));
    h.busy = false;
    const ga_totals t = *h.h_totals;
    if (*h.h_bad) {
        if (!acc->error) { acc->error = GA_ERR_BAD_ARGUMENT; acc->error_detail = (uint32_t)h.s0; }
        return ga_fail(e, GA_ERR_BAD_ARGUMENT, (*h.h_bad & 1u) ? "ga_run_host: seq4 records are not in read order inside a dataset (seq_off16 must ascend without overlap)"
                                                             : "ga_run_host: sparse quality records are not in read order (qual_off16 must ascend without overlap)");
    }
    if (t.error == GA_ERR_CAPACITY && t.error_detail == 0xffffffffu) { *retry = true; return GA_OK; }
    if (t.error) {
        if (!acc->error) { acc->error = t.error; acc->error_detail = t.error_detail; }
        return (int)t.error;
    }
    const int32_t ns = h.s1 - h.s0;
    const int64_t n = (int64_t)t.n_modified;
    const bool fits = (int64_t)(acc->n_modified + t.n_modified) <= out->cap_records &&
                      (int64_t)(acc->seq16_used + t.seq16_used) <= out->cap_seq16 &&
                      (int64_t)(acc->qual16_used + t.qual16_used) <= out->cap_qual16;
    if (fits && n > 0) {
        ga::finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h.st>>>(h.O.mod_session, h.O.mod_read, h.O.mod_seq_off16, h.O.mod_qual_off16, n,
            h.s0, (int32_t)(h.t_hi - h.t_lo), (int32_t)h.t_lo, (int32_t)h.n_lo, (uint32_t)acc->seq16_used, (uint32_t)acc->qual16_used);
        e->launches++;
        D2H(out->mod_session + acc->n_modified, h.O.mod_session, 4 * n);
        D2H(out->mod_read + acc->n_modified, h.O.mod_read, 4 * n);
        D2H(out->mod_len + acc->n_modified, h.O.mod_len, 4 * n);
        D2H(out->mod_seq_off16 + acc->n_modified, h.O.mod_seq_off16, 4 * n);
        D2H(out->mod_qual_off16 + acc->n_modified, h.O.mod_qual_off16, 4 * n);
        D2H(out->out_seq4 + 16ull * acc->seq16_used, h.O.out_seq4, 16ll * (int64_t)t.seq16_used);
        D2H(out->out_qual + 32ull * acc->qual16_used, h.O.out_qual, 32ll * (int64_t)t.qual16_used);
    }
    D2H(out->sess_counts + 4ull * h.s0, h.O.sess_counts, 16ll * ns);
    GA_CUDA(cudaStreamSynchronize(h.st));
    acc->n_modified += t.n_modified; acc->seq16_used += t.seq16_used; acc->qual16_used += t.qual16_used;
    acc->session_reads += t.session_reads; acc->session_bases += t.session_bases; acc->indel_records += t.indel_records;
    for (int k = 0; k < 3; ++k) acc->masked[k] += t.masked[k];
    if (!fits && !acc->error) { acc->error = GA_ERR_CAPACITY; acc->error_detail = 0xffffffffu; }
    return GA_OK;
}

extern "C" {

void ga_last_host_traffic(const ga_engine* e, int64_t* h2d, int64_t* d2h) {
    if (h2d) *h2d = e ? e->last_h2d : 0;
    if (d2h) *d2h = e ? e->last_d2h : 0;
}

static int run_host_source(ga_engine* e, const HostSource& src, const ga_sessions* S, ga_result* out, int64_t chunk_sessions, int32_t maxspan);

int ga_run_host(ga_engine* e, const ga_reads* R, const ga_sessions* S, ga_result* out, int64_t chunk_sessions) {
    if (!e || !R || !S || !out || !out->totals) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_host: null argument");
    if (R->n_reads < 0 || R->n_tumor < 0 || R->n_tumor > R->n_reads || S->n_sessions < 0)
        return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_host: negative or inconsistent sizes");
    int32_t maxspan = R->max_ref_span;
    if (maxspan <= 0) {                                    // host scan; callers that care about speed pass the bound
        maxspan = 1;
        for (int64_t r = 0; r < R->n_reads; ++r) {
            int32_t sp = 0;
            for (uint32_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; ++c) {
                const uint32_t op = R->cigar[c] & 15u;
                if (op == 0u || op == 2u || op == 3u || op == 7u || op == 8u) sp += (int32_t)(R->cigar[c] >> 4);
            }
            maxspan = std::max(maxspan, sp);
        }
    }
    HostSource src; src.R = R;
    return run_host_source(e, src, S, out, chunk_sessions, maxspan);
}

int ga_run_wire(ga_engine* e, const ga_reads_wire* W, const ga_sessions* S, ga_result* out, int64_t chunk_sessions) {
    if (!e || !W || !S || !out || !out->totals) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_wire: null argument");
    if (W->n_reads < 0 || W->n_tumor < 0 || W->n_tumor > W->n_reads || S->n_sessions < 0 || W->n_blocks < 0 || W->n_tumor_blocks < 0 ||
        W->n_tumor_blocks > W->n_blocks || (W->n_reads > 0 && (!W->blob || !W->dir)) || (W->n_blocks == 0) != (W->n_reads == 0))
        return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_wire: negative or inconsistent sizes");
    if ((reinterpret_cast<uintptr_t>(W->blob) & 15u) != 0u) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_wire: the blob must be 16-byte aligned");
    if (W->n_reads > 0 && (W->dir[W->n_tumor_blocks].read != (uint32_t)W->n_tumor || W->dir[W->n_blocks].read != (uint32_t)W->n_reads ||
                           W->dir[W->n_blocks].byte != (uint64_t)W->blob_bytes))
        return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_wire: the directory does not describe this batch");
    if (W->max_ref_span <= 0 && W->n_reads > 0) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_wire: max_ref_span must be set (ga_wire_pack_sizes reports it)");
    if (W->n_qual > 0 && (!W->qual || !W->qual_reads || !W->qual_off16)) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_wire: quality arrays missing");
    HostSource src; src.W = W;
    return run_host_source(e, src, S, out, chunk_sessions, W->max_ref_span);
}

}  // extern "C"

static int run_host_source(ga_engine* e, const HostSource& src, const ga_sessions* S, ga_result* out, int64_t chunk_sessions, int32_t maxspan) {
    if (e->refs.find(src.contig_id()) == e->refs.end()) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_host: reference contig was not uploaded");
    GA_CUDA(cudaSetDevice(e->device));
    int rc = ensure_slots(e); if (rc) return rc;
    e->last_h2d = 0; e->last_d2h = 0;
    ga_totals acc; memset(&acc, 0, sizeof acc);
    const int32_t ns = S->n_sessions;
    if (ns > 0) memset(out->sess_counts, 0, 16ull * (size_t)ns);
    if (ns == 0 || src.n_reads() == 0) { *out->totals = acc; return GA_OK; }
    for (int32_t s = 1; s < ns; ++s)
        if (S->first[s] < S->first[s - 1]) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_run_host: sessions must be sorted by first");
    if (chunk_sessions <= 0) chunk_sessions = 4096;
    const int32_t step = (int32_t)std::min<int64_t>(chunk_sessions, ns);
    struct Pending { int32_t s0, s1; int64_t cap_rec, cap_seq, cap_qual; };
    int lane = 0;
    int status = GA_OK;
    Pending inflight[kLanes] = {}; bool has[kLanes] = {};
    const int64_t total_units = src.R ? src.R->seq4_bytes / 16 : (int64_t)src.W->dir[src.W->n_blocks].unit;
    auto default_caps = [&](int32_t s0, int32_t s1, Pending* p) {
        // generous first guess from the slice size; an overflowing chunk is re-run with the exact need
        int64_t max_last = S->last[s0];
        for (int32_t s = s0 + 1; s < s1; ++s) max_last = std::max<int64_t>(max_last, S->last[s]);
        const int64_t lo_v = (int64_t)S->first[s0] - maxspan + 1;
        int64_t t_lo, t_hi, n_lo, n_hi, blocks[4];
        chunk_read_ranges(src, lo_v, max_last, &t_lo, &t_hi, &n_lo, &n_hi, blocks);
        const int64_t nr = (t_hi - t_lo) + (n_hi - n_lo);
        p->s0 = s0; p->s1 = s1;
        p->cap_rec = nr / 3 + 1024;
        const int64_t units = std::max<int64_t>(1, total_units / std::max<int64_t>(1, src.n_reads()));
        p->cap_seq = p->cap_rec * (units + 1);
        p->cap_qual = p->cap_rec * (units + 1) / 2 + 1024;
    };
    auto finish = [&](int l) -> int {
        if (!has[l]) return GA_OK;
        bool retry = false;
        int r = collect_chunk(e, l, out, &acc, &retry);
        has[l] = false;
        if (r) return r;
        if (retry) {
            const ga_totals t = *e->slots[l].h_totals;
            Pending p = inflight[l];
            p.cap_rec = (int64_t)t.n_modified + 16; p.cap_seq = (int64_t)t.seq16_used + 16; p.cap_qual = (int64_t)t.qual16_used + 16;
            r = submit_chunk(e, l, src, S, p.s0, p.s1, maxspan, p.cap_rec, p.cap_seq, p.cap_qual);
            if (r) return r;
            inflight[l] = p;
            bool again = false;
            r = collect_chunk(e, l, out, &acc, &again);
            if (r) return r;
            if (again) return ga_fail(e, GA_ERR_CAPACITY, "ga_run_host: chunk overflowed twice");
        }
        return GA_OK;
    };
    for (int32_t s0 = 0; s0 < ns && status == GA_OK; s0 += step) {
        const int32_t s1 = std::min<int32_t>(ns, s0 + step);
        // results are appended in chunk order: the oldest chunk in flight is collected after this one has been
        // queued, so its download overlaps the upload and the kernels of the younger ones
        status = finish(lane);                              // lane reuse: the chunk from kLanes steps ago
        if (status) break;
        Pending p; default_caps(s0, s1, &p);
        status = submit_chunk(e, lane, src, S, s0, s1, maxspan, p.cap_rec, p.cap_seq, p.cap_qual);
        if (status) break;
        inflight[lane] = p; has[lane] = true;
        lane = (lane + 1) % kLanes;
        status = finish(lane);                              // the oldest chunk still in flight
    }
    for (int k = 0; k < kLanes && status == GA_OK; ++k) { status = finish(lane); lane = (lane + 1) % kLanes; }
    for (int l = 0; l < kLanes; ++l) if (e->slots[l].busy) { cudaStreamSynchronize(e->slots[l].st); e->slots[l].busy = false; }
    *out->totals = acc;
    if (status) return status;
    if (acc.error == GA_ERR_CAPACITY) return ga_fail(e, GA_ERR_CAPACITY, "ga_run_host: caller result capacity too small (totals hold the need)");
    return GA_OK;
}
