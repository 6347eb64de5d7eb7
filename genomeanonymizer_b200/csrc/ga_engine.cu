// ga_engine.cu - engine handle, reference upload, session assignment and the C ABI of include/ga_b200.h.
// Built for sm_100a only; there is no CPU fallback: ga_engine_create fails without a CUDA device.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <map>
#include <string>
#include <vector>
#include <algorithm>

#include "ga_engine_internal.h"
#include "ga_emit_kernel.cuh"

namespace ga {

// ------------------------------------------------------------------ small kernels

// ASCII reference -> resident 4-bit codes, upper-cased (variation_classifier.py:89,194 apply .upper()).
// Output word w (w >= 1) holds bases [8(w-1), 8(w-1)+8); word 0 and the tail are N padding.
__global__ void pack_reference_kernel(const uint8_t* __restrict__ asc, int64_t n, uint32_t* __restrict__ out, int64_t n_words) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t v = 0xffffffffu;
    if (w >= 1) {
        v = 0;
        const int64_t p0 = (w - 1) * 8;
        for (int k = 0; k < 8; ++k) {
            uint32_t code = 15u;
            if (p0 + k < n) {
                uint8_t ch = asc[p0 + k];
                if (ch >= 'a' && ch <= 'z') ch -= 32;
                switch (ch) {
                    case '=': code = 0; break;  case 'A': code = 1; break;  case 'C': code = 2; break;  case 'M': code = 3; break;
                    case 'G': code = 4; break;  case 'R': code = 5; break;  case 'S': code = 6; break;  case 'V': code = 7; break;
                    case 'T': code = 8; break;  case 'W': code = 9; break;  case 'Y': code = 10; break; case 'H': code = 11; break;
                    case 'K': code = 12; break; case 'D': code = 13; break; case 'B': code = 14; break; default: code = 15; break;
                }
            }
            v |= code << (k * 4);
        }
    }
    out[w] = v;
}

// max over reads of the reference span (only when the caller did not provide ga_reads.max_ref_span)
__global__ void max_span_kernel(const uint32_t* __restrict__ cigar_off, const uint32_t* __restrict__ cigar, int64_t n, int32_t* out) {
    int m = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        m = max(m, ref_span_of(cigar, cigar_off[r], cigar_off[r + 1]));
    for (int d = 16; d; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}


// K-assign: binary search of the coordinate-sorted read positions for every session
// (replaces the index fetch inside pileup(), pileup_io.pyx:12-17).
// Lower bound of v in pos[b, e) by K lanes (the K-lane group of a warp that holds `lane`; every lane of the group gets the
// answer): each round the lanes test K pivots that cut the range into K + 1 parts, so with K = 8 a search over 2^25
// positions is 8 dependent loads deep instead of 25 (K = 1 is the binary search) - the chain of dependent loads is the
// whole cost of assign_sessions_kernel.
template <int K>
__device__ __forceinline__ int32_t lower_bound_pos_k(const int32_t* __restrict__ pos, int64_t b, int64_t e, int64_t v, int lane) {
    const int j = lane % K, g0 = lane - j;
    for (;;) {
        const bool live = b < e;                                      // uniform over the group; the groups of a warp finish at different rounds
        if (!__any_sync(0xffffffffu, live)) break;
        const int64_t n = e - b;
        const int64_t m = b + (n * (j + 1)) / (K + 1);                // b <= m < e
        const bool below = live && (int64_t)__ldg(pos + m) < v;       // monotone in j: the array is sorted
        const int c = __popc((__ballot_sync(0xffffffffu, below) >> g0) & ((1u << K) - 1u));
        if (live) {
            const int64_t nb = c ? b + (n * c) / (K + 1) + 1 : b;     // the last pivot below v, plus one
            const int64_t ne = c < K ? b + (n * (c + 1)) / (K + 1) : e;   // the first pivot not below v
            b = nb; e = ne;
        }
    }
    return (int32_t)b;
}

// 4 * K lanes per session, K lanes per search: the four searches of the read positions run side by side, then the four
// searches of the sparse quality index; the session's first lane does the rest.  The engine picks the largest K whose
// threads still fit the device in one wave: few sessions are searched wide and shallow, many sessions one lane each.
template <int K>
__global__ void assign_sessions_kernel(BatchView B, SessView S, const int32_t* __restrict__ maxspan_p, SessionDesc* __restrict__ descs,
                                       int32_t* __restrict__ big_list, int32_t* __restrict__ n_big) {
    constexpr int T = 4 * K;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, sl = lane % T, q4 = sl / K, sb = lane - sl;
    const int64_t s_raw = gid / T;
    const bool active = s_raw < S.n_sessions;
    const int s = active ? (int)s_raw : S.n_sessions - 1;
    const int maxspan = max(1, *maxspan_p);
    const int64_t first = S.first[s], last = S.last[s];
    SessionDesc d;
    {
        const bool normal = q4 >= 2;
        const int mine = lower_bound_pos_k<K>(B.pos, normal ? B.n_tumor : 0, normal ? B.n_reads : B.n_tumor, (q4 & 1) ? last : first - maxspan + 1, lane);
        d.t_begin = __shfl_sync(0xffffffffu, mine, sb); d.t_end = __shfl_sync(0xffffffffu, mine, sb + K);
        d.n_begin = __shfl_sync(0xffffffffu, mine, sb + 2 * K); d.n_end = __shfl_sync(0xffffffffu, mine, sb + 3 * K);
    }
    d.qt_begin = d.qt_end = d.qn_begin = d.qn_end = 0;
    if (B.qual_reads) {                                               // warp-uniform
        const int target = q4 == 0 ? d.t_begin : q4 == 1 ? d.t_end : q4 == 2 ? d.n_begin : d.n_end;
        const int mine = lower_bound_pos_k<K>(B.qual_reads, 0, B.n_qual, target, lane);
        d.qt_begin = __shfl_sync(0xffffffffu, mine, sb); d.qt_end = __shfl_sync(0xffffffffu, mine, sb + K);
        d.qn_begin = __shfl_sync(0xffffffffu, mine, sb + 2 * K); d.qn_end = __shfl_sync(0xffffffffu, mine, sb + 3 * K);
    }
    if (!active || sl != 0) return;
    int64_t lo = last;
    if (d.t_end > d.t_begin) lo = min(lo, (int64_t)B.pos[d.t_begin]);
    if (d.n_end > d.n_begin) lo = min(lo, (int64_t)B.pos[d.n_begin]);
    d.col_begin = (int32_t)lo;
    d.n_cols = (int32_t)max((int64_t)0, last - 1 + maxspan - lo + 1);
    const int n_range = (d.t_end - d.t_begin) + (d.n_end - d.n_begin);
    int64_t ops = 0;
    if (d.t_end > d.t_begin) ops += (int64_t)B.cigar_off[d.t_end] - B.cigar_off[d.t_begin];
    if (d.n_end > d.n_begin) ops += (int64_t)B.cigar_off[d.n_end] - B.cigar_off[d.n_begin];
    d.obs_bound = (int32_t)max((int64_t)0, ops - n_range);
    // oversize sessions go to the global-scratch kernel; table overflows found at run time join them
    d.big = (d.n_cols > kCols2 || n_range > kReads2) ? 1 : 0;
    if (n_range == 0) { d.n_cols = 0; d.big = 0; }
    d.t_seq_lo = d.t_seq_n = d.n_seq_lo = d.n_seq_n = 0u;
    d.t_cig_lo = d.t_cig_n = d.n_cig_lo = d.n_cig_n = 0u;
    if (d.t_end > d.t_begin) {
        const uint32_t L = B.len_flag[d.t_end - 1] & 0xffffu;
        d.t_seq_lo = B.seq_off16[d.t_begin];
        const uint32_t hi = B.seq_off16[d.t_end - 1] + (L ? (L + 31u) / 32u : 1u);
        d.t_seq_n = hi > d.t_seq_lo ? hi - d.t_seq_lo : 0u;
        d.t_cig_lo = B.cigar_off[d.t_begin]; d.t_cig_n = B.cigar_off[d.t_end] - d.t_cig_lo;
    }
    if (d.n_end > d.n_begin) {
        const uint32_t L = B.len_flag[d.n_end - 1] & 0xffffu;
        d.n_seq_lo = B.seq_off16[d.n_begin];
        const uint32_t hi = B.seq_off16[d.n_end - 1] + (L ? (L + 31u) / 32u : 1u);
        d.n_seq_n = hi > d.n_seq_lo ? hi - d.n_seq_lo : 0u;
        d.n_cig_lo = B.cigar_off[d.n_begin]; d.n_cig_n = B.cigar_off[d.n_end] - d.n_cig_lo;
    }
    descs[s] = d;
    if (d.big) { big_list[atomicAdd(n_big, 1)] = s; atomicAdd(n_big + 4, 1); }
}

__global__ void clear_kernel(ga_totals* totals, int32_t* n_big, unsigned int* tickets, int32_t* maxspan, int32_t given_span) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        memset(totals, 0, sizeof(ga_totals));
        *n_big = 0; tickets[0] = 0; tickets[1] = 0; *maxspan = given_span;
        for (int k = 4; k < 32; ++k) n_big[k] = 0;                    // fallback reasons, n_large, n_special, second fallback ticket, n_many, n_many_recs
    }
}

}  // namespace ga

// ------------------------------------------------------------------ engine

// the two instantiations of the one-warp resolve kernel (ga_resolve_kernel.cuh)
static const auto kResolveLean = ga::resolve_warp_kernel<ga::kReadsL, ga::kModL, ga::kObsL, ga::kEntL, ga::kLeanWarps, 11, false, 1>;
static const auto kResolveMid = ga::resolve_warp_kernel<ga::kReadsM, ga::kModM, ga::kObsM, ga::kEntM, ga::kMidTeam, 7, true, ga::kMidTeam>;

int ga_fail(ga_engine* e, int code, const char* what, cudaError_t ce) {
    if (e) {
        e->err = what;
        if (ce != cudaSuccess) { e->err += ": "; e->err += cudaGetErrorString(ce); }
    }
    return code;
}
static int fail(ga_engine* e, int code, const char* what, cudaError_t ce = cudaSuccess) { return ga_fail(e, code, what, ce); }

extern "C" {

int ga_abi_version(void) { return GA_ABI_VERSION; }

const char* ga_status_string(int status) {
    switch (status) {
        case GA_OK: return "ok";
        case GA_ERR_BAD_ARGUMENT: return "bad argument";
        case GA_ERR_OFFSET_RANGE: return "offset out of range";
        case GA_ERR_LENGTH_MISMATCH: return "length mismatch";
        case GA_ERR_CUDA: return "CUDA error";
        case GA_ERR_CAPACITY: return "output or scratch capacity exceeded";
        case GA_ERR_UNSUPPORTED: return "unsupported input";
        case GA_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}

int ga_engine_create(int device, ga_engine** out) {
    if (!out) return GA_ERR_BAD_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return GA_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GA_ERR_NO_DEVICE;
    if (prop.major < 10) return GA_ERR_NO_DEVICE;             // kernels are built for sm_100a only
    ga_engine* e = new ga_engine();
    e->device = device;
    e->n_sm = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess) { delete e; return GA_ERR_CUDA; }
    for (int l = 0; l < kLanes; ++l) {
        if (cudaMalloc(&e->lanes[l].d_small, 128) != cudaSuccess) { ga_engine_destroy(e); return GA_ERR_CUDA; }
        for (int j = 0; j < 5; ++j) for (int k = 0; k < kTimedRuns; ++k) cudaEventCreate(&e->lanes[l].ev[j][k]);
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&e->lanes[l].side, cudaStreamNonBlocking, hi);
        cudaEventCreateWithFlags(&e->lanes[l].ev_fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e->lanes[l].ev_join, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&e->lanes[l].ev_done, cudaEventDisableTiming);
    }
    cudaFuncSetAttribute(ga::scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ga::WarpSmem) * (ga::kScanThreads / 32)));
    cudaFuncSetAttribute(ga::scan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(ga::resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ga::SmemR));
    cudaFuncSetAttribute(ga::resolve_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(ga::session_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ga::SmemLayout));
    cudaFuncSetAttribute(kResolveLean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ga::SmemL) * ga::kLeanWarps));
    cudaFuncSetAttribute(kResolveLean, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(kResolveMid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ga::SmemM));
    cudaFuncSetAttribute(kResolveMid, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // persistent kernels: exactly one wave of resident CTAs (a partial second wave would wait for the first to drain)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->occ_scan, ga::scan_kernel, ga::kScanThreads, sizeof(ga::WarpSmem) * (ga::kScanThreads / 32));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->occ_lean, kResolveLean, 32 * ga::kLeanWarps, sizeof(ga::SmemL) * ga::kLeanWarps);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->occ_mid, kResolveMid, 32 * ga::kMidTeam, sizeof(ga::SmemM));
    if (e->occ_mid < 1) e->occ_mid = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->occ_res, ga::resolve_kernel, ga::kResThreads, sizeof(ga::SmemR));
    cudaFuncSetAttribute(ga::emit_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(ga::RecWarp) * ga::kRecWarps));
    cudaFuncSetAttribute(ga::emit_records_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e->occ_rec, ga::emit_records_kernel, 32 * ga::kRecWarps, sizeof(ga::RecWarp) * ga::kRecWarps);
    if (e->occ_rec < 1) e->occ_rec = 1;
    if (const char* v = getenv("GA_OCC_SCAN")) e->occ_scan = std::min(e->occ_scan, std::max(1, atoi(v)));   // tuning knob: CTAs per SM of the scan kernel
    if (e->occ_scan < 1) e->occ_scan = 1;
    if (e->occ_lean < 1) e->occ_lean = 1;
    if (e->occ_res < 1) e->occ_res = 1;
    *out = e;
    return GA_OK;
}

void ga_engine_destroy(ga_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    ga_host_slots_destroy(e);
    for (auto& kv : e->refs) cudaFree(kv.second.d_ref4);
    cudaFree(e->d_fastq_sums);
    for (int l = 0; l < kLanes; ++l) {
        Lane& L = e->lanes[l];
        cudaFree(L.d_descs); cudaFree(L.d_big_list); cudaFree(L.d_large_list); cudaFree(L.d_large2_list); cudaFree(L.d_small); cudaFree(L.d_big_scratch);
        if (L.side) cudaStreamDestroy(L.side);
        if (L.ev_fork) cudaEventDestroy(L.ev_fork);
        if (L.ev_join) cudaEventDestroy(L.ev_join);
        if (L.ev_done) cudaEventDestroy(L.ev_done);
        cudaFree(L.d_kind); cudaFree(L.d_edesc); cudaFree(L.d_special); cudaFree(L.d_rare_list); cudaFree(L.d_edit_keep); cudaFree(L.d_many); cudaFree(L.d_many_recs); cudaFree(L.d_germ); cudaFree(L.d_ent); cudaFree(L.d_obs); cudaFree(L.d_cnt);
        for (int j = 0; j < 5; ++j) for (int k = 0; k < kTimedRuns; ++k) if (L.ev[j][k]) cudaEventDestroy(L.ev[j][k]);
    }
    delete e;
}

const char* ga_last_error(const ga_engine* e) { return e ? e->err.c_str() : "null engine"; }
int64_t ga_launch_count(const ga_engine* e) { return e ? e->launches : 0; }

int ga_upload_reference(ga_engine* e, int contig_id, const uint8_t* bases, int64_t n_bases, void* stream_) {
    if (!e || !bases || n_bases < 0) return fail(e, GA_ERR_BAD_ARGUMENT, "ga_upload_reference: null argument");
    cudaStream_t st = (cudaStream_t)stream_;
    GA_CUDA(cudaSetDevice(e->device));
    cudaPointerAttributes attr;
    const uint8_t* d_asc = bases;
    uint8_t* tmp = nullptr;
    bool is_dev = (cudaPointerGetAttributes(&attr, bases) == cudaSuccess) && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    // every failure path releases the staging copy and leaves the entry consistent (no pointer without its length)
#define GA_UP(call) do { cudaError_t _ce = (call); if (_ce != cudaSuccess) { if (tmp) cudaFree(tmp); return ga_fail(e, GA_ERR_CUDA, #call, _ce); } } while (0)
    if (!is_dev) {
        GA_UP(cudaMalloc(&tmp, (size_t)n_bases + 16));
        GA_UP(cudaMemcpyAsync(tmp, bases, (size_t)n_bases, cudaMemcpyHostToDevice, st));
        d_asc = tmp;
    }
    RefEntry& re = e->refs[contig_id];
    const int64_t n_words = (n_bases + 7) / 8 + 1 + 8;          // 1 pad word in front, 8 behind
    if (n_words > re.cap_words) {                               // the resident copy only grows (one session per call re-uploads often)
        if (re.d_ref4) { cudaFree(re.d_ref4); re.d_ref4 = nullptr; }
        re.n = 0; re.cap_words = 0;
        const int64_t want = n_words + n_words / 4;
        GA_UP(cudaMalloc(&re.d_ref4, (size_t)want * 4));
        re.cap_words = want;
    }
    re.n = n_bases;
    const int threads = 256;
    ga::pack_reference_kernel<<<(unsigned)((n_words + threads - 1) / threads), threads, 0, st>>>(d_asc, n_bases, re.d_ref4, n_words);
    e->launches++;
    GA_UP(cudaGetLastError());
    if (tmp) GA_UP(cudaStreamSynchronize(st));
#undef GA_UP
    if (tmp) cudaFree(tmp);
    return GA_OK;
}

static int ensure_session_scratch(ga_engine* e, Lane& L, int64_t n_sessions) {
    if (n_sessions <= L.cap_sessions) return GA_OK;
    cudaFree(L.d_descs); cudaFree(L.d_big_list); cudaFree(L.d_large_list); cudaFree(L.d_large2_list);
    L.d_descs = nullptr; L.d_big_list = nullptr; L.d_large_list = nullptr; L.d_large2_list = nullptr; L.cap_sessions = 0;
    const int64_t cap = n_sessions + n_sessions / 4 + 1024;
    GA_CUDA(cudaMalloc(&L.d_descs, (size_t)cap * sizeof(ga::SessionDesc)));
    GA_CUDA(cudaMalloc(&L.d_big_list, (size_t)cap * sizeof(int32_t)));
    GA_CUDA(cudaMalloc(&L.d_large_list, (size_t)cap * sizeof(int32_t)));
    GA_CUDA(cudaMalloc(&L.d_large2_list, (size_t)cap * sizeof(int32_t)));
    L.cap_sessions = cap;
    return GA_OK;
}

static int ensure_stream_scratch(ga_engine* e, Lane& L, int64_t cap_records, int64_t n_sessions) {
    if (cap_records > L.cap_kind) {
        cudaFree(L.d_kind); cudaFree(L.d_edesc); cudaFree(L.d_special); cudaFree(L.d_rare_list); cudaFree(L.d_edit_keep); L.d_edit_keep = nullptr; L.d_kind = nullptr; L.d_edesc = nullptr; L.d_special = nullptr; L.d_rare_list = nullptr; L.cap_kind = 0;
        const int64_t cap = cap_records + cap_records / 8 + 1024;
        GA_CUDA(cudaMalloc(&L.d_kind, (size_t)cap * sizeof(uint32_t)));
        GA_CUDA(cudaMalloc(&L.d_edesc, (size_t)cap * sizeof(uint4)));
        GA_CUDA(cudaMalloc(&L.d_special, (size_t)cap * 4 * sizeof(uint4)));
        GA_CUDA(cudaMalloc(&L.d_rare_list, (size_t)cap * sizeof(uint32_t)));
        cudaFree(L.d_many); cudaFree(L.d_many_recs); L.d_many = nullptr; L.d_many_recs = nullptr; L.cap_many = 0;
        const int64_t cap_many = std::min<int64_t>(std::max<int64_t>(1 << 16, cap / 4), 1 << 26);   // edit lists of reads with more than two germline indels
        GA_CUDA(cudaMalloc(&L.d_many, (size_t)cap_many * sizeof(uint4)));
        GA_CUDA(cudaMalloc(&L.d_many_recs, (size_t)cap_many * sizeof(uint32_t)));
        L.cap_many = cap_many;
        L.cap_kind = cap;
    }
    if (n_sessions > L.cap_germ) {
        cudaFree(L.d_germ); L.d_germ = nullptr; L.cap_germ = 0;
        const int64_t cap = n_sessions + n_sessions / 4 + 1024;
        GA_CUDA(cudaMalloc(&L.d_germ, (size_t)cap * ga::kGermStride * sizeof(uint32_t)));
        L.cap_germ = cap;
    }
    if (2 * n_sessions > L.cap_items) {
        cudaFree(L.d_ent); cudaFree(L.d_obs); cudaFree(L.d_cnt); L.d_ent = nullptr; L.d_obs = nullptr; L.d_cnt = nullptr; L.cap_items = 0;
        const int64_t cap = 2 * (n_sessions + n_sessions / 4 + 1024);
        GA_CUDA(cudaMalloc(&L.d_ent, (size_t)cap * ga::kEntHalf * sizeof(uint32_t)));
        GA_CUDA(cudaMalloc(&L.d_obs, (size_t)cap * ga::kObsHalf * sizeof(ga::ObsRec)));
        GA_CUDA(cudaMalloc(&L.d_cnt, (size_t)cap * sizeof(uint4)));
        L.cap_items = cap;
    }
    return GA_OK;
}

static int ensure_big_scratch(ga_engine* e, Lane& L) {
    if (L.d_big_scratch) return GA_OK;
    const int64_t per = 4ll * e->big_cols_cap * 2 + 4ll * ((e->big_reads_cap + 31) / 32) + 20ll * e->big_obs_cap + 8ll * e->big_reads_cap;
    e->big_bytes_per_cta = (per + 255) / 256 * 256;
    e->big_ctas = e->n_sm;
    GA_CUDA(cudaMalloc(&L.d_big_scratch, (size_t)e->big_bytes_per_cta * e->big_ctas));
    return GA_OK;
}

// A run uses one lane (scratch + side stream).  Runs launched on ONE stream share a lane - stream order keeps them
// apart; a run on another stream takes another lane, so that up to kLanes runs (e.g. the contigs of a genome, one
// stream each) overlap: the persistent kernels of one run fill the SMs that the tail of another leaves idle.  With
// more streams than lanes a lane is reused and the new run waits for the lane's previous one.
int ga_run(ga_engine* e, const ga_reads* R, const ga_sessions* S, ga_result* out, void* stream_) {
    if (!e) return GA_ERR_BAD_ARGUMENT;
    cudaStream_t st = (cudaStream_t)stream_;
    int lane = -1;
    for (int l = 0; l < kLanes && lane < 0; ++l) if (e->lanes[l].used && e->lanes[l].last_stream == st) lane = l;
    for (int l = 0; l < kLanes && lane < 0; ++l) if (!e->lanes[l].used) lane = l;
    if (lane < 0) { lane = e->next_lane; e->next_lane = (e->next_lane + 1) % kLanes; }
    e->last_lane = lane;
    return ga_run_lane(e, lane, R, S, out, st);
}

// Duration (ms) between events ev[a] and ev[b] of the most recent runs, out[0] = latest.
static int stage_history(ga_engine* e, int a, int b, float* out, int n) {
    if (!e || !out || n < 0) return 0;
    Lane& L = e->lanes[e->last_lane];
    const int have = (int)std::min<int64_t>(std::min<int64_t>(L.runs, kTimedRuns), n);
    for (int k = 0; k < have; ++k) {
        const int slot = (int)((L.runs - 1 - k) % kTimedRuns);
        float ms = -1.f;
        if (cudaEventSynchronize(L.ev[b][slot]) != cudaSuccess || cudaEventElapsedTime(&ms, L.ev[a][slot], L.ev[b][slot]) != cudaSuccess) ms = -1.f;
        out[k] = ms;
    }
    return have;
}

int ga_kernel_ms_history(ga_engine* e, float* out, int n) { return stage_history(e, 0, 4, out, n); }

int ga_stage_ms_history(ga_engine* e, int stage, float* out, int n) {
    if (stage < 0 || stage > 3) return 0;
    return stage_history(e, stage, stage + 1, out, n);
}

int ga_last_fallback_sessions(ga_engine* e, int32_t* reasons, int n_reasons) {
    if (!e) return -1;
    int32_t h[24] = {0};
    if (cudaSetDevice(e->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpy(h, e->lanes[e->last_lane].d_small, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    for (int k = 0; reasons && k < n_reasons && k < 8; ++k) reasons[k] = h[4 + k];
    if (reasons && n_reasons > 8) reasons[8] = h[20];      // mid one-warp kernel -> one-CTA kernel
    if (reasons && n_reasons > 9) reasons[9] = h[12];      // lean one-warp kernel -> mid one-warp kernel
    return h[0];
}

int ga_engine_keep_edits(ga_engine* e, int on) {
    if (!e) return GA_ERR_BAD_ARGUMENT;
    e->keep_edits = on != 0;
    return GA_OK;
}

int ga_record_edits(ga_engine* e, const int64_t* rec_idx, int64_t n, uint32_t* out) {
    if (!e || (n > 0 && (!rec_idx || !out)) || n < 0) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_record_edits: null argument");
    Lane& L = e->lanes[e->last_lane];
    if (!e->keep_edits || !L.d_edit_keep) return ga_fail(e, GA_ERR_BAD_ARGUMENT, "ga_record_edits: ga_engine_keep_edits was not switched on before the run");
    GA_CUDA(cudaSetDevice(e->device));
    GA_CUDA(cudaDeviceSynchronize());
    for (int64_t k = 0; k < n;) {                                     // one copy per run of consecutive indices
        if (rec_idx[k] < 0 || rec_idx[k] >= L.cap_kind) return ga_fail(e, GA_ERR_OFFSET_RANGE, "ga_record_edits: record index out of range");
        int64_t m = 1;
        while (k + m < n && rec_idx[k + m] == rec_idx[k] + m && rec_idx[k + m] < L.cap_kind) ++m;
        GA_CUDA(cudaMemcpy(out + 8 * k, static_cast<const uint8_t*>(L.d_edit_keep) + 32 * rec_idx[k], (size_t)(32 * m), cudaMemcpyDeviceToHost));
        k += m;
    }
    return GA_OK;
}

float ga_last_kernel_ms(ga_engine* e) {
    float ms = -1.f;
    return ga_kernel_ms_history(e, &ms, 1) == 1 ? ms : -1.f;
}

}  // extern "C"

int ga_run_lane(ga_engine* e, int lane, const ga_reads* R, const ga_sessions* S, ga_result* out, cudaStream_t st) {
    if (!e || !R || !S || !out || !out->totals) return fail(e, GA_ERR_BAD_ARGUMENT, "ga_run: null argument");
    if (lane < 0 || lane >= kLanes) return fail(e, GA_ERR_BAD_ARGUMENT, "ga_run: bad lane");
    if (R->n_reads < 0 || R->n_tumor < 0 || R->n_tumor > R->n_reads || S->n_sessions < 0)
        return fail(e, GA_ERR_BAD_ARGUMENT, "ga_run: negative or inconsistent sizes");
    if (R->n_reads > 0x7fffffffll) return fail(e, GA_ERR_UNSUPPORTED, "ga_run: more than 2^31-1 reads in one batch");
    if ((reinterpret_cast<uintptr_t>(R->seq4) & 15u) || (reinterpret_cast<uintptr_t>(R->qual) & 15u) ||
        (reinterpret_cast<uintptr_t>(out->out_seq4) & 15u) || (reinterpret_cast<uintptr_t>(out->out_qual) & 15u))
        return fail(e, GA_ERR_BAD_ARGUMENT, "ga_run: seq4 / qual / out_seq4 / out_qual must be 16-byte aligned (128-bit and TMA accesses)");
    auto it = e->refs.find(R->contig_id);
    if (it == e->refs.end()) return fail(e, GA_ERR_BAD_ARGUMENT, "ga_run: reference contig was not uploaded");
    Lane& L = e->lanes[lane];
    GA_CUDA(cudaSetDevice(e->device));
    if (L.used && L.last_stream != st) GA_CUDA(cudaStreamWaitEvent(st, L.ev_done, 0));   // the lane's scratch is still the previous run's
    L.used = true; L.last_stream = st;
    struct DoneMark { Lane& L; cudaStream_t st; ~DoneMark() { cudaEventRecord(L.ev_done, st); } } done_mark{L, st};
    int rc = ensure_session_scratch(e, L, S->n_sessions); if (rc) return rc;
    rc = ensure_big_scratch(e, L); if (rc) return rc;

    ga::BatchView B;
    B.pos = R->pos; B.len_flag = R->len_flag; B.seq_off16 = R->seq_off16; B.cigar_off = R->cigar_off; B.cigar = R->cigar;
    B.seq4 = R->seq4; B.qual = R->qual; B.qual_reads = R->qual_reads; B.qual_off16 = R->qual_off16;
    B.n_reads = R->n_reads; B.n_tumor = R->n_tumor; B.n_qual = R->qual_reads ? R->n_qual : 0;
    B.ref4 = it->second.d_ref4; B.ref_len = it->second.n;
    ga::SessView V;
    V.first = S->first; V.last = S->last; V.keep_type = S->keep_type; V.keep_pos = S->keep_pos; V.keep_end = S->keep_end;
    V.keep_len = S->keep_len; V.keep_allele_off = S->keep_allele_off; V.keep_alleles = S->keep_alleles; V.n_sessions = S->n_sessions;
    ga::ResultView O;
    O.cap_records = out->cap_records; O.cap_seq16 = out->cap_seq16; O.cap_qual16 = out->cap_qual16;
    O.mod_session = out->mod_session; O.mod_read = out->mod_read; O.mod_len = out->mod_len;
    O.mod_seq_off16 = out->mod_seq_off16; O.mod_qual_off16 = out->mod_qual_off16;
    O.out_seq4 = out->out_seq4; O.out_qual = out->out_qual; O.sess_counts = out->sess_counts; O.totals = out->totals;

    int32_t* d_nbig = L.d_small;
    int32_t* d_maxspan = L.d_small + 1;
    unsigned int* d_tickets = reinterpret_cast<unsigned int*>(L.d_small + 2);
    ga::clear_kernel<<<1, 32, 0, st>>>(out->totals, d_nbig, d_tickets, d_maxspan, R->max_ref_span);
    e->launches++;
    if (S->n_sessions == 0 || R->n_reads == 0) {
        if (S->n_sessions > 0) GA_CUDA(cudaMemsetAsync(out->sess_counts, 0, 16ull * S->n_sessions, st));
        GA_CUDA(cudaGetLastError());
        return GA_OK;
    }
    if (R->max_ref_span <= 0) {
        ga::max_span_kernel<<<e->n_sm * 4, 256, 0, st>>>(R->cigar_off, R->cigar, R->n_reads, d_maxspan);
        e->launches++;
    }
    {
        const int64_t room = (int64_t)e->n_sm * 2048 / std::max<int64_t>(1, 4 * (int64_t)S->n_sessions);   // lanes per search that keep the kernel in one wave
        const auto grid = [&](int k) { return (int)((4 * k * (int64_t)S->n_sessions + 127) / 128); };
        if (room >= 8) ga::assign_sessions_kernel<8><<<grid(8), 128, 0, st>>>(B, V, d_maxspan, L.d_descs, L.d_big_list, d_nbig);
        else if (room >= 4) ga::assign_sessions_kernel<4><<<grid(4), 128, 0, st>>>(B, V, d_maxspan, L.d_descs, L.d_big_list, d_nbig);
        else if (room >= 2) ga::assign_sessions_kernel<2><<<grid(2), 128, 0, st>>>(B, V, d_maxspan, L.d_descs, L.d_big_list, d_nbig);
        else ga::assign_sessions_kernel<1><<<grid(1), 128, 0, st>>>(B, V, d_maxspan, L.d_descs, L.d_big_list, d_nbig);
    }
    e->launches++;
    ga::BigScratch scr;
    scr.base = L.d_big_scratch; scr.bytes_per_cta = e->big_bytes_per_cta;
    scr.cols_cap = e->big_cols_cap; scr.reads_cap = e->big_reads_cap; scr.obs_cap = e->big_obs_cap;
    rc = ensure_stream_scratch(e, L, out->cap_records, S->n_sessions); if (rc) return rc;
    ga::ScanScratch X; X.ent = L.d_ent; X.obs = reinterpret_cast<ga::ObsRec*>(L.d_obs); X.cnt = reinterpret_cast<uint4*>(L.d_cnt);
    ga::EmitScratch2 E; E.kind1_list = reinterpret_cast<uint32_t*>(L.d_kind); E.edesc = reinterpret_cast<uint4*>(L.d_edesc); E.germ = L.d_germ;
    E.sdesc = reinterpret_cast<uint4*>(L.d_special); E.n_special = reinterpret_cast<uint32_t*>(L.d_small + 13);
    E.many = reinterpret_cast<uint4*>(L.d_many); E.n_many = reinterpret_cast<uint32_t*>(L.d_small + 15); E.cap_many = (uint32_t)L.cap_many;
    E.many_recs = L.d_many_recs; E.n_many_recs = reinterpret_cast<uint32_t*>(L.d_small + 16); E.n_kind1 = reinterpret_cast<uint32_t*>(L.d_small + 17); E.ticket_large = reinterpret_cast<unsigned int*>(L.d_small + 18); E.ticket_lean = reinterpret_cast<unsigned int*>(L.d_small + 19);
    E.n_rare = reinterpret_cast<uint32_t*>(L.d_small + 22); E.rare_list = L.d_rare_list;
    E.edit_keep = nullptr;
    if (e->keep_edits) {                                              // asked for by the per-sample driver (quirk Q12 of DESIGN.md), small batches
        if (!L.d_edit_keep) GA_CUDA(cudaMalloc(&L.d_edit_keep, (size_t)L.cap_kind * 32));
        GA_CUDA(cudaMemsetAsync(L.d_edit_keep, 0xff, (size_t)std::max<int64_t>(out->cap_records, 1) * 32, st));
        E.edit_keep = reinterpret_cast<uint4*>(L.d_edit_keep);
    }
    const int tslot = (int)(L.runs % kTimedRuns);
    L.runs++;
    // stage 1: allele discovery, one warp per (session, dataset) item, persistent CTAs
    const int64_t n_items = 2 * (int64_t)S->n_sessions;
    const int grid_scan = (int)std::min<int64_t>((int64_t)e->n_sm * e->occ_scan, (n_items + ga::kScanThreads / 32 - 1) / (ga::kScanThreads / 32));
    GA_CUDA(cudaEventRecord(L.ev[0][tslot], st));
    ga::scan_kernel<<<grid_scan, ga::kScanThreads, sizeof(ga::WarpSmem) * (ga::kScanThreads / 32), st>>>(B, V, L.d_descs, X, d_tickets, out->totals);
    GA_CUDA(cudaEventRecord(L.ev[1][tslot], st));
    // stage 2: germline set, modified-record list, output slots, headers - one warp per session, then one CTA per
    // session for those whose tables did not fit the lean capacities
    int32_t* d_nlarge = L.d_small + 12;
    const int lean_ctas = (int)std::min<int64_t>((int64_t)e->n_sm * e->occ_lean, (S->n_sessions + ga::kLeanWarps - 1) / ga::kLeanWarps);
    int32_t* d_nlarge2 = L.d_small + 20;
    kResolveLean<<<lean_ctas, 32 * ga::kLeanWarps, sizeof(ga::SmemL) * ga::kLeanWarps, st>>>(B, V, L.d_descs, L.d_big_list, d_nbig, nullptr, nullptr, E.ticket_lean,
                                                                                           L.d_large_list, d_nlarge, O, X, E);
    // sessions beyond the lean capacities (indel-dense ones): the same kernel with larger tables, then the one-CTA kernel
    const int mid_ctas = (int)std::min<int64_t>((int64_t)e->n_sm * e->occ_mid, (int64_t)S->n_sessions);
    kResolveMid<<<mid_ctas, 32 * ga::kMidTeam, sizeof(ga::SmemM), st>>>(B, V, L.d_descs, L.d_big_list, d_nbig, L.d_large_list, d_nlarge,
                                                                                        reinterpret_cast<unsigned int*>(L.d_small + 21), L.d_large2_list, d_nlarge2, O, X, E);
    const int grid_res = (int)std::min<int64_t>((int64_t)e->n_sm * e->occ_res, S->n_sessions);
    ga::resolve_kernel<<<grid_res, ga::kResThreads, sizeof(ga::SmemR), st>>>(B, V, L.d_descs, L.d_big_list, d_nbig, L.d_large2_list, d_nlarge2, O, X, E);
    GA_CUDA(cudaEventRecord(L.ev[2][tslot], st));
    // oversize sessions and whatever the tables of stages 1-2 could not hold: global-scratch kernel, complete records;
    // it runs on a side stream beside stage 3 (it only appends records, which the emission kernel skips), behind the
    // emission of the few records whose edit lists travel in the side buffer
    GA_CUDA(cudaEventRecord(L.ev_fork, st));
    GA_CUDA(cudaStreamWaitEvent(L.side, L.ev_fork, 0));
    ga::emit_many_kernel<<<e->n_sm * 4, ga::kThreads, 0, L.side>>>(B, O, E);   // records of reads with more than two germline indels (usually few)
    ga::session_kernel<false><<<e->n_sm * 2, ga::kThreads, sizeof(ga::SmemLayout), L.side>>>(B, V, L.d_descs, L.d_big_list, d_nbig, O, scr,
                                                                                               reinterpret_cast<unsigned int*>(L.d_small + 14), E.edit_keep);
    ga::session_kernel<true><<<e->big_ctas, ga::kThreads, 0, L.side>>>(B, V, L.d_descs, L.d_big_list, d_nbig, O, scr, d_tickets + 1, E.edit_keep);
    GA_CUDA(cudaEventRecord(L.ev_join, L.side));
    // stage 3: record bodies
    ga::emit_kernel<<<e->n_sm * 8, ga::kThreads, 0, st>>>(B, L.d_descs, O, E);
    ga::emit_records_kernel<<<e->n_sm * e->occ_rec, 32 * ga::kRecWarps, sizeof(ga::RecWarp) * ga::kRecWarps, st>>>(B, O, E);
    ga::emit_special_kernel<<<e->n_sm * 4, ga::kThreads, ga::kEmitSpecialSmem, st>>>(B, L.d_descs, O, E);
    GA_CUDA(cudaEventRecord(L.ev[3][tslot], st));
    GA_CUDA(cudaStreamWaitEvent(st, L.ev_join, 0));
    GA_CUDA(cudaEventRecord(L.ev[4][tslot], st));
    e->launches += 10;
    GA_CUDA(cudaGetLastError());
    return GA_OK;
}
