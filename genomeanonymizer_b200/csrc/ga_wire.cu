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
