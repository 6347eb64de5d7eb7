/*
 * ga_digest.h - record digest of a masking result: the arithmetic both sides of a parity check agree on.
 *
 * A ga_result holds the modified (session, read) records in no particular order, and a sharded run holds them in
 * several results.  To compare two results - engine vs oracle, 1 GPU vs 8 GPUs - without sorting or moving
 * 100-byte records around, every record is reduced to a key (global session, global read id) and a 128-bit hash of
 * exactly the bytes that are observable (ga_b200.h "output"): key, new length, whether it carries qualities, the
 * new base codes [0, len) and, when present, the new qualities [0, len).  Padding nibbles / bytes are NOT hashed.
 * The digest of a result is the wrapping 64-bit sum of the record hashes, so it is independent of record order and
 * of how the sessions were cut into shards:
 *     digest[0] = sum hash_lo   digest[1] = sum hash_hi   digest[2] = records   digest[3] = sum of new lengths
 *
 * Record hash (w = little-endian 32-bit words):
 *     h = seed
 *     h = mix(h, contig); h = mix(h, session_global); h = mix(h, read_gid); h = mix(h, len | has_qual << 32)
 *     for k in 0 .. ceil(len/8)-1:  h = mix(h, seq4 word k, nibbles >= len cleared)
 *     if has_qual: for k in 0 .. ceil(len/4)-1:  h = mix(h, quality word k, bytes >= len cleared)
 *     hash = fin(h)
 * computed twice with (seed, multiplier) = (GA_DIGEST_SEED_LO, GA_DIGEST_MUL_LO) and (.._HI, .._HI).
 * read_gid = dataset << 40 | ordinal of the read inside its dataset, counted over the whole (unsharded) input.
 *
 * Implemented by ga_result_digest (device, csrc/ga_wire.cu) and, separately, by the oracle (oracle/ga_oracle.c).
 */
#ifndef GA_DIGEST_H
#define GA_DIGEST_H
#include <stdint.h>

#ifdef __CUDACC__
#define GA_DIGEST_FN __host__ __device__ static inline
#else
#define GA_DIGEST_FN static inline
#endif

#define GA_DIGEST_SEED_LO 0x243F6A8885A308D3ull
#define GA_DIGEST_SEED_HI 0x13198A2E03707344ull
#define GA_DIGEST_MUL_LO  0x9E3779B97F4A7C15ull
#define GA_DIGEST_MUL_HI  0xC2B2AE3D27D4EB4Full

GA_DIGEST_FN uint64_t ga_digest_mix(uint64_t h, uint64_t w, uint64_t mul) {
    h ^= w;
    h *= mul;
    h ^= h >> 29;
    return h;
}
GA_DIGEST_FN uint64_t ga_digest_fin(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}
#endif /* GA_DIGEST_H */
