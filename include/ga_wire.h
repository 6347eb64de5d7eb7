/*
 * ga_wire.h - compact host-side form of a read batch for the end-to-end entry (ga_run_wire).
 *
 * ga_run_host moves 100 bytes per session read over PCIe (80-byte padded seq4 record, four meta words, one CIGAR
 * word) and the device is idle 97 % of such a step.  The wire form carries the same information in ~40 bytes:
 *   bases      2 bits each (A C G T = 0 1 2 3), the reads of a block back to back in one bit stream; any other base
 *              code (N, IUPAC, '=') travels in an exception list and is patched in on the device
 *   position   8-bit difference to the previous read of the block; larger differences (below 65,536) are listed
 *   flag       a 4-bit index into the block's dictionary of up to 16 flag values (16 bits per read when a block has more)
 *   length     the block's most frequent length once; the reads that differ are listed
 *   record offsets  not sent: exclusive scans of the lengths on the device
 *   CIGAR      only for reads whose CIGAR is not a single M op spanning the read ("generic" reads)
 * Reads are grouped in BLOCKS of at most GA_WIRE_BLOCK_READS reads of one dataset; a block is one contiguous,
 * self-describing byte range, so that the reads a chunk of sessions needs are ONE slice of the blob per dataset
 * (one cudaMemcpyAsync instead of a dozen) and one CTA expands one block into the engine's ga_reads arrays.
 * A block is closed early when a position difference does not fit 16 bits.
 *
 * Block (little-endian; starts 16-byte aligned, size a multiple of 16; format version 3):
 *   uint32 n_reads, n_words, n_gen, n_gen_ops, n_exc, len_common | n_flags << 16, n_lenx, n_posx    (32 bytes)
 *   n_flags > 0:  uint16 fdict[16]                        the block's distinct BAM flags (n_flags of them)
 *                 uint8  fidx[ceil(n_reads / 2)]          (padded to 4 bytes) two 4-bit dictionary indices per byte, low nibble first
 *   n_flags == 0: uint16 flag[n_reads]                    (padded to 4 bytes) BAM flag (more than 16 distinct flags in the block)
 *   uint8  dpos8[n_reads]           (padded to 4 bytes)   min(pos - pos of the previous read, 255); 0 for the first
 *   uint32 posx[n_posx]                                   (index in the block << 16) | difference, for every difference >= 255, ascending
 *   uint32 lenx[n_lenx]                                   (index in the block << 16) | length of every read whose length is
 *                                                         not len_common, ascending
 *   uint16 gen_idx[n_gen]           (padded to 4 bytes)   index in the block of every generic read, ascending
 *   uint32 gen_off[n_gen + 1]                             its first op in gen_cigar; the last entry is n_gen_ops
 *   uint32 gen_cigar[n_gen_ops]                           their ops, BAM encoding, read after read
 *   uint32 exc[n_exc]                                     (index in block << 20) | (query offset << 4) | base code
 *   uint32 bases2[n_words]                                base k of the block (the reads back to back) = bits 2 (k % 16) of word
 *                                                         k / 16; n_words = ceil(sum of lengths / 16) + 3 (zero words behind)
 * Qualities stay what they are in ga_reads: sparse records of the reads that have an I or D op.
 *
 * The directory (host only) lists, per block, where it starts in the blob and the running totals the host needs to
 * size and place a chunk: entry b describes block b, entry n_blocks closes the last block.  Tumor blocks come first.
 *
 * Implemented in genomeanonymizer_b200/csrc/ga_wire.cu (host packer: plain C++ threads; expansion: CUDA).
 */
#ifndef GA_WIRE_H
#define GA_WIRE_H

#include <stdint.h>
#include "ga_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GA_WIRE_BLOCK_READS 1024

typedef struct ga_wire_dir {
    uint64_t byte;      /* offset of the block in the blob                                   */
    uint32_t read;      /* batch index of its first read                                     */
    uint32_t unit;      /* seq4 units (16 bytes) of all reads before it, over the whole batch */
    uint32_t ops;       /* CIGAR ops of all reads before it                                  */
    int32_t  pos;       /* reference_start of its first read                                 */
    uint32_t reserved[2];
} ga_wire_dir;

typedef struct ga_reads_wire {
    int64_t n_reads, n_tumor;
    int64_t n_blocks, n_tumor_blocks;
    const uint8_t* blob;          /* 16-byte aligned                                          */
    int64_t blob_bytes;
    const ga_wire_dir* dir;       /* [n_blocks + 1]                                           */
    /* sparse qualities, as in ga_reads; records contiguous and ascending with qual_reads      */
    const uint8_t*  qual;
    int64_t n_qual;
    const int32_t*  qual_reads;
    const uint32_t* qual_off16;
    int64_t qual_units;           /* 32-byte units used in qual                               */
    int32_t max_ref_span;         /* > 0 (the packer computes it)                             */
    int32_t contig_id;
} ga_reads_wire;

/* Host packer, pass 1: number of blocks and blob bytes the batch needs (reads of a dataset in coordinate order, as
 * ga_run requires).  *max_ref_span receives the largest reference span (may be NULL). */
int ga_wire_pack_sizes(const ga_reads* reads, int64_t* n_blocks, int64_t* n_tumor_blocks, int64_t* blob_bytes, int32_t* max_ref_span);
/* Pass 2: writes the blob (blob_bytes bytes, 16-byte aligned) and the directory (n_blocks + 1 entries) on n_threads
 * host threads (0 = all).  Fails with GA_ERR_BAD_ARGUMENT when the batch changed between the passes. */
int ga_wire_pack(const ga_reads* reads, uint8_t* blob, int64_t blob_bytes, ga_wire_dir* dir, int64_t n_blocks, int n_threads);

/* ga_run_host over the wire form: all pointers are HOST pointers (pinned for full speed).  Same chunking, lanes,
 * result layout and error behaviour as ga_run_host; mod_read are batch indices of the ORIGINAL batch. */
int ga_run_wire(ga_engine* e, const ga_reads_wire* reads, const ga_sessions* sessions, ga_result* result, int64_t chunk_sessions);

#ifdef __cplusplus
}
#endif
#endif /* GA_WIRE_H */
