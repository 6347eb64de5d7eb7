"""Synthetic tumor/normal session batches of the BASELINE.json shapes (wrapper of csrc/ga_synth.cu).

`generate_device` builds a shard of windows directly in HBM (any size, any GPU, bit-reproducible);
`generate_host` runs the identical arithmetic on the host for the small shapes the CPU tests use.
Benchmark / test input only - nothing here is on the masking path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, replace
from typing import Optional, Tuple

import numpy as np

from . import _abi, _lib
from .batch import ReadBatch, SessionTable


@dataclass(frozen=True)
class SynthConfig:
    name: str
    contig_len: int
    total_windows: int
    cov_tumor: float
    cov_normal: float
    read_len: int = 150
    seed: int = 20261018
    window_half: int = 1000
    snp_rate: float = 1e-3
    indel_rate: float = 1e-4
    err_rate: float = 1e-3
    n_rate: float = 1e-4
    somatic_vaf: float = 0.3
    clip_frac: float = 0.0
    max_indel: int = 10
    max_clip: int = 50
    depth_var_pct: int = 0          # > 0: every window keeps a uniformly drawn fraction in [1 - v/100, 1] of its reads (the others are flagged 0x4)

    def params(self, window_begin: int = 0, n_windows: Optional[int] = None) -> _abi.GaSynthParams:
        p = _abi.GaSynthParams()
        p.contig_len, p.seed, p.read_len = self.contig_len, self.seed, self.read_len
        p.total_windows, p.window_begin = self.total_windows, window_begin
        p.n_windows = self.total_windows - window_begin if n_windows is None else n_windows
        p.window_half, p.max_indel, p.max_clip, p.depth_var_pct = self.window_half, self.max_indel, self.max_clip, self.depth_var_pct
        p.cov_tumor, p.cov_normal = self.cov_tumor, self.cov_normal
        p.snp_rate, p.indel_rate, p.err_rate, p.n_rate = self.snp_rate, self.indel_rate, self.err_rate, self.n_rate
        p.somatic_vaf, p.clip_frac = self.somatic_vaf, self.clip_frac
        return p

    def plan(self, window_begin: int = 0, n_windows: Optional[int] = None) -> _abi.GaSynthPlan:
        pl = _abi.GaSynthPlan()
        p = self.params(window_begin, n_windows)
        st = _lib.synth_lib().ga_synth_plan_sizes(C.byref(p), C.byref(pl))
        if st != _abi.GA_OK:
            raise ValueError(f"invalid synthetic configuration {self.name}: windows do not fit the contig")
        return pl


# The shapes BASELINE.json names.  Coverage counts session reads only: reads outside
# the variant windows never reach the masking path.
WORKLOADS = {
    # configs[0]: chr22 plumbing case, 1M reads per dataset ~ 2.95x, 1k SNVs
    "chr22-1k": SynthConfig("chr22-1k", 50_818_468, 1_000, 2.95, 2.95),
    # configs[1]: 30x tumor / 30x normal chr1, 50k SNVs (the metric's configuration)
    "chr1-30x-50k": SynthConfig("chr1-30x-50k", 248_956_422, 50_000, 30.0, 30.0),
    # configs[2]: CIGAR-walk stress - indel- and soft-clip-heavy reads
    "cigar-stress": SynthConfig("cigar-stress", 248_956_422, 50_000, 30.0, 30.0, indel_rate=2.5e-3, clip_frac=0.2,
                                max_indel=20),
    # configs[3] at 1/10 scale: high variant density, 60x tumor / 30x normal
    "dense-60x30x": SynthConfig("dense-60x30x", 310_000_000, 100_000, 60.0, 30.0),
    # capacity check: a realistic 1 % mismatch rate at 60x / 30x - an order of magnitude more SNV
    # candidates per window than the error model of the other workloads
    "noisy-60x30x": SynthConfig("noisy-60x30x", 62_000_000, 20_000, 60.0, 30.0, err_rate=1e-2),
    # load imbalance: the depth of a window varies between 6x and 60x per dataset, window by window
    "varied-depth": SynthConfig("varied-depth", 124_000_000, 40_000, 60.0, 60.0, depth_var_pct=90),
    # small shapes for tests
    "tiny": SynthConfig("tiny", 400_000, 40, 30.0, 30.0, indel_rate=4e-4, clip_frac=0.1),
    "tiny-stress": SynthConfig("tiny-stress", 300_000, 24, 20.0, 25.0, read_len=100, indel_rate=3e-3, clip_frac=0.3,
                               max_indel=16, snp_rate=3e-3),
    "tiny-varied": SynthConfig("tiny-varied", 400_000, 40, 40.0, 40.0, indel_rate=1e-3, clip_frac=0.1, depth_var_pct=90),
}


# GRCh38 primary assembly lengths: the contigs of the whole-genome workloads
HUMAN_CONTIGS = (("chr1", 248_956_422), ("chr2", 242_193_529), ("chr3", 198_295_559), ("chr4", 190_214_555), ("chr5", 181_538_259),
                 ("chr6", 170_805_979), ("chr7", 159_345_973), ("chr8", 145_138_636), ("chr9", 138_394_717), ("chr10", 133_797_422),
                 ("chr11", 135_086_622), ("chr12", 133_275_309), ("chr13", 114_364_328), ("chr14", 107_043_718), ("chr15", 101_991_189),
                 ("chr16", 90_338_345), ("chr17", 83_257_441), ("chr18", 80_373_285), ("chr19", 58_617_616), ("chr20", 64_444_167),
                 ("chr21", 46_709_983), ("chr22", 50_818_468), ("chrX", 156_040_895), ("chrY", 57_227_415))
GENOME_LEN = sum(n for _, n in HUMAN_CONTIGS)

# Whole-genome workloads: (per-contig template, somatic variants over the whole genome).  Contig k gets the template's
# densities on its own length, round(total * len / genome) windows and seed + 7919 k, so contig 0 of "wgs-30x" IS the
# workload "chr1-30x-50k" (BASELINE config 2) and the genome is the north star's "synthetic 30x WGS tumor-normal pair";
# "wgs-60x30x" is BASELINE configs 4 / 5 at full scale: 1,000,000 somatic variants, 60x tumor / 30x normal.
GENOMES = {
    "wgs-30x": ("chr1-30x-50k", round(50_000 * GENOME_LEN / 248_956_422)),
    "wgs-60x30x": ("dense-60x30x", 1_000_000),
    # BASELINE config 3 at its stated size: the CIGAR-stress read model, 100 M session reads (116,600 windows over the genome)
    "cigar-stress-100m": ("cigar-stress", 116_600),
}


def genome_contigs(name: str):
    """The per-contig SynthConfigs of workload `name`, in genome order; a single-contig workload is a genome of one."""
    if name in WORKLOADS:
        return [WORKLOADS[name]]
    base_name, total = GENOMES[name]
    base = WORKLOADS[base_name]
    out = []
    for k, (cname, clen) in enumerate(HUMAN_CONTIGS):
        nw = base.total_windows if (clen == base.contig_len and name == "wgs-30x") else max(1, round(total * clen / GENOME_LEN))
        out.append(replace(base, name=f"{name}:{cname}", contig_len=clen, total_windows=nw, seed=base.seed + 7919 * k))
    return out


def genome_session_weights(contigs) -> np.ndarray:
    """Session reads of every session of the genome-ordered session list (what shard_sessions balances by)."""
    parts = []
    for c in contigs:
        pl = c.plan(0, 0)
        parts.append(np.full(c.total_windows, int(pl.reads_per_window[0]) + int(pl.reads_per_window[1]), np.int64))
    return np.concatenate(parts)


def genome_pieces(contigs, s_begin: int, s_end: int):
    """Sessions [s_begin, s_end) of the genome-ordered session list as per-contig pieces
    [(contig index, first window, windows, global index of the piece's first session)]."""
    out, base = [], 0
    for k, c in enumerate(contigs):
        lo, hi = max(s_begin, base), min(s_end, base + c.total_windows)
        if hi > lo:
            out.append((k, lo - base, hi - lo, lo))
        base += c.total_windows
    return out


def _check(st, what):
    if st != _abi.GA_OK:
        raise RuntimeError(f"{what} failed with ga status {st}")


def shard_windows(total_windows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous genome-ordered window range [begin, begin+n) of `rank` (sessions are the
    independent units; synthetic windows all hold the same number of reads, so an even split is balanced)."""
    base, rem = divmod(total_windows, world_size)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def generate_host(cfg: SynthConfig, window_begin: int = 0, n_windows: Optional[int] = None, with_reference: bool = True):
    """(ReadBatch, SessionTable, reference bytes | None) on the host; small shapes only (single thread)."""
    L = _lib.synth_lib()
    p = cfg.params(window_begin, n_windows)
    pl = cfg.plan(window_begin, n_windows)
    n, nw = int(pl.n_reads), int(p.n_windows)
    n_ops = np.zeros(n, np.uint32)
    has_indel = np.zeros(n, np.uint8)
    span = C.c_int32(0)
    _check(L.ga_synth_reads_count_host(C.byref(p), n_ops.ctypes.data, has_indel.ctypes.data, C.addressof(span)), "count")
    cigar_off = np.zeros(n + 1, np.uint32)
    np.cumsum(n_ops, out=cigar_off[1:])
    qual_reads = np.nonzero(has_indel)[0].astype(np.int32)
    qual_slot = np.full(n, -1, np.int32)
    qual_slot[qual_reads] = np.arange(len(qual_reads), dtype=np.int32)
    units = int(pl.units_per_read)
    b = ReadBatch(n_tumor=int(pl.n_tumor), pos=np.zeros(n, np.int32), len_flag=np.zeros(n, np.uint32),
                  seq_off16=np.zeros(n, np.uint32), cigar_off=cigar_off, cigar=np.zeros(int(cigar_off[-1]) + 4, np.uint32),
                  seq4=np.zeros(int(pl.seq4_bytes) + 64, np.uint8), qual=np.zeros(len(qual_reads) * 32 * units + 64, np.uint8),
                  max_ref_span=0, contig_id=0, qual_reads=qual_reads,
                  qual_off16=(np.arange(len(qual_reads), dtype=np.uint32) * units).astype(np.uint32))
    keep = []
    R = b.as_struct(keep)
    _check(L.ga_synth_reads_fill_host(C.byref(p), C.byref(R), qual_slot.ctypes.data), "fill")
    b.max_ref_span = int(span.value)
    b.cigar = b.cigar[:int(cigar_off[-1])]
    b.seq4 = b.seq4[:int(pl.seq4_bytes)]
    b.qual = b.qual[:len(qual_reads) * 32 * units]
    i32 = lambda: np.zeros(max(1, nw), np.int32)
    first, last, kt, kp, ke, kl = i32(), i32(), i32(), i32(), i32(), i32()
    koff = np.zeros(nw + 1, np.uint32)
    kall = np.zeros(nw + 1, np.uint8)
    _check(L.ga_synth_sessions_host(C.byref(p), first.ctypes.data, last.ctypes.data, kt.ctypes.data, kp.ctypes.data,
                                    ke.ctypes.data, kl.ctypes.data, koff.ctypes.data, kall.ctypes.data), "sessions")
    sess = SessionTable(first[:nw], last[:nw], kt[:nw], kp[:nw], ke[:nw], kl[:nw], koff, kall)
    ref = None
    if with_reference:
        buf = np.zeros(cfg.contig_len, np.uint8)
        _check(L.ga_synth_reference_host(C.byref(p), buf.ctypes.data, 0, cfg.contig_len), "reference")
        ref = buf.tobytes()
    return b, sess, ref


def generate_device(cfg: SynthConfig, device, window_begin: int = 0, n_windows: Optional[int] = None):
    """(DeviceBatch, DeviceSessions) of a window shard, generated in HBM on `device`."""
    import torch
    from .engine import DeviceBatch, DeviceSessions
    L = _lib.synth_lib()
    p = cfg.params(window_begin, n_windows)
    pl = cfg.plan(window_begin, n_windows)
    n, nw, units = int(pl.n_reads), int(p.n_windows), int(pl.units_per_read)
    with torch.cuda.device(device):
        st = torch.cuda.current_stream().cuda_stream
        i32 = dict(dtype=torch.int32, device=device)
        n_ops = torch.zeros(n, **i32)
        has_indel = torch.zeros(n, dtype=torch.uint8, device=device)
        span = torch.zeros(1, **i32)
        _check(L.ga_synth_reads_count(C.byref(p), n_ops.data_ptr(), has_indel.data_ptr(), span.data_ptr(), st), "count")
        cigar_off = torch.zeros(n + 1, **i32)
        torch.cumsum(n_ops, 0, dtype=torch.int32, out=cigar_off[1:])
        n_cigar = int(cigar_off[-1].item())
        del n_ops
        qual_reads = torch.nonzero(has_indel).flatten().to(torch.int32)
        nq = int(qual_reads.numel())
        qual_slot = torch.full((n,), -1, **i32)
        qual_slot[qual_reads.long()] = torch.arange(nq, **i32)
        del has_indel
        db = object.__new__(DeviceBatch)
        db.n_reads, db.n_tumor, db.contig_id = n, int(pl.n_tumor), 0
        db.pos = torch.empty(n, **i32)
        db.len_flag = torch.empty(n, **i32)
        db.seq_off16 = torch.empty(n, **i32)
        db.cigar_off = cigar_off
        db.cigar = torch.zeros(n_cigar + 4, **i32)
        db.seq4 = torch.zeros(int(pl.seq4_bytes) + 64, dtype=torch.uint8, device=device)
        db.qual = torch.zeros(nq * 32 * units + 64, dtype=torch.uint8, device=device)
        db.qual_reads = qual_reads
        db.qual_off16 = (torch.arange(nq, **i32) * units).to(torch.int32)
        db.seq4_bytes = int(pl.seq4_bytes)
        db.max_ref_span = 0
        R = db.as_struct()
        _check(L.ga_synth_reads_fill(C.byref(p), C.byref(R), qual_slot.data_ptr(), st), "fill")
        db.max_ref_span = int(span.item())
        del qual_slot
        ds = object.__new__(DeviceSessions)
        ds.n_sessions = nw
        z = lambda k=0: torch.zeros(max(1, nw) + k, **i32)
        ds.first, ds.last, ds.keep_type, ds.keep_pos, ds.keep_end, ds.keep_len = z(), z(), z(), z(), z(), z()
        ds.keep_allele_off = z(1)
        ds.keep_alleles = torch.zeros(nw + 1, dtype=torch.uint8, device=device)
        _check(L.ga_synth_sessions(C.byref(p), ds.first.data_ptr(), ds.last.data_ptr(), ds.keep_type.data_ptr(),
                                   ds.keep_pos.data_ptr(), ds.keep_end.data_ptr(), ds.keep_len.data_ptr(),
                                   ds.keep_allele_off.data_ptr(), ds.keep_alleles.data_ptr(), st), "sessions")
        torch.cuda.current_stream().synchronize()
    return db, ds


def reference_device(cfg: SynthConfig, device, begin: int = 0, n: Optional[int] = None):
    """ASCII reference bases [begin, begin+n) as a uint8 CUDA tensor."""
    import torch
    n = cfg.contig_len - begin if n is None else n
    with torch.cuda.device(device):
        out = torch.empty(n, dtype=torch.uint8, device=device)
        p = cfg.params(0, 0)
        _check(_lib.synth_lib().ga_synth_reference(C.byref(p), out.data_ptr(), begin, n, torch.cuda.current_stream().cuda_stream),
               "reference")
        torch.cuda.current_stream().synchronize()
    return out


def scaled(cfg: SynthConfig, total_windows: int) -> SynthConfig:
    """Same densities on a proportionally shorter contig (for bounded CPU samples and tests)."""
    frac = total_windows / cfg.total_windows
    return replace(cfg, total_windows=total_windows, contig_len=max(int(cfg.contig_len * frac), 8 * (2 * cfg.window_half + 2 * cfg.read_len + 128)),
                   name=f"{cfg.name}/{total_windows}w")
