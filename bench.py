#!/usr/bin/env python
"""bench.py - reads/s of the per-read germline-masking hot path on synthetic tumor-normal sessions.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the whole hot path (session assignment, allele discovery, germline resolution,
masking, compaction) over every session of the workload.  `value` times ga_run() on a batch already
resident in HBM; `e2e` times ga_run_host() - host SoA buffers in pinned memory in, compacted modified
records in host memory out, H2D/D2H inside the timed region.  `roofline` is the session kernel against
the measured HBM copy peak; `cpu_baseline` is the CPU oracle (a C restatement of the reference's
algorithm, all host threads) on a bounded sample of the same sessions.

N > 1 (torchrun): region sharding of ONE synthetic whole genome, "wgs-30x" (24 contigs, 30x / 30x,
one somatic SNV per 4,979 bp; contig 0 is the N=1 workload chr1-30x-50k).  `value` is weak scaling: rank r masks
sessions [50,000 r, 50,000 (r+1)) of the genome-ordered session list; `strong` cuts the WHOLE genome's session list
across the N ranks with sharding.shard_sessions (total work fixed).  No data-path collective; the masking counters
and the record digest are all-reduced over NCCL.  At every N the digest of all modified records (include/ga_digest.h)
is compared with the oracle's digest of the same sessions (tests/golden/workload_digests.json, written by
tools/make_workload_digests.py), so the merged output is proven identical for N = 1, 2, 4, 8.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "session reads/s masked (germline-variant masking hot path)"
UNIT = "reads/s"


def algorithmic_bytes(read_len, n_reads, n_cigar_ops, n_sessions, n_cols_per_session, n_modified, n_indel_records, seq16_used, qual16_used):
    """Compulsory traffic of ONE pass with every array touched once (DESIGN.md "Algorithmic bytes").

    This is the conservative single-pass figure.  SURVEY.md 8(d) counts the read arrays twice (discover,
    then mask); the fused session kernel stages a session once, so the single-pass figure is the honest
    floor for this design.  Both are reported."""
    per_read = (read_len + 1) // 2 + 4 + 4 + 4 + 8          # seq4 + pos + len_flag + seq_off16 + two cigar offsets
    once = n_reads * per_read + 4 * n_cigar_ops + n_sessions * (n_cols_per_session // 2 + 40)
    out = 20 * n_modified + 16 * seq16_used + 32 * qual16_used + 16 * n_sessions
    out_reads = 16 * seq16_used + 32 * qual16_used           # modified records are re-read (sequence, qualities)
    single = once + out + out_reads
    survey = 2 * (n_reads * per_read + 4 * n_cigar_ops) + n_sessions * n_cols_per_session + out + out_reads
    return single, survey


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv:
            self._stop.clear()
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        if self.t:
            self._stop.set()
            self.t.join()
            self.t = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_sample(cfg, n_windows, device=None):
    """Host copy (ReadBatch, SessionTable, reference prefix bytes) of windows [0, n_windows) of `cfg`."""
    from genomeanonymizer_b200 import synthdev as SD
    pl = cfg.plan(0, n_windows)
    prefix = min(cfg.contig_len, (n_windows + 1) * int(pl.window_stride) + 8 * cfg.window_half)
    if device is not None:
        db, ds = SD.generate_device(cfg, device, 0, n_windows)
        ref = SD.reference_device(cfg, device, 0, prefix).cpu().numpy().tobytes()
        return db.to_host(), ds.to_host(), ref
    import ctypes as C
    import numpy as np
    from genomeanonymizer_b200 import _lib
    b, s, _ = SD.generate_host(cfg, 0, n_windows, with_reference=False)
    buf = np.zeros(prefix, np.uint8)
    p = cfg.params(0, 0)
    _lib.synth_lib().ga_synth_reference_host(C.byref(p), buf.ctypes.data, 0, prefix)
    return b, s, buf.tobytes()


def time_oracle(batch, sessions, ref, threads):
    from oracle import oracle
    t0 = time.perf_counter()
    raw, st = oracle.run(batch, sessions, ref, threads=threads, decode=False)
    dt = time.perf_counter() - t0
    if st != 0:
        raise RuntimeError(f"oracle failed with status {st}")
    return dt, raw


def run_reference_arm(args):
    """--impl reference: the CPU implementation of the path (oracle port; the Python reference needs pysam and
    cannot travel to the box) on all host threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from genomeanonymizer_b200 import synthdev as SD
    from oracle import oracle
    oracle.build()
    cfg = SD.WORKLOADS[args.workload]
    threads = oracle.n_threads()
    n_w = min(cfg.total_windows, args.sample_windows or 1500)
    batch, sessions, ref = host_sample(cfg, n_w, device=None if not _has_cuda() else 0)
    for _ in range(args.warmup):
        time_oracle(batch, sessions, ref, threads)
    t = 0.0
    reads = bases = 0
    for _ in range(args.steps):
        dt, raw = time_oracle(batch, sessions, ref, threads)
        t += dt
        reads, bases = int(raw["totals"].session_reads), int(raw["totals"].session_bases)
    v = reads * args.steps / t
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "impl": "reference",
            "config": {"workload": cfg.name, "sample": f"windows [0,{n_w}) of {cfg.total_windows}", "read_len": cfg.read_len,
                       "session_reads_per_step": reads},
            "bases_per_s": bases * args.steps / t,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{n_w} of {cfg.total_windows} windows ({reads} session reads) per step"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def bench_bam_decode(n_reads=400_000, read_len=150):
    """N4 measurement (host only, SURVEY.md 8(d) level iii input side): BGZF inflate + CRC + BAM record packing into
    the engine's SoA batch by genome_files (csrc/ga_genome_io.cpp) on all host threads.  The BAM is synthetic:
    fixed-shape records built with numpy and deflated per 64 KiB block by Python's zlib (test-side writer)."""
    import struct, tempfile, zlib
    import numpy as np
    from genomeanonymizer_b200 import genome_files as GF
    rng = np.random.default_rng(7)
    name_len, nb = 9, (read_len + 1) // 2
    rec_len = 4 + 32 + name_len + 4 + nb + read_len
    recs = np.zeros((n_reads, rec_len), np.uint8)
    pos = np.sort(rng.integers(0, 50_000_000, n_reads)).astype("<i4")
    fixed = np.zeros(n_reads, dtype=[("bs", "<u4"), ("ref", "<i4"), ("pos", "<i4"), ("lname", "u1"), ("mapq", "u1"), ("bin", "<u2"),
                                     ("ncig", "<u2"), ("flag", "<u2"), ("lseq", "<u4"), ("nref", "<i4"), ("npos", "<i4"), ("tlen", "<i4")])
    fixed["bs"], fixed["pos"], fixed["lname"], fixed["mapq"], fixed["ncig"] = rec_len - 4, pos, name_len, 60, 1
    fixed["flag"], fixed["lseq"], fixed["nref"], fixed["npos"] = 99, read_len, -1, -1
    recs[:, :36] = fixed.view(np.uint8).reshape(n_reads, 36)
    names = np.char.add("T", np.char.zfill(np.arange(n_reads).astype(str), 7)).astype("S8")
    recs[:, 36:44] = np.frombuffer(names.tobytes(), np.uint8).reshape(n_reads, 8)
    recs[:, 45:49] = np.frombuffer(struct.pack("<I", read_len << 4), np.uint8)
    codes = np.array([1, 2, 4, 8], np.uint8)[rng.integers(0, 4, (n_reads, 2 * nb))]
    recs[:, 49:49 + nb] = (codes[:, 0::2] << 4) | codes[:, 1::2]
    recs[:, 49 + nb:] = rng.integers(2, 41, (n_reads, read_len), dtype=np.uint8)
    text = "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chr22\tLN:50818468\n"
    stream = (b"BAM\1" + struct.pack("<I", len(text)) + text.encode() + struct.pack("<I", 1) + struct.pack("<I", 6) + b"chr22\0" +
              struct.pack("<I", 50818468) + recs.tobytes())
    tmp = tempfile.mkdtemp(prefix="ga_bam_")
    path = os.path.join(tmp, "synthetic.bam")
    with open(path, "wb") as fh:
        for o in range(0, len(stream), 0xff00):
            d = stream[o:o + 0xff00]
            co = zlib.compressobj(1, zlib.DEFLATED, -15)
            cd = co.compress(d) + co.flush()
            fh.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(cd) + 25) + cd +
                     struct.pack("<II", zlib.crc32(d) & 0xffffffff, len(d)))
        fh.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    size = os.path.getsize(path)
    best_open = best_pack = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        f = GF.BamFile(path)
        t1 = time.perf_counter()
        cb = GF.pack_tumor_normal(f, f, "chr22")
        t2 = time.perf_counter()
        f.close()
        best_open, best_pack = min(best_open, t1 - t0), min(best_pack, (t2 - t1) / 2)
    assert cb.batch.n_reads == 2 * n_reads and np.array_equal(cb.batch.pos[:n_reads], pos)
    os.remove(path); os.rmdir(tmp)
    return {"api": "ga_bam_open + ga_bam_pack_contig (C ABI, host threads)", "reads": n_reads, "bam_bytes": size, "inflated_bytes": len(stream),
            "threads": os.cpu_count(), "open_ms": best_open * 1e3, "pack_ms": best_pack * 1e3,
            "reads_per_s": n_reads / (best_open + best_pack), "inflated_gbs": len(stream) / (best_open + best_pack) / 1e9}


def bench_file_path(device_index, n_pairs=100000):
    """Level (iii) of SURVEY.md 8(d): tumor / normal BAM + VCF + FASTA -> the reference's FASTQ and statistics files through
    run_short_read_tumor_normal_anonymizer (C++ readers, native plan, one masking pass per contig, device FASTQ text).  The
    samples are seeded synthetic ones written by the test-side BAM / FASTA / VCF writers: one contig with all the reads,
    and the same number of reads over four contigs (the entry point prepares contig k+1 while contig k is masked and
    written); the best of four runs each is reported."""
    import shutil, tempfile
    from genomeanonymizer_b200 import synth
    from genomeanonymizer_b200.engine import Engine
    from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import run_short_read_tumor_normal_anonymizer
    from tests import helpers as H
    eng = Engine(device_index)                                           # its own engine: the sample brings its own reference

    def case_of(seed, pairs):
        contig_len = 150 * 2 * pairs // 30                               # ~30x per dataset
        return synth.make_case(seed=seed, contig_len=contig_len, n_pairs=(pairs, pairs), read_len=150,
                               somatic_positions=list(range(3000, contig_len - 3000, 4000)))

    def timed(tmp, t, n, fa, vc):
        best, res = 1e9, None
        for _ in range(4):
            for f in os.listdir(tmp):                                    # a fresh output directory every run (truncating 30 MB files is not part of the path)
                if f.endswith(".fastq") or f.endswith(".statistics.txt"):
                    os.remove(os.path.join(tmp, f))
            t0 = time.perf_counter()
            res = run_short_read_tumor_normal_anonymizer([vc], [(t, n)], fa, eng, [(os.path.join(tmp, "T.out"), os.path.join(tmp, "N.out"))], True, 0, False)
            best = min(best, time.perf_counter() - t0)
        out_bytes = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp) if f.endswith(".fastq"))
        return {"reads": res[0]["reads"], "sessions": res[0]["sessions"], "ms": best * 1e3, "reads_per_s": res[0]["reads"] / best, "fastq_bytes": out_bytes}

    tmp = tempfile.mkdtemp(prefix="ga_files_")
    tmp4 = tempfile.mkdtemp(prefix="ga_files4_")
    try:
        case = case_of(9, n_pairs)
        vcf = [[case["contig"], w["keep"]["pos"] + 1, w["keep"]["pos"] + 1, 1, "N", w["keep"]["allele"], "SNV"] for w in case["windows"]]
        one = timed(tmp, *H.write_sample_files(tmp, case, vcf))
        del case
        # four contigs, a quarter of the reads each
        cases = [case_of(11 + k, n_pairs // 4) for k in range(4)]
        names = [f"ctg{k}" for k in range(4)]
        contigs = [(names[k], len(c["reference"])) for k, c in enumerate(cases)]
        reads = [dict(r, contig=names[k], name=f"k{k}_{r['name']}") for k, c in enumerate(cases) for r in c["reads"]]
        t, n = os.path.join(tmp4, "T.bam"), os.path.join(tmp4, "N.bam")
        H.write_bam(t, contigs, [r for r in reads if r["dataset"] == 0])
        H.write_bam(n, contigs, [r for r in reads if r["dataset"] == 1])
        fa, vc = os.path.join(tmp4, "ref.fa"), os.path.join(tmp4, "somatic.vcf")
        H.write_fasta(fa, [(names[k], c["reference"]) for k, c in enumerate(cases)])
        H.write_vcf(vc, [[names[k], w["keep"]["pos"] + 1, w["keep"]["pos"] + 1, 1, "N", w["keep"]["allele"], "SNV"] for k, c in enumerate(cases) for w in c["windows"]])
        del reads
        four = timed(tmp4, t, n, fa, vc)
        out = {"api": "run_short_read_tumor_normal_anonymizer (BAM + VCF + FASTA -> FASTQ + statistics files)"}
        out.update(one)
        out.update({"host_threads": os.cpu_count(), "sample": "one contig", "four_contigs": four})
        return out
    finally:
        eng.close()
        shutil.rmtree(tmp, ignore_errors=True)
        shutil.rmtree(tmp4, ignore_errors=True)


def bench_fastq(eng, cfg, dev, n_w, peak):
    """ga_fastq_layout + ga_fastq_render over every read of the first n_w windows: masked where a session modified
    the read (lowest record index wins), as it came in otherwise.  Dense synthetic qualities, 10-character names."""
    import ctypes as C
    import torch
    from genomeanonymizer_b200 import _abi
    from genomeanonymizer_b200 import synthdev as SD
    from genomeanonymizer_b200.engine import DeviceResult
    db, ds = SD.generate_device(cfg, dev, 0, n_w)
    n = db.n_reads
    units = db.seq4_bytes // 16 // max(1, n)
    g = torch.Generator(device=dev); g.manual_seed(7)
    db.qual = torch.randint(2, 41, (db.seq4_bytes * 2 + 64,), dtype=torch.uint8, device=dev, generator=g)   # dense: record r at 32 * seq_off16[r]
    db.qual_reads, db.qual_off16 = None, None
    cap = n // 3 + 1024
    dres = DeviceResult(n_w, cap, cap * (units + 1), cap * (units + 1) // 2 + 1024, dev)
    eng.run_device(db, ds, dres)
    torch.cuda.synchronize()
    tot = eng.check_device_status(dres)
    n_mod = int(tot.n_modified)
    rec = torch.full((n,), 2 ** 31 - 1, dtype=torch.int32, device=dev)
    rec.scatter_reduce_(0, dres.mod_read[:n_mod].long(), torch.arange(n_mod, dtype=torch.int32, device=dev), reduce="amin")
    rec[rec == 2 ** 31 - 1] = -1
    ids = torch.arange(n, device=dev)
    digits = torch.stack([(ids // 10 ** k) % 10 for k in range(8, -1, -1)], 1).to(torch.uint8) + 48
    lead = torch.where(ids < db.n_tumor, 84, 78).to(torch.uint8).unsqueeze(1)                          # 'T' / 'N'
    names = torch.cat([lead, digits], 1).contiguous().flatten()
    name_off = torch.arange(n + 1, dtype=torch.int64, device=dev) * 10
    reads_idx = torch.arange(n, dtype=torch.int32, device=dev)
    t_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    items = _abi.GaFastqItems(n, reads_idx.data_ptr(), rec.data_ptr(), names.data_ptr(), name_off.data_ptr())
    R, O = db.as_struct(), dres.as_struct()
    L, st = eng._L, torch.cuda.current_stream(dev).cuda_stream
    status = torch.zeros(_abi.TOTALS_BYTES, dtype=torch.uint8, device=dev)
    eng._check(L.ga_fastq_layout(eng._h, C.byref(R), C.byref(O), n_mod, C.byref(items), t_off.data_ptr(), st))
    total = int(t_off[-1].item())
    text = torch.empty(total, dtype=torch.uint8, device=dev)
    def layout():
        eng._check(L.ga_fastq_layout(eng._h, C.byref(R), C.byref(O), n_mod, C.byref(items), t_off.data_ptr(), st))

    def render():
        eng._check(L.ga_fastq_render(eng._h, C.byref(R), C.byref(O), n_mod, C.byref(items), t_off.data_ptr(), text.data_ptr(), total,
                                     status.data_ptr(), st))

    def timed(fn, reps=5):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms_layout, ms_render = timed(layout), timed(render)
    ms = ms_layout + ms_render
    err = _abi.GaTotals.from_buffer_copy(status.cpu().numpy().tobytes()).error
    L_read = cfg.read_len
    in_bytes = n * ((L_read + 1) // 2 + L_read + 10 + 4 + 4 + 4 + 8)       # seq4 + qual + name + len_flag + seq_off16 + indices + name offsets
    bytes_all = in_bytes + total + 8 * n
    return {"api": "ga_fastq_layout + ga_fastq_render (C ABI), device resident", "reads": n, "masked_reads": int((rec >= 0).sum().item()),
            "text_bytes": total, "ms": ms, "ms_layout": ms_layout, "ms_render": ms_render, "reads_per_s": n / (ms * 1e-3), "achieved_gbs": bytes_all / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": bytes_all / (ms * 1e-3) / 1e9 / peak, "device_error": int(err),
            "sample": f"every read of the first {n_w} windows"}


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


REGION = 50_000          # sessions per rank of the weak-scaling line (= the N=1 workload's session count)
DIGESTS = os.path.join(ROOT, "tests", "golden", "workload_digests.json")


def committed_digest(name, s_begin=None, s_end=None):
    """Oracle digest of sessions [s_begin, s_end) of workload `name` (whole workload by default) from the committed
    table, or None when the table has no entry or the range does not fall on block boundaries."""
    try:
        with open(DIGESTS) as f:
            t = json.load(f)[name]
    except Exception:
        return None
    if s_begin is None:
        return [int(x) for x in t["total"]]
    acc, lo_ok, hi_ok = [0, 0, 0, 0], s_begin == s_end, s_begin == s_end
    for b in t["blocks"]:
        if b[0] >= s_begin and b[1] <= s_end:
            acc = [(a + x) & ((1 << 64) - 1) for a, x in zip(acc, b[2:6])]
            lo_ok |= b[0] == s_begin
            hi_ok |= b[1] == s_end
        elif b[0] < s_end and b[1] > s_begin:
            return None
    return acc if lo_ok and hi_ok else None


class Piece:
    """Sessions [w0, w0 + nw) of one contig, resident in HBM, with their caller-owned result buffers."""
    pass


class ShardRun:
    """This rank's share of a workload: per-contig pieces (device batch, session table, result buffers)."""

    def __init__(self, eng, dev, contigs, pieces):
        import torch
        from genomeanonymizer_b200 import synthdev as SD
        from genomeanonymizer_b200.engine import DeviceResult
        self.eng, self.dev, self.contigs, self.pieces = eng, dev, contigs, []
        done_refs = set()
        for k, w0, nw, g in pieces:
            cfg = contigs[k]
            if k not in done_refs:
                ref = SD.reference_device(cfg, dev)
                eng.upload_reference(k, ref)
                del ref
                done_refs.add(k)
            p = Piece()
            p.k, p.w0, p.nw, p.g, p.cfg = k, w0, nw, g, cfg
            p.db, p.ds = SD.generate_device(cfg, dev, w0, nw)
            p.db.contig_id = k
            pl = cfg.plan(0, 0)
            p.ids = dict(session_base=g, tumor_base=w0 * int(pl.reads_per_window[0]), normal_base=w0 * int(pl.reads_per_window[1]),
                         n_tumor=p.db.n_tumor, contig=k)
            # size the caller-owned result once: a trial run reports what the piece needs
            units = p.db.seq4_bytes // 16
            upr = max(1, units // max(1, p.db.n_reads))
            cap = p.db.n_reads // 3 + 1024
            trial = DeviceResult(nw, cap, cap * (upr + 1), cap * (upr + 1) // 2 + 1024, dev)
            eng.run_device(p.db, p.ds, trial)
            torch.cuda.synchronize()
            t0 = trial.read_totals()
            if int(t0.error) not in (0, 5):
                eng.check_device_status(trial)
            del trial
            p.dres = DeviceResult(nw, int(t0.n_modified * 1.02) + 1024, int(t0.seq16_used * 1.02) + 1024, int(t0.qual16_used * 1.02) + 1024, dev)
            self.pieces.append(p)
        torch.cuda.empty_cache()
        self.sessions = sum(p.nw for p in self.pieces)
        # several pieces (the contigs of a genome): one stream each, up to the engine's three lanes, so that the
        # persistent kernels of one contig fill the SMs the tail of another leaves idle (ga_run: a lane per stream)
        self.streams = [torch.cuda.Stream(dev) for _ in range(min(3, len(self.pieces)))] if len(self.pieces) > 1 else []
        self.ev_fork = torch.cuda.Event()
        self.ev_join = [torch.cuda.Event() for _ in self.streams]

    def step(self):
        import torch
        if not self.streams:
            for p in self.pieces:
                self.eng.run_device(p.db, p.ds, p.dres)
            return
        cur = torch.cuda.current_stream(self.dev)
        self.ev_fork.record(cur)
        for i, p in enumerate(self.pieces):
            st = self.streams[i % len(self.streams)]
            if i < len(self.streams):
                st.wait_event(self.ev_fork)
            self.eng.run_device(p.db, p.ds, p.dres, stream=st)
        for st, ev in zip(self.streams, self.ev_join):
            ev.record(st)
            cur.wait_event(ev)

    def totals(self):
        """Checks every piece's device status; sums of the totals over the pieces."""
        out = {"session_reads": 0, "session_bases": 0, "n_modified": 0, "indel_records": 0, "seq16_used": 0, "qual16_used": 0,
               "masked": [0, 0, 0]}
        for p in self.pieces:
            t = self.eng.check_device_status(p.dres)
            p.tot = t
            for f in ("session_reads", "session_bases", "n_modified", "indel_records", "seq16_used", "qual16_used"):
                out[f] += int(getattr(t, f))
            out["masked"] = [a + int(b) for a, b in zip(out["masked"], t.masked)]
        return out

    def counters(self, out):
        """Masking counters of this shard (SR.py:198-204) into the int64[8] device tensor `out`."""
        import torch
        out.zero_()
        for p in self.pieces:
            out[:3] += p.dres.totals.view(torch.int64)[6:9]          # ga_totals.masked = sum of the sessions' counters

    def digest(self):
        import torch
        acc = torch.zeros(4, dtype=torch.int64, device=self.dev)
        for p in self.pieces:
            self.eng.digest(p.dres, int(p.tot.n_modified), accumulate=acc, **p.ids)
        return acc

    def fallbacks(self):
        """(sessions that took the global-scratch fallback kernel, per-reason counts) summed over the pieces (one more run each)."""
        n, why = 0, [0] * 10
        for p in self.pieces:
            self.eng.run_device(p.db, p.ds, p.dres)
            a, b = self.eng.fallback_sessions()
            n, why = n + a, [x + y for x, y in zip(why, b)]
        return n, why

    def free(self):
        import torch
        self.pieces = []
        torch.cuda.empty_cache()


def timed_steps(fn, steps, warmup, barrier, sampler=None):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    if sampler:
        sampler.stop()
    return ev0.elapsed_time(ev1)


def u64(t):
    return [int(x) & ((1 << 64) - 1) for x in t.tolist()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="", help="default: chr1-30x-50k at N=1, 50,000-session regions of wgs-30x at N>1")
    ap.add_argument("--windows", type=int, default=0, help="debug: use only the first WINDOWS windows per rank")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--chunk-sessions", type=int, default=4096)
    ap.add_argument("--sample-windows", type=int, default=0, help="CPU baseline sample (0 = the whole N=1 workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fastq", action="store_true")
    ap.add_argument("--no-bam", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling run of the whole wgs-30x genome")
    ap.add_argument("--strong-workload", default="wgs-30x")
    ap.add_argument("--others", default="auto", help="comma list of other workloads run for 3 steps each (auto: per GPU count; none)")
    ap.add_argument("--fastq-windows", type=int, default=10000, help="windows rendered by the FASTQ measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        args.workload = args.workload or "chr1-30x-50k"
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from genomeanonymizer_b200 import sharding
    from genomeanonymizer_b200 import synthdev as SD
    from genomeanonymizer_b200.engine import Engine, HostBatch, HostResult

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the masking path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = sharding.bind_to_gpu_numa_node(local)                   # host batches local to this rank's GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allsum(t):
        if world > 1:
            dist.all_reduce(t)
        return t

    def allmax(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    eng = Engine(local)
    counters = torch.zeros(8, dtype=torch.int64, device=dev)

    # ---- the workload of `value`
    if args.workload:
        wname = args.workload
        contigs = SD.genome_contigs(wname)
        total_sessions = sum(c.total_windows for c in contigs)
        spans = sharding.shard_sessions(SD.genome_session_weights(contigs), world)
        s_lo, s_hi = spans[rank]
        scaling, span_all = "strong", (0, total_sessions)
        sharding_note = f"{wname}: genome-ordered session list cut into {world} contiguous ranges balanced by session reads"
    elif world == 1:
        wname = "chr1-30x-50k"
        contigs = SD.genome_contigs(wname)
        s_lo, s_hi = 0, contigs[0].total_windows
        scaling, span_all = "weak", (0, s_hi)
        sharding_note = "single GPU"
    else:
        wname = "wgs-30x"
        contigs = SD.genome_contigs(wname)
        s_lo, s_hi = REGION * rank, REGION * (rank + 1)
        scaling, span_all = "weak", (0, REGION * world)
        sharding_note = (f"wgs-30x: rank r masks sessions [{REGION} r, {REGION} (r+1)) of the genome-ordered session list "
                         f"(rank 0 = chr1-30x-50k); no data-path collective")
    if args.windows:
        s_hi = min(s_hi, s_lo + args.windows)
    pieces = SD.genome_pieces(contigs, s_lo, s_hi)
    run = ShardRun(eng, dev, contigs, pieces)
    cfg = run.pieces[0].cfg
    n_w = run.sessions

    def step():
        run.step()
        if world > 1:
            # the only collective of the path: masking counters (SR.py:198-204) summed over the shards
            run.counters(counters)
            dist.all_reduce(counters)

    sampler = ClockSampler(local)
    launches0 = eng.launch_count
    ms = timed_steps(step, args.steps, args.warmup, barrier, sampler)
    launches = eng.launch_count - launches0
    tot = run.totals()
    session_reads, session_bases = tot["session_reads"], tot["session_bases"]
    n_hist = min(args.steps * len(run.pieces), 32)
    kernel_ms = [x for x in eng.kernel_ms_history(n_hist) if x > 0]
    stage_ms = []
    for st in range(4):
        h = [x for x in eng.stage_ms_history(st, n_hist) if x >= 0]
        stage_ms.append(sum(h) / len(h) * len(run.pieces) if h else float("nan"))
    n_fallback, fallback_reasons = run.fallbacks()
    ms_max = float(allmax(torch.tensor([ms], dtype=torch.float64, device=dev)).item())
    work = allsum(torch.tensor([session_reads, session_bases, launches, tot["n_modified"]], dtype=torch.int64, device=dev))
    total_reads, total_bases, total_launches, total_modified = (int(x) for x in work.tolist())
    value = total_reads * args.steps / (ms_max * 1e-3)

    # ---- parity of EVERY modified record against the oracle's digest of the same sessions, all ranks together
    dig = u64(allsum(run.digest()))
    want = None if args.windows else committed_digest(wname, *span_all)
    parity_digest = {"digest": dig, "oracle_digest": want,
                     "records": "unchecked (no committed oracle digest for these sessions)" if want is None else ("ok" if dig == want else "MISMATCH"),
                     "sessions": list(span_all), "source": "tests/golden/workload_digests.json (oracle on every session)"}

    # ---- roofline of the masking pass (rank 0's shard)
    p0 = run.pieces[0]
    n_cigar = sum(int(p.db.cigar_off[-1].item()) for p in run.pieces)
    n_cols = 2 * cfg.window_half + 1 + 2 * (cfg.read_len - 1)
    single, two_pass = algorithmic_bytes(cfg.read_len, session_reads, n_cigar, n_w, n_cols, tot["n_modified"],
                                         tot["indel_records"], tot["seq16_used"], tot["qual16_used"])
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # the algorithmic bytes cover the whole masking pass (scan -> resolve -> [fallback] -> emission kernels, back to
    # back on one stream), so the duration is that of the pass; the per-kernel durations are listed beside it
    pass_ms = sum(kernel_ms) / len(kernel_ms) * len(run.pieces) if kernel_ms else float("nan")
    ach = single / (pass_ms * 1e-3) / 1e9
    scan_bytes = single - (tot["n_modified"] * 20 + 2 * 16 * tot["seq16_used"] + 2 * 32 * tot["qual16_used"])
    # DRAM bytes of one pass from the committed ncu --set full capture (profiles/), when it is of this workload
    traffic, traffic_src, tj = None, None, None
    for tf in ("r02_final_traffic.json", "r01_final_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", tf)) as f:
                tj = json.load(f)
            if tj["workload"] == cfg.name and tj["windows"] == n_w:
                traffic, traffic_src = tj["dram_bytes_per_pass"], tj["source"]
                break
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "kernel": "masking pass = scan_kernel + resolve_warp_kernel<lean> + resolve_warp_kernel<mid, four-warp teams> + resolve_kernel + emit_kernel + emit_records_kernel + emit_special_kernel (|| emit_many_kernel, session_kernel<fallback>)", "kernel_ms": pass_ms,
                "stage_ms": {"scan_kernel": stage_ms[0], "resolve_kernels": stage_ms[1], "emit_kernel": stage_ms[2],
                             "fallback_kernel_tail": stage_ms[3]},
                # the dominant kernel on its own: its algorithmic bytes are the inputs, read once
                "dominant_kernel": {"name": "scan_kernel", "ms": stage_ms[0], "algorithmic_bytes": scan_bytes,
                                    "achieved": scan_bytes / (stage_ms[0] * 1e-3) / 1e9 if stage_ms[0] > 0 else None,
                                    "frac": scan_bytes / (stage_ms[0] * 1e-3) / 1e9 / peak if stage_ms[0] > 0 else None,
                                    "traffic": (tj["per_kernel"]["scan_kernel"][0] + tj["per_kernel"]["scan_kernel"][1]) if traffic else None},
                "kernel_share_of_step": pass_ms / (ms / args.steps),
                "algorithmic_bytes_per_launch": single, "bytes_per_session_read": single / max(1, session_reads),
                "two_pass_reference_traffic": {"note": "SURVEY 8(d) counts the read arrays twice (discover, then mask); not a roofline fraction of this single-pass design",
                                               "bytes_per_launch": two_pass}}

    # ---- end to end through the host entries, every rank on its own shard: host buffers in pinned memory in, compacted
    #      records in host memory out.  `e2e` = ga_run_wire (include/ga_wire.h: ~44 bytes per read over PCIe, expanded on
    #      the device); `e2e.soa` = ga_run_host over the plain structure of arrays (~100 bytes per read).
    e2e = None
    host_pieces = None
    if not args.no_e2e:
        from genomeanonymizer_b200.engine import HostWire
        from genomeanonymizer_b200.wire import pack_wire
        host_pieces = []
        t_pack = 0.0
        for p in run.pieces:
            hb = HostBatch(p.db.to_host(), p.ds.to_host())
            t0 = time.perf_counter()
            wb = pack_wire(hb.batch)
            t_pack += time.perf_counter() - t0
            hw = HostWire(wb, hb.sessions)
            hres = HostResult(p.nw, p.dres.cap_records, p.dres.cap_seq16, p.dres.cap_qual16)
            host_pieces.append((p, hb, hres, hw))
        traffic_now = [0, 0]

        def e2e_step(wire=True):
            traffic_now[0] = traffic_now[1] = 0
            for p, hb, hres, hw in host_pieces:
                if wire:
                    eng.run_wire(hw, hres, args.chunk_sessions)
                else:
                    eng.run_host(hb, hres, args.chunk_sessions)
                a, b = eng.host_traffic()
                traffic_now[0] += a
                traffic_now[1] += b

        def timed_e2e(wire, steps):
            e2e_step(wire)                                         # warm-up: allocates the lane buffers
            barrier()
            sampler.start()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step(wire)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            sampler.stop()
            for p, hb, hres, hw in host_pieces:
                assert int(hres.totals.session_reads) == int(p.tot.session_reads) and int(hres.totals.n_modified) == int(p.tot.n_modified), "host path disagrees with device path"
            t_e = float(allmax(torch.tensor([dt], dtype=torch.float64, device=dev)).item())
            io = allsum(torch.tensor(traffic_now, dtype=torch.int64, device=dev))
            return {"value": total_reads * steps / t_e, "unit": UNIT, "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
                    "steps": steps, "ms_per_step": 1e3 * t_e / steps, "h2d_bytes_per_session_read": int(io[0].item()) / max(1, total_reads)}
        soa = timed_e2e(False, max(1, args.e2e_steps - 1))
        soa["api"] = "ga_run_host (C ABI), pinned host structure of arrays in / host records out"
        soa_digest = None
        if world == 1 and len(host_pieces) == 1:
            from oracle import oracle as _o
            soa_digest = _o.digest(host_pieces[0][2].as_struct(), int(host_pieces[0][2].totals.n_modified), records=True, **run.pieces[0].ids)
        e2e = timed_e2e(True, args.e2e_steps)
        e2e.update({"api": "ga_run_wire (C ABI, include/ga_wire.h), pinned host wire form in / host records out", "chunk_sessions": args.chunk_sessions,
                    "wire_pack_ms_outside_the_timed_region": 1e3 * t_pack, "wire_bytes_per_gpu": sum(x[3].bytes for x in host_pieces), "soa": soa})

    # ---- CPU baseline = the oracle on the whole N=1 workload, and record-by-record parity against it (rank 0, N=1);
    #      at N>1 every rank checks a 1,500-session sample of its shard against the oracle, live
    cpu = None
    parity_records = None
    if not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        threads = oracle.n_threads()
        if world == 1 and len(run.pieces) == 1:
            if host_pieces is not None:
                hb_np, hs_np = host_pieces[0][1].batch, host_pieces[0][1].sessions
            else:
                hb_np, hs_np = p0.db.to_host(), p0.ds.to_host()
            ref_host = SD.reference_device(cfg, dev).cpu().numpy().tobytes()
            if args.sample_windows:
                sb, ss, sref = host_sample(cfg, min(n_w, args.sample_windows), dev)
                dt, raw = time_oracle(sb, ss, sref, threads)
                sample = f"first {min(n_w, args.sample_windows)} of {n_w} windows"
            else:
                time_oracle(*host_sample(cfg, min(n_w, 64), dev), threads)         # pages the library in
                t0 = time.perf_counter()
                raw, st = oracle.run(hb_np, hs_np, ref_host, threads=threads, decode=False, cap_frac=0.25)
                dt = time.perf_counter() - t0
                if st != 0:
                    raise RuntimeError(f"oracle failed with status {st}")
                sample = f"all {n_w} windows of the workload"
            cpu_reads = int(raw["totals"].session_reads)
            cpu = {"value": cpu_reads / dt, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{sample} ({cpu_reads} session reads), one pass, {dt:.2f} s"}
            if not args.sample_windows:
                n_exp = int(raw["totals"].n_modified)
                edig, ekeys, ehash = oracle.digest(raw["result"], n_exp, records=True, **p0.ids)
                _, gkeys, ghash = eng.digest(p0.dres, int(p0.tot.n_modified), records=True, **p0.ids)
                bad = oracle.compare_records(gkeys.cpu().numpy(), ghash.cpu().numpy(), ekeys, ehash)
                counts_ok = np.array_equal(p0.dres.sess_counts.view(-1, 4)[:n_w].cpu().numpy().view(np.uint32), raw["counts"][:n_w])
                parity_records = {"ga_run": "ok" if bad == 0 and counts_ok else f"MISMATCH({bad} records, counters {'ok' if counts_ok else 'differ'})",
                                  "records_compared": n_exp, "sessions_compared": n_w,
                                  "how": "every modified record: (session, read) key, new length, bases and printed qualities as a 128-bit hash (include/ga_digest.h), engine vs oracle run live on the same sessions"}
                if host_pieces is not None:
                    hres = host_pieces[0][2]                       # holds the result of the last ga_run_wire call
                    _, hkeys, hhash = oracle.digest(hres.as_struct(), int(hres.totals.n_modified), records=True, **p0.ids)
                    bad_h = oracle.compare_records(hkeys, hhash, ekeys, ehash)
                    hc_ok = np.array_equal(hres.sess_counts.numpy().view(np.uint32)[:4 * n_w].reshape(-1, 4), raw["counts"][:n_w])
                    parity_records["ga_run_wire"] = "ok" if bad_h == 0 and hc_ok else f"MISMATCH({bad_h} records, counters {'ok' if hc_ok else 'differ'})"
                    if soa_digest is not None:
                        bad_s = oracle.compare_records(soa_digest[1], soa_digest[2], ekeys, ehash)
                        parity_records["ga_run_host"] = "ok" if bad_s == 0 else f"MISMATCH({bad_s} records)"
                del ekeys, ehash, gkeys, ghash
            del raw
        else:
            n_s = min(1500, p0.nw)
            sb, ss = SD.generate_device(p0.cfg, dev, p0.w0, n_s)
            hb_s, hs_s = sb.to_host(), ss.to_host()
            del sb, ss
            ref_host = SD.reference_device(p0.cfg, dev).cpu().numpy().tobytes()
            raw, st = oracle.run(hb_s, hs_s, ref_host, threads=max(1, threads // world), decode=False)
            ids = dict(p0.ids)
            ids["n_tumor"] = hb_s.n_tumor
            _, ekeys, ehash = oracle.digest(raw["result"], int(raw["totals"].n_modified), records=True, **ids)
            _, gkeys, ghash = eng.digest(p0.dres, int(p0.tot.n_modified), records=True, **p0.ids)
            # the piece's tumor reads of the sampled sessions keep their ids; normal ordinals are counted per dataset too
            sel = (gkeys[:, 0] >= p0.g) & (gkeys[:, 0] < p0.g + n_s)
            bad = oracle.compare_records(gkeys[sel].cpu().numpy(), ghash[sel].cpu().numpy(), ekeys, ehash) if st == 0 else -1
            bad_all = allsum(torch.tensor([bad if bad >= 0 else 1 << 40, int(raw["totals"].n_modified)], dtype=torch.int64, device=dev))
            parity_records = {"every_rank_sample_vs_live_oracle": "ok" if int(bad_all[0].item()) == 0 else f"MISMATCH({int(bad_all[0].item())})",
                              "records_compared": int(bad_all[1].item()), "sessions_per_rank": n_s}
            del raw
    if host_pieces is not None:
        del host_pieces
    torch.cuda.empty_cache()

    # ---- next rows of the scope table: BAM decode, file path, FASTQ rendering (rank 0, N=1)
    bam = files = fastq = None
    if rank == 0 and world == 1 and not args.no_bam:
        bam = bench_bam_decode()
        try:
            files = bench_file_path(local)
        except Exception as exc:                                          # a side measurement must not take the bench line down
            files = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_fastq:
        fastq = bench_fastq(eng, cfg, dev, min(n_w, args.fastq_windows), peak)
    main_cfg = {"workload": wname if world == 1 or args.workload else f"{wname} sessions [0, {REGION * world})", "windows_per_gpu": n_w, "read_len": cfg.read_len,
                "coverage": [cfg.cov_tumor, cfg.cov_normal], "session_reads_per_gpu": session_reads,
                "modified_records_per_gpu": tot["n_modified"], "modified_records": total_modified,
                "fallback_sessions_per_gpu": n_fallback, "fallback_reasons": fallback_reasons, "masked_snv_del_ins": tot["masked"],
                "sharding": sharding_note, "host_numa_binding": numa,
                "l2_policy": f"inputs ({sum(p.db.seq4_bytes + 20 * p.db.n_reads for p in run.pieces) / 1e9:.2f} GB per step) exceed the 126 MB L2"}
    run.free()
    del run

    # ---- strong scaling: the WHOLE wgs-30x genome cut across the ranks, and the other BASELINE workloads
    def side_run(name, steps=3):
        cs = SD.genome_contigs(name)
        spans = sharding.shard_sessions(SD.genome_session_weights(cs), world)
        r = ShardRun(eng, dev, cs, SD.genome_pieces(cs, *spans[rank]))

        def st():
            r.step()
            if world > 1:
                r.counters(counters)
                dist.all_reduce(counters)
        t_ms = timed_steps(st, steps, 3, barrier)
        tt = r.totals()
        hist = min(steps * len(r.pieces), 32)
        km = [x for x in eng.kernel_ms_history(hist) if x > 0]
        stg = []
        for s_i in range(4):
            h = [x for x in eng.stage_ms_history(s_i, hist) if x >= 0]
            stg.append(sum(h) / len(h) * len(r.pieces) if h else None)
        nfb, why = r.fallbacks()
        t_max = float(allmax(torch.tensor([t_ms], dtype=torch.float64, device=dev)).item())
        w = allsum(torch.tensor([tt["session_reads"], tt["session_bases"], tt["n_modified"], nfb, why[8], why[9]] + tt["masked"], dtype=torch.int64, device=dev)).tolist()
        d = u64(allsum(r.digest()))
        wd = committed_digest(name)
        out = {"workload": name, "contigs": len(cs), "sessions": int(sum(c.total_windows for c in cs)), "session_reads": int(w[0]),
               "modified_records": int(w[2]), "masked_snv_del_ins": [int(x) for x in w[6:9]], "steps": steps, "ms_per_step": t_max / steps,
               "value": w[0] * steps / (t_max * 1e-3), "unit": UNIT, "bases_per_s": w[1] * steps / (t_max * 1e-3),
               "sessions_per_rank": [b - a for a, b in spans], "pieces_this_rank": len(r.pieces),
               "pass_ms_rank0": sum(km) / len(km) * len(r.pieces) if km else None,
               "stage_ms_rank0": {"scan_kernel": stg[0], "resolve_kernels": stg[1], "emit_kernel": stg[2], "fallback_kernel_tail": stg[3]},
               "fallback_sessions": int(w[3]), "sessions_resolved_by_the_one_cta_kernel": int(w[4]), "sessions_resolved_by_the_mid_one_warp_kernel": int(w[5]) - int(w[4]),
               "digest": d, "oracle_digest": wd,
               "parity_records": "unchecked (no committed oracle digest)" if wd is None else ("ok" if d == wd else "MISMATCH")}
        r.free()
        return out

    strong = None
    if not args.no_strong and not args.windows and not args.workload:
        strong = side_run(args.strong_workload)
        strong["scaling"] = "strong"
        strong["parity_across_n"] = strong["parity_records"] + " (every GPU count is compared with the same oracle digest of the whole genome)"
    others = {}
    if args.others == "auto":
        names = ["cigar-stress", "cigar-stress-100m", "dense-60x30x", "noisy-60x30x", "varied-depth", "chr22-1k"] if world == 1 else ["dense-60x30x", "cigar-stress", "cigar-stress-100m"] + (["wgs-60x30x"] if world >= 4 else [])
    else:
        names = [x for x in args.others.split(",") if x and x != "none"]
    if not args.windows:
        for nm in names:
            others[nm] = side_run(nm)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "u8", "data": "synthetic", "config": main_cfg,
                "bases_per_s": total_bases * args.steps / (ms_max * 1e-3),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity_records": parity_records, "parity_digest": parity_digest,
                "strong": strong, "other_workloads": others,
                "fastq": fastq, "bam_decode": bam, "file_path": files, "gpu_launches": total_launches,
                "clocks": sampler.summary()}
        try:
            with open(os.path.join(ROOT, "tests", "golden", "ref_timing.json")) as f:
                line["cpu_baseline_reference_python"] = json.load(f)
        except Exception:
            pass
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
