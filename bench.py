#!/usr/bin/env python
"""bench.py - reads/s of the per-read germline-masking hot path on synthetic tumor-normal sessions.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one pass of the whole hot path (session assignment, allele discovery, germline resolution,
masking, compaction) over every session of the workload.  `value` times ga_run() on a batch already
resident in HBM; `e2e` times ga_run_host() - host SoA buffers in pinned memory in, compacted modified
records in host memory out, H2D/D2H inside the timed region.  `roofline` is the session kernel against
the measured HBM copy peak; `cpu_baseline` is the CPU oracle (a C restatement of the reference's
algorithm, all host threads) on a bounded sample of the same sessions.

N > 1 (torchrun): each rank owns one contig-sized shard of the genome (region sharding),
no data-path collective; the masking counters are all-reduced once per step over NCCL.  Weak scaling.
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


def bench_file_path(device_index, n_pairs=20000):
    """Level (iii) of SURVEY.md 8(d): tumor / normal BAM + VCF + FASTA -> the reference's FASTQ and statistics files through
    run_short_read_tumor_normal_anonymizer (C++ readers, native plan, one masking pass, device FASTQ text).  The sample is
    a seeded synthetic one written by the test-side BAM / FASTA / VCF writers; the best of four runs is reported."""
    import shutil, tempfile
    from genomeanonymizer_b200 import synth
    from genomeanonymizer_b200.engine import Engine
    from genomeanonymizer_b200.short_read_tumor_normal_anonymizer import run_short_read_tumor_normal_anonymizer
    from tests import helpers as H
    eng = Engine(device_index)                                           # its own engine: the sample brings its own reference
    contig_len = 150 * 2 * n_pairs // 30                                 # ~30x per dataset
    case = synth.make_case(seed=9, contig_len=contig_len, n_pairs=(n_pairs, n_pairs), read_len=150,
                           somatic_positions=list(range(3000, contig_len - 3000, 4000)))
    tmp = tempfile.mkdtemp(prefix="ga_files_")
    try:
        vcf = [[case["contig"], w["keep"]["pos"] + 1, w["keep"]["pos"] + 1, 1, "N", w["keep"]["allele"], "SNV"] for w in case["windows"]]
        t, n, fa, vc = H.write_sample_files(tmp, case, vcf)
        best = 1e9
        for _ in range(4):
            t0 = time.perf_counter()
            res = run_short_read_tumor_normal_anonymizer([vc], [(t, n)], fa, eng, [(os.path.join(tmp, "T.out"), os.path.join(tmp, "N.out"))], True, 0, False)
            best = min(best, time.perf_counter() - t0)
        out_bytes = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp) if f.endswith(".fastq"))
        return {"api": "run_short_read_tumor_normal_anonymizer (BAM + VCF + FASTA -> FASTQ + statistics files)", "reads": res[0]["reads"],
                "sessions": res[0]["sessions"], "ms": best * 1e3, "reads_per_s": res[0]["reads"] / best, "fastq_bytes": out_bytes,
                "host_threads": os.cpu_count()}
    finally:
        eng.close()
        shutil.rmtree(tmp, ignore_errors=True)


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="chr1-30x-50k")
    ap.add_argument("--windows", type=int, default=0, help="debug: use only the first WINDOWS windows per rank")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--chunk-sessions", type=int, default=4096)
    ap.add_argument("--sample-windows", type=int, default=0, help="CPU baseline sample (0 = auto, ~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fastq", action="store_true")
    ap.add_argument("--no-bam", action="store_true")
    ap.add_argument("--fastq-windows", type=int, default=10000, help="windows rendered by the FASTQ measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from genomeanonymizer_b200 import synthdev as SD
    from genomeanonymizer_b200.engine import DeviceResult, Engine, HostBatch, HostResult

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the masking path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from genomeanonymizer_b200.sharding import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local)                            # host batches local to this rank's GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    base = SD.WORKLOADS[args.workload]
    # region sharding: rank r owns contig r of the synthetic genome (same shape, its own seed)
    from dataclasses import replace
    cfg = replace(base, seed=base.seed + 7919 * rank, name=base.name)
    n_w = args.windows or cfg.total_windows
    eng = Engine(local)
    db, ds = SD.generate_device(cfg, dev, 0, n_w)
    ref_dev = SD.reference_device(cfg, dev)
    eng.upload_reference(0, ref_dev)
    del ref_dev
    torch.cuda.empty_cache()
    units = db.seq4_bytes // 16
    cap_rec = db.n_reads // 3 + 1024
    upr = max(1, units // max(1, db.n_reads))
    dres = DeviceResult(n_w, cap_rec, cap_rec * (upr + 1), cap_rec * (upr + 1) // 2 + 1024, dev)
    # size the caller-owned result once: a trial run reports what the workload needs (ga_totals holds the need
    # when the capacities are exceeded)
    eng.run_device(db, ds, dres)
    torch.cuda.synchronize()
    t0 = dres.read_totals()
    if int(t0.error) == 5:                                    # GA_ERR_CAPACITY
        cap_rec = int(t0.n_modified * 1.05) + 1024
        del dres
        torch.cuda.empty_cache()
        dres = DeviceResult(n_w, cap_rec, int(t0.seq16_used * 1.05) + 1024, int(t0.qual16_used * 1.05) + 1024, dev)
    counters = torch.zeros(8, dtype=torch.int64, device=dev)

    def step():
        eng.run_device(db, ds, dres)
        if world > 1:
            # the only collective of the path: masking counters (SR.py:198-204) summed over the shards
            counters.zero_()
            counters[:3] = dres.sess_counts.view(-1, 4)[:, :3].sum(0)
            dist.all_reduce(counters)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    tot = eng.check_device_status(dres)
    session_reads, session_bases = int(tot.session_reads), int(tot.session_bases)

    sampler = ClockSampler(local)
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    kernel_ms = [x for x in eng.kernel_ms_history(min(args.steps, 32)) if x > 0]
    stage_ms = []
    for st in range(4):
        h = [x for x in eng.stage_ms_history(st, min(args.steps, 32)) if x >= 0]
        stage_ms.append(sum(h) / len(h) if h else float("nan"))
    eng.check_device_status(dres)
    n_fallback, fallback_reasons = eng.fallback_sessions()
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    work = torch.tensor([session_reads, session_bases, launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(work)
    ms_max = float(t_ms.item())
    total_reads, total_bases, total_launches = (int(x) for x in work.tolist())
    value = total_reads * args.steps / (ms_max * 1e-3)

    # ---- roofline of the session kernel (rank 0's shard)
    n_cigar = int(db.cigar_off[-1].item())
    n_cols = 2 * cfg.window_half + 1 + 2 * (cfg.read_len - 1)
    single, survey = algorithmic_bytes(cfg.read_len, session_reads, n_cigar, n_w, n_cols, int(tot.n_modified),
                                       int(tot.indel_records), int(tot.seq16_used), int(tot.qual16_used))
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # the algorithmic bytes cover the whole masking pass (scan -> resolve -> [fallback] -> emission kernels, back to
    # back on one stream), so the duration is that of the pass; the per-kernel durations are listed beside it
    pass_ms = sum(kernel_ms) / len(kernel_ms) if kernel_ms else float("nan")
    ach = single / (pass_ms * 1e-3) / 1e9
    scan_bytes = single - (int(tot.n_modified) * 20 + 2 * 16 * int(tot.seq16_used) + 2 * 32 * int(tot.qual16_used))
    # DRAM bytes of one pass from the committed ncu --set full capture (profiles/), when it is of this workload
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_final_traffic.json")) as f:
            tj = json.load(f)
        if tj["workload"] == cfg.name and tj["windows"] == n_w:
            traffic, traffic_src = tj["dram_bytes_per_pass"], tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "kernel": "masking pass = scan_kernel + resolve_lean_kernel + resolve_kernel + emit_kernel + emit_special_kernel (|| emit_many_kernel, session_kernel<fallback>)", "kernel_ms": pass_ms,
                "stage_ms": {"scan_kernel": stage_ms[0], "resolve_kernels": stage_ms[1], "emit_kernel": stage_ms[2],
                             "fallback_kernel_tail": stage_ms[3]},
                "scan_kernel_gbs": scan_bytes / (stage_ms[0] * 1e-3) / 1e9 if stage_ms[0] > 0 else None,
                # the dominant kernel on its own: its algorithmic bytes are the inputs, read once
                "dominant_kernel": {"name": "scan_kernel", "ms": stage_ms[0], "algorithmic_bytes": scan_bytes,
                                    "achieved": scan_bytes / (stage_ms[0] * 1e-3) / 1e9 if stage_ms[0] > 0 else None,
                                    "frac": scan_bytes / (stage_ms[0] * 1e-3) / 1e9 / peak if stage_ms[0] > 0 else None,
                                    "traffic": (tj["per_kernel"]["scan_kernel"][0] + tj["per_kernel"]["scan_kernel"][1]) if traffic else None},
                "kernel_share_of_step": pass_ms / (ms / args.steps),
                "algorithmic_bytes_per_launch": single, "bytes_per_session_read": single / max(1, session_reads),
                "survey_8d_bytes_per_launch": survey, "survey_8d_achieved": survey / (pass_ms * 1e-3) / 1e9,
                "survey_8d_frac": survey / (pass_ms * 1e-3) / 1e9 / peak}

    # ---- end to end through the host entry (pinned host SoA in, host records out)
    e2e = None
    if not args.no_e2e:
        hb = HostBatch(db.to_host(), ds.to_host())
        hres = HostResult(n_w, dres.cap_records, dres.cap_seq16, dres.cap_qual16)
        eng.run_host(hb, hres, args.chunk_sessions)            # warm-up: allocates the lane buffers
        barrier()
        sampler.start()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            th = eng.run_host(hb, hres, args.chunk_sessions)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        sampler.stop()
        h2d, d2h = eng.host_traffic()
        assert int(th.session_reads) == session_reads and int(th.n_modified) == int(tot.n_modified), "host path disagrees with device path"
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        io = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
            dist.all_reduce(io)
        e2e = {"value": total_reads * args.e2e_steps / float(t_e.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
               "steps": args.e2e_steps, "ms_per_step": 1e3 * float(t_e.item()) / args.e2e_steps,
               "api": "ga_run_host (C ABI), pinned host SoA in / host records out", "chunk_sessions": args.chunk_sessions}
        del hb, hres

    # ---- next row of the scope table: FASTQ rendering of the masked reads, device resident
    bam = None
    if rank == 0 and world == 1 and not args.no_bam:
        bam = bench_bam_decode()
    files = None
    if rank == 0 and world == 1 and not args.no_bam:
        try:
            files = bench_file_path(local)
        except Exception as exc:                                          # a side measurement must not take the bench line down
            files = {"error": repr(exc)}
    fastq = None
    if rank == 0 and world == 1 and not args.no_fastq:
        fastq = bench_fastq(eng, cfg, dev, min(n_w, args.fastq_windows), peak)

    # ---- CPU baseline on a bounded sample + parity of the sampled sessions (rank 0, N=1)
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        threads = oracle.n_threads()
        n_s = args.sample_windows
        if not n_s:
            pb, ps, pref = host_sample(cfg, min(n_w, 64), dev)
            dt, _ = time_oracle(pb, ps, pref, threads)
            dt, _ = time_oracle(pb, ps, pref, threads)
            n_s = int(min(n_w, max(64, 12.0 / max(dt / min(n_w, 64), 1e-6))))
        sb, ss, sref = host_sample(cfg, n_s, dev)
        dt, raw = time_oracle(sb, ss, sref, threads)
        cpu_reads = int(raw["totals"].session_reads)
        cpu = {"value": cpu_reads / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {n_s} of {n_w} windows ({cpu_reads} session reads), one pass, {dt:.2f} s"}
        got = dres.sess_counts.view(-1, 4)[:n_s].cpu().numpy().view(np.uint32)
        parity = "ok" if np.array_equal(got, raw["counts"][:n_s]) else "MISMATCH"

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": cfg.name, "windows_per_gpu": n_w, "read_len": cfg.read_len,
                           "coverage": [cfg.cov_tumor, cfg.cov_normal], "session_reads_per_gpu": session_reads,
                           "modified_records_per_gpu": int(tot.n_modified),
                           "fallback_sessions_per_gpu": n_fallback, "fallback_reasons": fallback_reasons, "masked_snv_del_ins": [int(x) for x in tot.masked],
                           "sharding": "one contig-sized region per GPU, no data-path collective" if world > 1 else "single GPU",
                           "host_numa_binding": numa,
                           "l2_policy": f"inputs ({(db.seq4_bytes + 20 * db.n_reads) / 1e9:.2f} GB per step) exceed the 126 MB L2"},
                "bases_per_s": total_bases * args.steps / (ms_max * 1e-3),
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "fastq": fastq, "bam_decode": bam, "file_path": files, "gpu_launches": total_launches,
                "clocks": sampler.summary(), "parity_vs_oracle_on_sample": parity}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
