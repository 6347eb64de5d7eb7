"""Python host side above the C ABI: device buffers (torch tensors), engine handle, result decoding.

PyTorch is used for device memory, streams and torch.distributed only; every kernel is in
libga_b200.so (csrc/).  Nothing here computes masking on the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _abi, _lib
from .batch import MaskResult, ReadBatch, SessionTable, decode_result


def _dev(a: Optional[np.ndarray], device, pad=0):
    if a is None:
        return None
    t = torch.from_numpy(np.ascontiguousarray(a))
    if not pad:
        return t.to(device, non_blocking=False)
    out = torch.empty(t.numel() + pad, dtype=t.dtype, device=device)      # padded on the device: no host copy of the whole array
    out[:t.numel()].copy_(t)
    out[t.numel():].zero_()
    return out


class DeviceBatch:
    """ReadBatch resident in HBM (one torch tensor per array)."""

    def __init__(self, b: ReadBatch, device):
        self.n_reads, self.n_tumor = b.n_reads, b.n_tumor
        self.max_ref_span, self.contig_id = b.max_ref_span, b.contig_id
        v = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)) if a is not None else None
        self.pos = _dev(b.pos, device)
        self.len_flag = v(b.len_flag).to(device)
        self.seq_off16 = v(b.seq_off16).to(device)
        self.cigar_off = v(b.cigar_off).to(device)
        self.cigar = _dev(np.ascontiguousarray(b.cigar).view(np.int32), device, pad=4)
        self.seq4 = _dev(b.seq4, device, pad=64)
        self.qual = _dev(b.qual, device, pad=64) if b.qual is not None else None
        self.qual_reads = _dev(b.qual_reads, device) if b.qual_reads is not None else None
        self.qual_off16 = v(b.qual_off16).to(device) if b.qual_off16 is not None else None
        self.seq4_bytes = int(b.seq4.shape[0])

    def as_struct(self) -> _abi.GaReads:
        p = lambda t: t.data_ptr() if t is not None else None
        s = _abi.GaReads()
        s.n_reads, s.n_tumor = self.n_reads, self.n_tumor
        s.pos, s.len_flag, s.seq_off16 = p(self.pos), p(self.len_flag), p(self.seq_off16)
        s.cigar_off, s.cigar, s.seq4, s.qual = p(self.cigar_off), p(self.cigar), p(self.seq4), p(self.qual)
        s.seq4_bytes = self.seq4_bytes
        s.n_qual = 0 if self.qual_reads is None else int(self.qual_reads.shape[0])
        s.qual_reads, s.qual_off16 = p(self.qual_reads), p(self.qual_off16)
        s.max_ref_span, s.contig_id = int(self.max_ref_span), int(self.contig_id)
        return s


class DeviceSessions:
    def __init__(self, t: SessionTable, device):
        self.n_sessions = t.n_sessions
        f = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else np.ascontiguousarray(a)).to(device)
        self.first, self.last = f(t.first), f(t.last)
        self.keep_type, self.keep_pos, self.keep_end, self.keep_len = f(t.keep_type), f(t.keep_pos), f(t.keep_end), f(t.keep_len)
        self.keep_allele_off = f(t.keep_allele_off)
        self.keep_alleles = torch.from_numpy(np.ascontiguousarray(t.keep_alleles)).to(device)

    def as_struct(self) -> _abi.GaSessions:
        s = _abi.GaSessions()
        s.n_sessions = self.n_sessions
        s.first, s.last = self.first.data_ptr(), self.last.data_ptr()
        s.keep_type, s.keep_pos = self.keep_type.data_ptr(), self.keep_pos.data_ptr()
        s.keep_end, s.keep_len = self.keep_end.data_ptr(), self.keep_len.data_ptr()
        s.keep_allele_off, s.keep_alleles = self.keep_allele_off.data_ptr(), self.keep_alleles.data_ptr()
        return s


class DeviceResult:
    """Caller-owned output buffers in HBM."""

    def __init__(self, n_sessions, cap_records, cap_seq16, cap_qual16, device):
        i32 = dict(dtype=torch.int32, device=device)
        self.n_sessions = n_sessions
        self.cap_records, self.cap_seq16, self.cap_qual16 = int(cap_records), int(cap_seq16), int(cap_qual16)
        self.mod_session = torch.empty(self.cap_records, **i32)
        self.mod_read = torch.empty(self.cap_records, **i32)
        self.mod_len = torch.empty(self.cap_records, **i32)
        self.mod_seq_off16 = torch.empty(self.cap_records, **i32)
        self.mod_qual_off16 = torch.empty(self.cap_records, **i32)
        self.out_seq4 = torch.empty(self.cap_seq16 * 16, dtype=torch.uint8, device=device)
        self.out_qual = torch.empty(self.cap_qual16 * 32, dtype=torch.uint8, device=device)
        self.sess_counts = torch.zeros(max(1, n_sessions) * 4, **i32)
        self.totals = torch.zeros(_abi.TOTALS_BYTES, dtype=torch.uint8, device=device)

    def as_struct(self) -> _abi.GaResult:
        r = _abi.GaResult()
        r.cap_records, r.cap_seq16, r.cap_qual16 = self.cap_records, self.cap_seq16, self.cap_qual16
        r.mod_session, r.mod_read, r.mod_len = self.mod_session.data_ptr(), self.mod_read.data_ptr(), self.mod_len.data_ptr()
        r.mod_seq_off16, r.mod_qual_off16 = self.mod_seq_off16.data_ptr(), self.mod_qual_off16.data_ptr()
        r.out_seq4, r.out_qual = self.out_seq4.data_ptr(), self.out_qual.data_ptr()
        r.sess_counts, r.totals = self.sess_counts.data_ptr(), self.totals.data_ptr()
        return r

    def read_totals(self) -> _abi.GaTotals:
        raw = self.totals.cpu().numpy().tobytes()
        return _abi.GaTotals.from_buffer_copy(raw)

    def to_host(self, edits=None) -> MaskResult:
        t = self.read_totals()
        n = int(t.n_modified)
        u32 = lambda x, k: x[:k].cpu().numpy().view(np.uint32)
        return decode_result(self.n_sessions, t, self.mod_session[:n].cpu().numpy(), self.mod_read[:n].cpu().numpy(),
                             u32(self.mod_len, n), u32(self.mod_seq_off16, n), u32(self.mod_qual_off16, n),
                             self.out_seq4[:int(t.seq16_used) * 16].cpu().numpy(),
                             self.out_qual[:int(t.qual16_used) * 32].cpu().numpy(),
                             self.sess_counts[:self.n_sessions * 4].cpu().numpy().view(np.uint32), edits=edits)


class Engine:
    """Owns a ga_engine handle on one GPU."""

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("genomeanonymizer_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._L = _lib.lib()
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        st = self._L.ga_engine_create(self.device_index, C.byref(h))
        if st != _abi.GA_OK:
            raise _abi.GaError(st, self._L.ga_status_string(st).decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.ga_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != _abi.GA_OK:
            _abi.raise_for_status(st, self._L.ga_last_error(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(self._L.ga_launch_count(self._h))

    def last_kernel_ms(self) -> float:
        return float(self._L.ga_last_kernel_ms(self._h))

    def kernel_ms_history(self, n: int = 32):
        buf = (C.c_float * n)()
        k = int(self._L.ga_kernel_ms_history(self._h, buf, n))
        return [float(buf[i]) for i in range(k)]

    def fallback_sessions(self):
        """(sessions of the last run that took the fallback kernel, per-reason counts) - see ga_last_fallback_sessions."""
        buf = (C.c_int32 * 10)()
        n = int(self._L.ga_last_fallback_sessions(self._h, buf, 10))
        return n, [int(x) for x in buf]

    def stage_ms_history(self, stage: int, n: int = 32):
        """stage: 0 scan, 1 resolve, 2 emission kernel, 3 tail of the fallback kernel."""
        buf = (C.c_float * n)()
        k = int(self._L.ga_stage_ms_history(self._h, stage, buf, n))
        return [float(buf[i]) for i in range(k)]

    def upload_reference(self, contig_id: int, bases) -> None:
        """bases: str / bytes (host) or a uint8 CUDA tensor of ASCII codes."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if isinstance(bases, torch.Tensor):
            assert bases.dtype == torch.uint8 and bases.is_cuda
            self._check(self._L.ga_upload_reference(self._h, contig_id, bases.data_ptr(), bases.numel(), stream))
            torch.cuda.current_stream(self.device).synchronize()
            return
        if isinstance(bases, str):
            bases = bases.encode("ascii")
        buf = np.frombuffer(bases, dtype=np.uint8)
        self._check(self._L.ga_upload_reference(self._h, contig_id, buf.ctypes.data, len(buf), stream))

    def keep_edits(self, on: bool = True) -> None:
        """The following runs keep every record's edit description (ga_engine_keep_edits; per-sample driver, quirk Q12)."""
        self._check(self._L.ga_engine_keep_edits(self._h, 1 if on else 0))

    def record_edits(self, rec_idx) -> np.ndarray:
        """[n, 8] uint32: the edit descriptions of records `rec_idx` of the last run (ga_record_edits)."""
        idx = np.ascontiguousarray(rec_idx, np.int64)
        out = np.zeros((len(idx), 8), np.uint32)
        if len(idx):
            self._check(self._L.ga_record_edits(self._h, idx.ctypes.data_as(C.POINTER(C.c_int64)), len(idx), out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out

    def run_device(self, dbatch: DeviceBatch, dsess: DeviceSessions, dres: DeviceResult, stream=None) -> None:
        """Asynchronous launch over device-resident buffers (the `value` path of bench.py)."""
        s = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        R, S, O = dbatch.as_struct(), dsess.as_struct(), dres.as_struct()
        self._check(self._L.ga_run(self._h, C.byref(R), C.byref(S), C.byref(O), s))

    def _pinned_text(self, n_bytes: int):
        """A pinned host buffer of at least n_bytes for rendered text; it only grows (page-locking is slow)."""
        buf = getattr(self, "_text_buf", None)
        if buf is None or buf.numel() < n_bytes:
            self._text_buf = buf = torch.empty(max(n_bytes + n_bytes // 4, 1 << 20), dtype=torch.uint8, pin_memory=True)
        return buf

    def render_fastq(self, dbatch: DeviceBatch, names, reads_idx, records_idx=None, dres: "DeviceResult" = None, n_records: int = 0, as_view: bool = False):
        """FASTQ text of reads `reads_idx` (ga_fastq_layout + ga_fastq_render on the device).

        names: query name of every read of the batch (list of str).  records_idx[k] >= 0 renders read reads_idx[k]
        with the sequence / length / qualities of modified record records_idx[k] of `dres` (the reference prints the
        masked read, anonymizer_methods.py:205-243); -1 renders the read as it came in.
        Returns (text: bytes, offsets: np.ndarray[int64] of n+1 entries).  as_view: the text is a memoryview of the engine's
        pinned download buffer instead (no copy; valid until the next call) - what the file writer appends to its files."""
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream(self.device).cuda_stream
            n = len(reads_idx)
            if isinstance(names, tuple):                              # (uint8 blob, int64 offsets), e.g. from genome_files
                blob, name_off = np.ascontiguousarray(names[0], np.uint8), np.ascontiguousarray(names[1], np.int64)
                if blob.size == 0:
                    blob = np.zeros(1, np.uint8)
            else:
                enc = [nm.encode("ascii") for nm in names]
                name_off = np.zeros(len(enc) + 1, np.int64)
                np.cumsum([len(b) for b in enc], out=name_off[1:])
                blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8)
            dev = self.device
            t_names = torch.from_numpy(blob.copy()).to(dev)
            t_noff = torch.from_numpy(name_off).to(dev)
            t_read = torch.as_tensor(np.asarray(reads_idx, np.int32)).to(dev)
            rec = np.full(n, -1, np.int32) if records_idx is None else np.asarray(records_idx, np.int32)
            t_rec = torch.as_tensor(rec).to(dev)
            t_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            items = _abi.GaFastqItems(n, t_read.data_ptr(), t_rec.data_ptr(), t_names.data_ptr(), t_noff.data_ptr())
            R = dbatch.as_struct()
            O = dres.as_struct() if dres is not None else None
            pO = C.byref(O) if O is not None else None
            self._check(self._L.ga_fastq_layout(self._h, C.byref(R), pO, int(n_records), C.byref(items), t_off.data_ptr(), s))
            total = int(t_off[-1].item())
            t_text = torch.empty(max(1, total), dtype=torch.uint8, device=dev)
            status = torch.zeros(_abi.TOTALS_BYTES, dtype=torch.uint8, device=dev)
            self._check(self._L.ga_fastq_render(self._h, C.byref(R), pO, int(n_records), C.byref(items), t_off.data_ptr(),
                                                t_text.data_ptr(), total, status.data_ptr(), s))
            torch.cuda.synchronize(dev)
            t = _abi.GaTotals.from_buffer_copy(status.cpu().numpy().tobytes())
            if t.error:
                _abi.raise_for_status(int(t.error), f"ga_fastq_render failed at item {t.error_detail}")
            if as_view:
                host = self._pinned_text(total)
                host[:total].copy_(t_text[:total])                    # synchronous for the host (pinned target, blocking copy)
                return memoryview(host.numpy())[:total], t_off.cpu().numpy()
            return t_text[:total].cpu().numpy().tobytes(), t_off.cpu().numpy()

    def check_device_status(self, dres: DeviceResult) -> _abi.GaTotals:
        t = dres.read_totals()
        if t.error:
            _abi.raise_for_status(int(t.error), f"device raised status {t.error} at index {t.error_detail} "
                                                f"(needs records={t.n_modified} seq16={t.seq16_used} qual16={t.qual16_used})")
        return t

    def run(self, batch: ReadBatch, sessions: SessionTable, cap_frac: float = 1.0, edits: bool = False) -> MaskResult:
        """Synchronous convenience: upload, run, download, decode.  edits: every indel-masked record also carries the
        description of what was applied to it (ga_record_edits; the plugin hands it to the reference's driver, quirk Q12)."""
        with torch.cuda.device(self.device):
            db, ds = DeviceBatch(batch, self.device), DeviceSessions(sessions, self.device)
            units = batch.seq4.shape[0] // 16
            dres = DeviceResult(sessions.n_sessions, max(16, int(2 * batch.n_reads * cap_frac) + 16),
                                int(2 * units * cap_frac) + 64, int(2 * units * cap_frac) + 64, self.device)
            self.keep_edits(edits)
            self.run_device(db, ds, dres)
            torch.cuda.synchronize(self.device)
            n = int(self.check_device_status(dres).n_modified)
            kept = self.record_edits(np.arange(n)) if edits else None
            if edits:
                self.keep_edits(False)                                # only this run pays for keeping them
            return dres.to_host(kept)


def _to_host_batch(db: DeviceBatch) -> ReadBatch:
    """Copy a device-resident batch back to host arrays (tests, CPU baseline sampling)."""
    u32 = lambda t: t.cpu().numpy().view(np.uint32)
    n_cigar = int(db.cigar_off[-1].item())
    return ReadBatch(n_tumor=db.n_tumor, pos=db.pos.cpu().numpy(), len_flag=u32(db.len_flag), seq_off16=u32(db.seq_off16),
                     cigar_off=u32(db.cigar_off), cigar=u32(db.cigar[:n_cigar]), seq4=db.seq4[:db.seq4_bytes].cpu().numpy(),
                     qual=None if db.qual is None else db.qual.cpu().numpy(), max_ref_span=db.max_ref_span,
                     contig_id=db.contig_id,
                     qual_reads=None if db.qual_reads is None else db.qual_reads.cpu().numpy(),
                     qual_off16=None if db.qual_off16 is None else u32(db.qual_off16))


def _to_host_sessions(ds: DeviceSessions) -> SessionTable:
    n = ds.n_sessions
    c = lambda t, k=n: t[:k].cpu().numpy()
    return SessionTable(c(ds.first), c(ds.last), c(ds.keep_type), c(ds.keep_pos), c(ds.keep_end), c(ds.keep_len),
                        c(ds.keep_allele_off, n + 1).view(np.uint32), ds.keep_alleles.cpu().numpy())


DeviceBatch.to_host = _to_host_batch
DeviceSessions.to_host = _to_host_sessions


class HostBatch:
    """A ReadBatch + SessionTable in pinned host memory: what the reference-facing plugin hands to ga_run_host."""

    def __init__(self, batch: ReadBatch, sessions: SessionTable, pin: bool = True):
        def pin_arr(a):
            if a is None:
                return None
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
            if pin and torch.cuda.is_available():
                t = t.pin_memory()
            return t
        self.batch, self.sessions = batch, sessions
        b, s = batch, sessions
        self._t = {k: pin_arr(getattr(b, k)) for k in ("pos", "len_flag", "seq_off16", "cigar_off", "cigar", "seq4", "qual", "qual_reads", "qual_off16")}
        self._s = {k: pin_arr(getattr(s, k)) for k in ("first", "last", "keep_type", "keep_pos", "keep_end", "keep_len", "keep_allele_off", "keep_alleles")}
        # CIGAR / seq4 slices are copied with their natural sizes; keep 64 spare bytes mapped behind seq4
        self.bytes = sum(t.numel() for t in self._t.values() if t is not None) + sum(t.numel() for t in self._s.values())

    def reads_struct(self) -> _abi.GaReads:
        p = lambda k: self._t[k].data_ptr() if self._t[k] is not None else None
        b = self.batch
        r = _abi.GaReads()
        r.n_reads, r.n_tumor = b.n_reads, b.n_tumor
        r.pos, r.len_flag, r.seq_off16, r.cigar_off, r.cigar = p("pos"), p("len_flag"), p("seq_off16"), p("cigar_off"), p("cigar")
        r.seq4, r.qual, r.seq4_bytes = p("seq4"), p("qual"), int(b.seq4.shape[0])
        r.n_qual = 0 if b.qual_reads is None else int(b.qual_reads.shape[0])
        r.qual_reads, r.qual_off16 = p("qual_reads"), p("qual_off16")
        r.max_ref_span, r.contig_id = int(b.max_ref_span), int(b.contig_id)
        return r

    def sessions_struct(self) -> _abi.GaSessions:
        p = lambda k: self._s[k].data_ptr()
        s = _abi.GaSessions()
        s.n_sessions = self.sessions.n_sessions
        s.first, s.last, s.keep_type, s.keep_pos = p("first"), p("last"), p("keep_type"), p("keep_pos")
        s.keep_end, s.keep_len, s.keep_allele_off, s.keep_alleles = p("keep_end"), p("keep_len"), p("keep_allele_off"), p("keep_alleles")
        return s


class HostResult:
    """Caller-owned host output buffers (pinned) for ga_run_host."""

    def __init__(self, n_sessions, cap_records, cap_seq16, cap_qual16, pin: bool = True):
        def buf(nbytes):
            t = torch.zeros(max(16, int(nbytes)), dtype=torch.uint8)
            return t.pin_memory() if pin and torch.cuda.is_available() else t
        self.n_sessions = n_sessions
        self.cap_records, self.cap_seq16, self.cap_qual16 = int(cap_records), int(cap_seq16), int(cap_qual16)
        self.mod_session, self.mod_read, self.mod_len = buf(4 * cap_records), buf(4 * cap_records), buf(4 * cap_records)
        self.mod_seq_off16, self.mod_qual_off16 = buf(4 * cap_records), buf(4 * cap_records)
        self.out_seq4, self.out_qual = buf(16 * cap_seq16), buf(32 * cap_qual16)
        self.sess_counts = buf(16 * max(1, n_sessions))
        self.totals = _abi.GaTotals()

    def as_struct(self) -> _abi.GaResult:
        r = _abi.GaResult()
        r.cap_records, r.cap_seq16, r.cap_qual16 = self.cap_records, self.cap_seq16, self.cap_qual16
        r.mod_session, r.mod_read, r.mod_len = self.mod_session.data_ptr(), self.mod_read.data_ptr(), self.mod_len.data_ptr()
        r.mod_seq_off16, r.mod_qual_off16 = self.mod_seq_off16.data_ptr(), self.mod_qual_off16.data_ptr()
        r.out_seq4, r.out_qual, r.sess_counts = self.out_seq4.data_ptr(), self.out_qual.data_ptr(), self.sess_counts.data_ptr()
        r.totals = C.addressof(self.totals)
        return r

    def decode(self) -> MaskResult:
        t = self.totals
        n = int(t.n_modified)
        v = lambda x, dt, k: x.numpy().view(dt)[:k]
        return decode_result(self.n_sessions, t, v(self.mod_session, np.int32, n), v(self.mod_read, np.int32, n),
                             v(self.mod_len, np.uint32, n), v(self.mod_seq_off16, np.uint32, n), v(self.mod_qual_off16, np.uint32, n),
                             self.out_seq4.numpy()[:int(t.seq16_used) * 16], self.out_qual.numpy()[:int(t.qual16_used) * 32],
                             v(self.sess_counts, np.uint32, self.n_sessions * 4))


class HostWire:
    """A WireBatch (include/ga_wire.h) + SessionTable in pinned host memory: what the plugin hands to ga_run_wire."""

    def __init__(self, wire, sessions: SessionTable, pin: bool = True):
        def pin_arr(a):
            if a is None:
                return None
            t = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
            if pin and torch.cuda.is_available():
                t = t.pin_memory()
            return t
        self.wire, self.sessions = wire, sessions
        self._t = {k: pin_arr(getattr(wire, k)) for k in ("blob", "dir", "qual", "qual_reads", "qual_off16")}
        if self._t["blob"] is not None and self._t["blob"].data_ptr() % 16:
            raise ValueError("the wire blob must be 16-byte aligned")
        s = sessions
        self._s = {k: pin_arr(getattr(s, k)) for k in ("first", "last", "keep_type", "keep_pos", "keep_end", "keep_len", "keep_allele_off", "keep_alleles")}
        self.bytes = sum(t.numel() for t in self._t.values() if t is not None) + sum(t.numel() for t in self._s.values())

    def reads_struct(self) -> _abi.GaReadsWire:
        p = lambda k: self._t[k].data_ptr() if self._t[k] is not None and self._t[k].numel() else None
        w = self.wire
        r = _abi.GaReadsWire()
        r.n_reads, r.n_tumor, r.n_blocks, r.n_tumor_blocks = w.n_reads, w.n_tumor, w.n_blocks, w.n_tumor_blocks
        r.blob, r.blob_bytes, r.dir = p("blob"), int(w.blob.nbytes), p("dir")
        r.qual, r.qual_reads, r.qual_off16 = p("qual"), p("qual_reads"), p("qual_off16")
        r.n_qual = 0 if w.qual_reads is None else int(len(w.qual_reads))
        r.qual_units = int(w.qual_units)
        r.max_ref_span, r.contig_id = int(w.max_ref_span), int(w.contig_id)
        return r

    sessions_struct = HostBatch.sessions_struct


def _run_wire(self, hw: HostWire, hres: HostResult, chunk_sessions: int = 0) -> _abi.GaTotals:
    """Wire form in, host records out (ga_run_wire): the end-to-end call with ~44 instead of ~100 bytes per read on PCIe."""
    R, S, O = hw.reads_struct(), hw.sessions_struct(), hres.as_struct()
    st = self._L.ga_run_wire(self._h, C.byref(R), C.byref(S), C.byref(O), int(chunk_sessions))
    if st != _abi.GA_OK:
        t = hres.totals
        _abi.raise_for_status(st, f"{self._L.ga_last_error(self._h).decode()} (device status {t.error} at {t.error_detail}; "
                                  f"needs records={t.n_modified} seq16={t.seq16_used} qual16={t.qual16_used})")
    return hres.totals


def _run_host(self, hb: HostBatch, hres: HostResult, chunk_sessions: int = 0) -> _abi.GaTotals:
    """Host buffers in, host buffers out: chunked H2D / kernels / D2H inside the C library (ga_run_host)."""
    R, S, O = hb.reads_struct(), hb.sessions_struct(), hres.as_struct()
    st = self._L.ga_run_host(self._h, C.byref(R), C.byref(S), C.byref(O), int(chunk_sessions))
    if st != _abi.GA_OK:
        t = hres.totals
        _abi.raise_for_status(st, f"{self._L.ga_last_error(self._h).decode()} (device status {t.error} at {t.error_detail}; "
                                  f"needs records={t.n_modified} seq16={t.seq16_used} qual16={t.qual16_used})")
    return hres.totals


def _host_traffic(self):
    a, b = C.c_int64(0), C.c_int64(0)
    self._L.ga_last_host_traffic(self._h, C.byref(a), C.byref(b))
    return int(a.value), int(b.value)


Engine.run_host = _run_host
Engine.run_wire = _run_wire
Engine.host_traffic = _host_traffic


def _digest(self, dres: DeviceResult, n_records: int, session_base: int = 0, tumor_base: int = 0, normal_base: int = 0,
            n_tumor: int = 0, contig: int = 0, records: bool = False, accumulate: "torch.Tensor" = None):
    """ga_result_digest over a device-resident result (include/ga_digest.h).  Returns the int64[4] digest tensor
    {sum lo, sum hi, records, sum of new lengths} (accumulated into `accumulate` when given) and, with records=True,
    the per-record (keys[n,2], hashes[n,2]) int64 tensors."""
    with torch.cuda.device(self.device):
        st = torch.cuda.current_stream(self.device).cuda_stream
        dig = accumulate if accumulate is not None else torch.zeros(4, dtype=torch.int64, device=self.device)
        keys = hashes = None
        if records:
            keys = torch.empty((max(1, n_records), 2), dtype=torch.int64, device=self.device)
            hashes = torch.empty((max(1, n_records), 2), dtype=torch.int64, device=self.device)
        ids = _abi.GaDigestIds(int(session_base), int(tumor_base), int(normal_base), int(n_tumor), int(contig))
        O = dres.as_struct()
        self._check(self._L.ga_result_digest(self._h, C.byref(O), int(n_records), C.byref(ids),
                                             keys.data_ptr() if records else None, hashes.data_ptr() if records else None,
                                             dig.data_ptr(), st))
        if records:
            return dig, keys[:n_records], hashes[:n_records]
        return dig


Engine.digest = _digest
