"""Batched sample driver: the reference's per-sample orchestration, planned on the host, executed on the engine.

SURVEY.md 8(f) N2 + N3.  The reference walks a sample section by section (`anonymize_genome`,
short_read_tumor_normal_anonymizer.py:625-760): variant windows become sessions (`anonymize_window`, :279-372), the
regions between them are fetched and passed through unless a tumor and a normal read island overlap, which makes one
more session (`anonymize_inter_window_region`, :498-558 over `pileup_io.iter_fetch_pair`, pileup_io.pyx:124-298).
Mates meet through `to_pair_anonymized_reads`, and a pair is written once (`written_read_ids`, :134-165) - by the
first section that completes it, so a read that was fetched and paired in the region before a window leaves
unmasked even though the window's session masks it.

None of this needs the bases: `plan_sample` works on (name, flag, dataset, start, end) alone and returns
  * the session table (variant windows and island sessions, in processing order), and
  * the write plan: for every output record which read it is and which session's result it shows (-1: as fetched).
The engine then masks all sessions in one `ga_run`, and `ga_fastq_render` prints the planned records.
`tests/test_genome_files.py` checks the four files byte for byte against what the reference itself wrote.

Scope: one contig per call, primary alignments and placed-unmapped mates (flag 0x4 at the mate's position: pileups never
see them, regions collect them for their end, what is left is paired after the last section); no supplementary /
secondary records; windows as `get_windows` makes them for SNVs and short indels.
"""
from __future__ import annotations

import bisect
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

WINDOW_HALF = 1000                      # window_size // 2, short_read_tumor_normal_anonymizer.py:71


def window_of_variant(pos: int, end: int) -> Tuple[int, int]:
    """(first, last) of the window around a VCF record with 1-based pos / end
    (short_read_tumor_normal_anonymizer.py:112-128: SNVs and indels shorter than the window)."""
    return pos - WINDOW_HALF, end + WINDOW_HALF + 1


def genome_sections(windows: Sequence[dict], contig_len: int) -> List[Tuple[int, int, Optional[int]]]:
    """(first, last, window index or None) in processing order for one contig
    (get_genome_sections, short_read_tumor_normal_anonymizer.py:245-276)."""
    if not windows:
        return [(0, 0, None)]                                     # the whole contig is one inter-window region
    out = []
    nxt = 1
    for k, w in enumerate(windows):
        out.append((nxt, w["first"] - 1, None))
        out.append((w["first"], w["last"], k))
        nxt = w["last"] + 1
    out.append((nxt, contig_len - 1, None))
    out.sort(key=lambda t: (t[0], t[1]))                          # sort_window_list, :67-68 (stable: a window keeps its place among ties)
    return out


@dataclass
class Plan:
    sessions: List[dict] = field(default_factory=list)            # {"first", "last", "keep", "window": index or None}
    pairs: List[Tuple[int, int, int, int, int]] = field(default_factory=list)   # (dataset, read1, version1, read2, version2) in write order
    singles: List[Tuple[int, int, int]] = field(default_factory=list)           # (dataset, read, version)


def _session_yield_order(reads, t_idx: List[int], n_idx: List[int]):
    """Pairs of one session in the order `CompleteGermlineAnonymizer.anonymize` yields them
    (anonymizer_methods.py:472-476, 489-512, 521-532): a pair whose two mates are in the session leaves at the first
    normal pileup column right of both mates, pairs in first-appearance order; everything else at the end, in
    registry order.  Reads appear at their own start (truncate=False, pileup_io.pyx:12-17), tumor column before
    normal column, file order inside a column.  Returns [(name, read index of mate 1 or None, of mate 2 or None)]."""
    order: Dict[str, int] = {}
    slots: Dict[str, list] = {}
    max_end: Dict[str, int] = {}
    seq = sorted([(reads[i]["pos"], 0, i) for i in t_idx] + [(reads[i]["pos"], 1, i) for i in n_idx])
    for _, _, i in seq:
        r = reads[i]
        order.setdefault(r["name"], len(order))
        slot = slots.setdefault(r["name"], [None, None])
        m = 0 if r["flag"] & 0x40 else 1
        if slot[m] is None:                                       # the first alignment of a (name, mate) is the read
            slot[m] = i
        max_end[r["name"]] = max(max_end.get(r["name"], -1), r["end"])
    # normal pileup columns = positions covered by a normal read of the session
    cover = sorted((reads[i]["pos"], reads[i]["end"]) for i in n_idx)
    merged: List[List[int]] = []
    for a, b in cover:
        if merged and a <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], b)
        else:
            merged.append([a, b])
    starts = [m[0] for m in merged]

    def first_normal_column_after(p: int) -> Optional[int]:      # smallest covered position > p
        k = bisect.bisect_right(starts, p + 1) - 1
        if k >= 0 and merged[k][1] > p + 1:
            return p + 1
        k += 1
        return merged[k][0] if k < len(merged) else None

    early, late = [], []
    for name, o in order.items():
        s = slots[name]
        col = first_normal_column_after(max_end[name]) if (s[0] is not None and s[1] is not None) else None
        if col is not None:
            early.append((col, o, name))
        else:
            late.append((o, name))
    return [(n, slots[n][0], slots[n][1]) for _, _, n in sorted(early)] + [(n, slots[n][0], slots[n][1]) for _, n in sorted(late)]


def _is_unmapped(r) -> bool:
    return bool(r["flag"] & 0x4)


def _islands(reads, idx: List[int], unmapped: Optional[List[int]] = None) -> List[List[int]]:
    """Chains of reads in which every read overlaps (or touches, or ends with) the one before it
    (collect_intersecting_reads, pileup_io.pyx:78-106 with compare, :44-59).  Unmapped reads behind the first fetched
    read are set aside in `unmapped` (:93-96); an unmapped FIRST read seeds an island like any read, as a read of
    length zero (compare_read_alignments_intersection uses its start as its end, :72-73)."""
    out: List[List[int]] = []
    for i in idx:
        r = reads[i]
        if out:
            if _is_unmapped(r):
                if unmapped is not None:
                    unmapped.append(i)
                continue
            last = reads[out[-1][-1]]
            last_end = last["pos"] if _is_unmapped(last) else last["end"]
            if (r["pos"] <= last_end and r["end"] >= last["pos"]) or r["end"] == last_end:
                out[-1].append(i)
                continue
        out.append([i])
    return out


def _island_right(reads, a: List[int]) -> int:
    """get_righmost_pos (pileup_io.pyx:109-121): the largest end of the island's mapped reads, 0 when there is none."""
    return max((reads[i]["end"] for i in a if not _is_unmapped(reads[i])), default=0)


def _cmp_islands(reads, a: List[int], b: List[int]) -> int:
    """compare() of pileup_io.pyx:44-59 on (first read's start, rightmost end) of two islands of one contig."""
    f1, l1 = reads[a[0]]["pos"], _island_right(reads, a)
    f2, l2 = reads[b[0]]["pos"], _island_right(reads, b)
    overlap = f2 <= l1 and l2 >= f1
    if l1 < l2:
        return -1 if overlap else -2
    if l2 < l1:
        return 1 if overlap else 2
    return -1 if f1 < f2 else (1 if f2 < f1 else 0)


def _fetch_pair_events(reads, t_idx: List[int], n_idx: List[int]):
    """What iter_fetch_pair (pileup_io.pyx:124-298) yields for the fetched tumor / normal reads of one region:
    ("single", dataset, [reads]) or ("both", (left, right)), and last ("unmapped", dataset, [reads]) for the unmapped
    reads it set aside (:298).  The last island of each dataset is always yielded singly, and when one dataset runs out
    the other's remaining islands are yielded singly too."""
    if not t_idx and not n_idx:
        return
    um: List[List[int]] = [[], []]
    ti, ni = _islands(reads, t_idx, um[0]), _islands(reads, n_idx, um[1])
    yield from _fetch_pair_island_events(reads, ti, ni)
    yield ("unmapped", 0, um[0])
    yield ("unmapped", 1, um[1])


REAPPLY = 1 << 30          # flag on the version of a planned read: its left-over indels are applied twice (quirk Q12)


def version_of(v: int) -> int:
    """The session whose masking a planned read prints (-1: as it came in), without the REAPPLY flag."""
    v = int(v)
    return v if v < 0 else v & ~REAPPLY


def reapply_pairs(plan) -> List[Tuple[int, int]]:
    """(session, read) of every planned read that carries the REAPPLY flag."""
    out = []
    for row in plan.pairs:
        out += [(version_of(v), int(r)) for r, v in ((row[1], row[2]), (row[3], row[4])) if int(v) >= 0 and int(v) & REAPPLY]
    for row in plan.singles:
        if int(row[2]) >= 0 and int(row[2]) & REAPPLY:
            out.append((version_of(row[2]), int(row[1])))
    return out


def _fetch_pair_island_events(reads, ti, ni):
    a = b = 0                                                     # current island of each dataset
    # "r is not None" in the reference = there is another island after the current one
    while True:
        more_t, more_n = a + 1 < len(ti), b + 1 < len(ni)
        if not more_t and not more_n:
            yield ("single", 0, ti[a] if a < len(ti) else [])
            yield ("single", 1, ni[b] if b < len(ni) else [])
            return
        if more_t and more_n:
            c = _cmp_islands(reads, ti[a], ni[b])
            if c < -1:
                yield ("single", 0, ti[a]); a += 1
            elif c > 1:
                yield ("single", 1, ni[b]); b += 1
            else:
                left = min(reads[ti[a][0]]["pos"], reads[ni[b][0]]["pos"])
                right = max(_island_right(reads, ti[a]), _island_right(reads, ni[b]))
                yield ("both", (left, right)); a += 1; b += 1
        else:
            if more_t:
                yield ("single", 0, ti[a]); a += 1
            if more_n:
                yield ("single", 1, ni[b]); b += 1


def plan_sample(reads: Sequence[dict], windows: Sequence[dict], contig_len: int) -> Plan:
    """reads: every read of the sample on this contig, each {"name", "flag", "dataset", "pos", "end"} with `end` the
    exclusive reference end, tumor and normal each in file (coordinate) order; windows: [{"first", "last", "keep"}]
    sorted as the reference sorts them."""
    import gc
    # The plan allocates a few small containers per read and keeps them alive; the cyclic collector then re-walks the
    # whole (acyclic) heap again and again - five times the cost of the planning itself.  Off for the duration.
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        return _plan_sample(reads, windows, contig_len)
    finally:
        if was_enabled:
            gc.enable()


def _plan_sample(reads: Sequence[dict], windows: Sequence[dict], contig_len: int) -> Plan:
    plan = Plan()
    t_all = [i for i, r in enumerate(reads) if r["dataset"] == 0]
    n_all = [i for i, r in enumerate(reads) if r["dataset"] == 1]
    to_pair: Dict[str, list] = {}                                 # name -> [(read, version) or None] * 2
    written = set()

    # AlignmentFile.fetch / pileup read selection: reads with pos < stop and end > start.  Each dataset is in coordinate
    # order, so the candidates are a bisected slice of it (pos in [start - longest span, stop)), not the whole list.
    pos_of = {id(t_all): [reads[i]["pos"] for i in t_all], id(n_all): [reads[i]["pos"] for i in n_all]}
    span = max((r["end"] - r["pos"] for r in reads), default=0)
    for idx in (t_all, n_all):
        if any(a > b for a, b in zip(pos_of[id(idx)], pos_of[id(idx)][1:])):
            raise ValueError("reads of a dataset must be in coordinate order")

    def overlapping(idx, start, stop, fetch=False):
        """pileup(): mapped reads with pos < stop and end > start (htslib drops unmapped reads); fetch(): unmapped reads
        too, as intervals of length one at their position."""
        ps = pos_of[id(idx)]
        lo, hi = bisect.bisect_left(ps, start - max(span, 1)), bisect.bisect_left(ps, stop)
        if fetch:
            return [i for i in idx[lo:hi] if (reads[i]["pos"] + 1 if _is_unmapped(reads[i]) else reads[i]["end"]) > start]
        return [i for i in idx[lo:hi] if reads[i]["end"] > start and not _is_unmapped(reads[i])]

    def write_pair(name, s1, s2, sink=None):                      # write_pair, :134-165
        if name in written:
            return
        written.add(name)
        (plan.pairs if sink is None else sink).append((reads[s1[0]]["dataset"], s1[0], s1[1], s2[0], s2[1]))

    def store(name, mate, value):                                 # add_*_to_collection: an occupied slot keeps its read
        slot = to_pair.setdefault(name, [None, None])
        if slot[mate] is None:
            slot[mate] = value
        return slot

    def run_session(first, last, keep, window_index):
        s = len(plan.sessions)
        plan.sessions.append({"first": first, "last": last, "keep": keep, "window": window_index})
        for name, m1, m2 in _session_yield_order(reads, overlapping(t_all, first, last), overlapping(n_all, first, last)):
            if m1 is not None and m2 is not None:                 # writeable as it comes (:310-312)
                write_pair(name, (m1, s), (m2, s))
                continue
            slot = None
            for mate, m in ((0, m1), (1, m2)):                    # :320-333
                if m is not None:
                    slot = store(name, mate, (m, s))
                    if slot[mate] != (m, s) and slot[mate][1] >= 0:
                        # quirk Q12: the read waits unpaired with the masking of an earlier session and is met again:
                        # update_anonymized_read_from_other (anonymizer_methods.py:281-287) switches its left-over indels on
                        # again, and whoever writes it applies them a second time
                        slot[mate] = (slot[mate][0], slot[mate][1] | REAPPLY)
            if slot[0] is not None and slot[1] is not None:       # :348-359
                write_pair(name, slot[0], slot[1])
                to_pair.pop(name)

    for first, last, k in genome_sections(windows, contig_len):
        if k is not None:
            run_session(first, last, windows[k].get("keep"), k)
            continue
        if first + last == 0:
            start, stop = 0, contig_len
        else:
            if first < 0 or first > last:                         # pysam rejects these coordinates
                raise ValueError(f"inter-window region ({first}, {last}) is not fetchable: variants closer than a window")
            start, stop = first, last
        # The region keeps its four output streams open from its first to its last read (:516-518, :558) while every
        # island session opens, writes and closes its own (:297-299, :366): the region's own records sit in the stream
        # buffers and reach the files when the region ends, behind the records of its island sessions.  (Exact while a
        # region's pass-through text per file stays below the platform's stream buffer; see DESIGN.md.)
        deferred: List[Tuple[int, int, int, int, int]] = []
        for ev in _fetch_pair_events(reads, overlapping(t_all, start, stop, fetch=True), overlapping(n_all, start, stop, fetch=True)):
            if ev[0] == "both":
                run_session(ev[1][0], ev[1][1], None, None)       # an island session has no variant to keep (:523-534)
                continue
            for i in ev[2]:                                       # pair_unmapped_or_non_pileup_pairs_and_write, :375-406
                r = reads[i]
                slot = store(r["name"], 0 if r["flag"] & 0x40 else 1, (i, -1))
                if slot[0] is not None and slot[1] is not None:
                    write_pair(r["name"], slot[0], slot[1], deferred)   # (stays in the collection until the end, :737-741)
        plan.pairs.extend(deferred)
    # pair_unmapped_mates (:561-600, called at :725-732 when the collection is not empty): every window is fetched again
    # (start = first - 1), tumor then normal, and an unmapped read whose name waits in the collection joins it
    if to_pair:
        for w in windows:
            for idx in (t_all, n_all):
                for i in overlapping(idx, max(w["first"] - 1, 0), w["last"], fetch=True):
                    r = reads[i]
                    if _is_unmapped(r) and r["name"] in to_pair:
                        slot = store(r["name"], 0 if r["flag"] & 0x40 else 1, (i, -1))
                        if slot[0] is not None and slot[1] is not None:
                            write_pair(r["name"], slot[0], slot[1])
    for name in written:
        to_pair.pop(name, None)
    for name, slot in to_pair.items():                            # write_single_end_reads, :603-622
        i, v = slot[0] if slot[0] is not None else slot[1]
        plan.singles.append((reads[i]["dataset"], i, v))
    return plan


def statistics_text(contig, plan: Optional[Plan] = None, sess_counts=None) -> str:
    """The `<normal_bam>.statistics.txt` file (AnonymizedVariantsStatistics.write_statistics,
    short_read_tumor_normal_anonymizer.py:212-242): one row per variant window with the masked variants by type
    (SNV, DEL, INS and the five structural types this path never calls), island sessions summed into the
    `outside_windows` row, then total / mean / median / max / min over ALL rows.
    statistics_text(contig, plan, counts) for one contig, or statistics_text([(contig, plan, counts), ...]) for a
    sample of several contigs in genome order."""
    import numpy as np
    parts = [(contig, plan, sess_counts)] if plan is not None else list(contig)
    rows = [("outside_windows", "-", "-", [0] * 8)]
    for name, pl, counts8 in parts:
        for s, ses in enumerate(pl.sessions):
            c = [int(x) for x in counts8[s][:3]] + [0] * 5
            if ses["window"] is None:
                rows[0] = rows[0][:3] + ([a + b for a, b in zip(rows[0][3], c)],)
            else:
                rows.append((name, str(ses["first"]), str(ses["last"]), c))
    out = ["\t".join(["#SEQ", "#FIRST", "#LAST", "#SNV", "#DEL", "#INS", "#DUP", "#INV", "#CNV", "#TRA", "#SGL"])]
    out += ["\t".join([a, b, c] + [str(x) for x in counts]) for a, b, c, counts in rows]
    out.append("### Overall statistics:")
    out.append("\t".join(["#SNV", "#DEL", "#INS", "#DUP", "#INV", "#CNV", "#TRA", "#SGL"]))
    cols = [np.array([r[3][k] for r in rows], dtype=np.int64) for k in range(8)]
    for stat, fn in (("total_counts", np.sum), ("average_counts", np.mean), ("median_counts", np.median), ("max_counts", np.max),
                     ("min_counts", np.min)):
        out.append(f"#{stat}\t" + "\t".join(str(fn(a)) for a in cols))
    return "\n".join(out) + "\n"


def anonymize_sample(engine, reads: Sequence[dict], windows: Sequence[dict], reference: str, contig: str = "c") -> Dict[str, str]:
    """One contig of a tumor-normal sample through the engine: plan, one masking pass over every session, FASTQ
    rendering on the device.  reads: tumor then normal, each in file order, dicts with name / flag / dataset / pos /
    cigar / seq / qual.  Returns the reference's output files as {suffix: text}: "T.1", "T.2", "T.single_end",
    "N.1", "N.2", "N.single_end", "statistics"."""
    import re
    from . import batch as B
    cig = re.compile(r"(\d+)([MIDNSHP=X])")
    rs = [dict(r, end=r["pos"] + sum(int(n) for n, op in cig.findall(r["cigar"]) if op in "MDN=X")) for r in reads]
    batch = B.pack_reads(rs)                                      # dense qualities: every read is printed
    return anonymize_packed(engine, batch, [r["name"] for r in rs], rs, windows, reference, contig)


_CODE2ASC = "=ACMGRSVTWYHKDBN"
_ASC2CODE = {c: k for k, c in enumerate(_CODE2ASC)}
_COMPLEMENT = [int(f"{k:04b}"[::-1], 2) for k in range(16)]       # the complement of a 4-bit base code is its bit reversal


def reapply_indel_edits(codes, printed_quals, reverse: bool, edits, reference) -> Tuple[List[int], List[int]]:
    """Quirk Q12: the second application of a read's left-over indels, as the reference performs it
    (anonymizer_methods.py:254-270 over the list that :281-287 switched on again; mask_or_modify_indel, :178-203).
    codes: the record's 4-bit base codes in alignment orientation, printed_quals: its qualities in printed (BAM) order, both
    as the first application left them; edits: batch.parse_edits of the record.  Returns (codes, printed qualities)."""
    seq = [int(c) for c in codes]
    q = [int(x) for x in (printed_quals[::-1] if reverse else printed_quals)]     # the reference edits the forward-orientation array (quirk Q1)
    for at, pos, ln, ins in edits:                                 # application order: all DELs, then all INSs
        if ins:                                                    # the inserted bases go (Python slices clamp)
            seq, q = seq[:at] + seq[at + ln:], q[:at] + q[at + ln:]
        else:                                                      # the deleted reference bases come back with floor(mean(qualities))
            ref = reference[pos:pos + ln]
            ref = ref.decode("ascii") if isinstance(ref, (bytes, bytearray)) else str(ref)
            avg = int(sum(q) / len(q)) if q else 0
            seq, q = seq[:at] + [_ASC2CODE.get(ch, 15) for ch in ref.upper()] + seq[at:], q[:at] + [avg] * ln + q[at:]
    return seq, (q[::-1] if reverse else q)


def fastq_text(name: bytes, flag: int, codes: Sequence[int], printed_quals: Sequence[int]) -> bytes:
    """One FASTQ record as the reference prints it (anonymizer_methods.py:205-243): reverse reads reverse-complemented."""
    if flag & 0x10:
        codes = [_COMPLEMENT[c] for c in reversed(codes)]
    return (b"@" + name + b"/" + (b"1" if flag & 0x40 else b"2") + b"\n" + "".join(_CODE2ASC[c] for c in codes).encode("ascii") + b"\n+\n" +
            bytes(int(x) + 33 for x in printed_quals) + b"\n")


def anonymize_packed(engine, batch, names, read_table: Sequence[dict], windows: Sequence[dict], reference,
                     contig: str = "c", plan: Optional[Plan] = None, as_bytes: bool = False, carry: Optional[dict] = None) -> Dict[str, str]:
    """Same for an already packed batch (batch.ReadBatch with dense qualities, e.g. from
    genome_files.pack_tumor_normal): names = list of str or (uint8 blob, int64 offsets); read_table = rows with name /
    flag / dataset / pos / end for the planner (not needed when `plan` is given); reference = str / bytes / uint8 array
    of the contig; as_bytes leaves the six FASTQ texts as bytes (the file writer appends them as they are); plan may be
    a concurrent.futures future of a Plan.
    carry: the sample's unpaired reads so far, {name bytes: (mate, dataset, FASTQ record)} in insertion order - the
    reference keeps them across contigs (to_pair_anonymized_reads, short_read_tumor_normal_anonymizer.py:646) and writes a
    pair the moment its second mate is processed.  With carry given (names must be the (blob, offsets) form and the plan
    a native one, whose singles say where such a pair goes) the unpaired reads of this contig join it instead of being
    returned as single-end text, and pairs completed by a carried mate are spliced into the pair files."""
    import numpy as np
    import torch
    from . import batch as B
    from .engine import DeviceBatch, DeviceResult, DeviceSessions
    engine.upload_reference(batch.contig_id, reference if isinstance(reference, (str, bytes)) else np.ascontiguousarray(reference, np.uint8).tobytes())
    db = DeviceBatch(batch, engine.device)                            # the reads travel while a plan handed over as a future is still being made
    if plan is None:
        plan = plan_sample(read_table, windows, len(reference))
    elif hasattr(plan, "result"):
        plan = plan.result()
    sessions = B.pack_sessions(plan.sessions)
    ds = DeviceSessions(sessions, engine.device)
    units = batch.seq4.shape[0] // 16
    dres = DeviceResult(sessions.n_sessions, 2 * batch.n_reads + 16, 2 * units + 64, 2 * units + 64, engine.device)
    P = np.asarray(plan.pairs, np.int64).reshape(-1, 5)
    S = np.asarray(plan.singles, np.int64)
    S = S.reshape(-1, S.shape[1] if S.ndim == 2 and S.shape[0] else 3)
    ver_cols = np.concatenate([P[:, 2], P[:, 4], S[:, 2]])
    any_reapply = bool(((ver_cols >= 0) & ((ver_cols & REAPPLY) != 0)).any())
    if hasattr(engine, "keep_edits"):
        engine.keep_edits(any_reapply)                                # quirk Q12: the edit descriptions of this run are needed on the host
    engine.run_device(db, ds, dres)
    # (while the device runs) items in file order: T.1, T.2, N.1, N.2, T.single_end, N.single_end - each file is one slice
    # of the rendered text
    groups = []
    for d in (0, 1):
        rows = P[P[:, 0] == d]
        groups += [(f"{'TN'[d]}.1", rows[:, 1], rows[:, 2]), (f"{'TN'[d]}.2", rows[:, 3], rows[:, 4])]
    for d in (0, 1):
        rows = S[S[:, 0] == d]
        groups.append((f"{'TN'[d]}.single_end", rows[:, 1], rows[:, 2]))
    item_read = np.concatenate([g[1] for g in groups]) if groups else np.zeros(0, np.int64)
    item_ver = np.concatenate([g[2] for g in groups]) if groups else np.zeros(0, np.int64)
    reapply = np.nonzero((item_ver >= 0) & ((item_ver & REAPPLY) != 0))[0] if any_reapply else np.zeros(0, np.int64)
    if any_reapply:
        item_ver = np.where(item_ver >= 0, item_ver & ~REAPPLY, item_ver)
    item_read32 = item_read.astype(np.int32)
    torch.cuda.synchronize(engine.device)
    n = int(engine.check_device_status(dres).n_modified)
    # which modified record (if any) each planned read prints: only reads that have a record at all are searched, by
    # (session, read) key in the sorted keys of the records
    mod_read = dres.mod_read[:n].cpu().numpy().astype(np.int64)
    mod_key = (dres.mod_session[:n].cpu().numpy().astype(np.int64) << 32) | mod_read
    by_key = np.argsort(mod_key, kind="stable")
    sorted_key = mod_key[by_key]
    has_record = np.zeros(batch.n_reads + 1, bool)
    has_record[mod_read] = True
    item_rec = np.full(len(item_read), -1, np.int32)
    cand = np.nonzero(has_record[item_read] & (item_ver >= 0))[0]
    if len(cand):
        want = (item_ver[cand] << 32) | item_read[cand]
        at = np.minimum(np.searchsorted(sorted_key, want), n - 1)
        found = sorted_key[at] == want
        item_rec[cand[found]] = by_key[at[found]]
    if os.environ.get("GA_FILE_TRACE"):
        print(f"[ga-file-trace]   masked, {n} records, items ready", flush=True)
    text, off = engine.render_fastq(db, names, item_read32, item_rec, dres, n, as_view=as_bytes)   # as_bytes: slices of the download buffer, no copies
    starts = np.concatenate([[0], np.cumsum([len(g[1]) for g in groups])]).astype(np.int64)
    # quirk Q12: the few planned reads whose left-over indels the reference applies twice are printed on the host
    replaced: Dict[int, bytes] = {}
    if len(reapply) and hasattr(engine, "record_edits"):
        cand = [int(k) for k in reapply if item_rec[k] >= 0]
        if cand:
            recs = np.asarray([int(item_rec[k]) for k in cand], np.int64)
            ridx = torch.as_tensor(recs, device=dres.mod_len.device)
            qoff = dres.mod_qual_off16[ridx].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            soff = dres.mod_seq_off16[ridx].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            mlen = dres.mod_len[ridx].cpu().numpy().astype(np.int64)
            aux = engine.record_edits(recs)
            blob, noff = names if isinstance(names, tuple) else (None, None)
            for k, q16, s16, L1, a in zip(cand, qoff, soff, mlen, aux):
                edits = None if q16 == 0xFFFFFFFF else B.parse_edits(a)
                if not edits:
                    continue                                          # no indel edits (nothing to re-apply), or a description that was not kept
                r = int(item_read[k])
                packed = dres.out_seq4[16 * int(s16):16 * int(s16) + (int(L1) + 1) // 2].cpu().numpy()
                codes = np.stack([packed & 15, packed >> 4], 1).reshape(-1)[:int(L1)]
                quals = dres.out_qual[32 * int(q16):32 * int(q16) + int(L1)].cpu().numpy()
                flag = int(batch.len_flag[r]) >> 16
                codes2, quals2 = reapply_indel_edits(codes, quals, bool(flag & 0x10), edits, reference)
                nm = bytes(blob[int(noff[r]):int(noff[r + 1])]) if blob is not None else names[r].encode("ascii")
                replaced[k] = fastq_text(nm, flag, codes2, quals2)
    if any_reapply and hasattr(engine, "keep_edits"):
        engine.keep_edits(False)                                      # the engine is the caller's: later runs do not pay for it
    rep_at = sorted(replaced)

    def piece(a, b):                                                  # records [a, b) of the item list
        a, b = int(a), int(b)
        lo, hi = bisect.bisect_left(rep_at, a), bisect.bisect_left(rep_at, b)
        if lo == hi:
            return text[int(off[a]):int(off[b])]
        parts, cur = [], a
        for k in rep_at[lo:hi]:
            parts += [bytes(text[int(off[cur]):int(off[k])]), replaced[k]]
            cur = k + 1
        parts.append(bytes(text[int(off[cur]):int(off[b])]))
        return b"".join(parts)
    out = {}
    if carry is None:
        for g, (name, rd, _) in enumerate(groups):
            out[name] = piece(starts[g], starts[g + 1])
    else:
        blob, noff = names
        flags = np.asarray(batch.len_flag) >> 16
        splice = {0: [], 1: []}                                       # dataset -> [(pairs of the contig before it, mate-1 text, mate-2 text)]
        for d in (0, 1):
            g = 4 + d                                                 # this dataset's unpaired reads, in spill order
            rows = S[S[:, 0] == d]
            for k in range(len(rows)):
                r = int(rows[k, 1])
                nm = bytes(blob[int(noff[r]):int(noff[r + 1])])
                m = 0 if int(flags[r]) & 0x40 else 1
                rec = bytes(piece(starts[g] + k, starts[g] + k + 1))  # outlives the download buffer
                held = carry.get(nm)
                if held is None:
                    carry[nm] = (m, d, rec)
                elif held[0] != m and held[1] == d:                   # the second mate of a pair whose first mate came earlier
                    del carry[nm]
                    at = int(rows[k, 3]) if rows.shape[1] > 3 else len(P)
                    splice[d].append((at, held[2] if held[0] == 0 else rec, rec if held[0] == 0 else held[2]))
        for d in (0, 1):
            before = np.concatenate([[0], np.cumsum(P[:, 0] == d)]) if len(P) else np.zeros(1, np.int64)   # pairs of d among the first k pairs
            for m in (0, 1):
                g = 2 * d + m
                parts, cur = [], 0
                for at, t1, t2 in sorted(splice[d], key=lambda t: t[0]):   # stable: equal positions keep their order
                    k = int(before[min(at, len(P))])
                    parts.append(piece(starts[g] + cur, starts[g] + k))
                    parts.append(t1 if m == 0 else t2)
                    cur = k
                parts.append(piece(starts[g] + cur, starts[g + 1]))
                out[groups[g][0]] = parts[0] if len(parts) == 1 else (parts if as_bytes else b"".join(parts))   # a list: written piece by piece
            out[groups[4 + d][0]] = b""
    if not as_bytes:
        out = {k: bytes(v).decode("ascii") for k, v in out.items()}
    counts = dres.sess_counts.view(-1, 4)[:sessions.n_sessions].cpu().numpy()
    out["statistics"] = statistics_text(contig, plan, counts)
    out["_plan"] = plan
    out["_counts"] = counts
    return out
