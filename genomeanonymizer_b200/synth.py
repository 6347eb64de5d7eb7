"""Small host-side synthetic tumor/normal generator (numpy) for parity tests and golden vectors.

Produces "cases": a reference contig, somatic-variant windows (sessions) and coordinate-sorted
tumor / normal read lists with the features the masking path has to survive: shared germline
SNPs / insertions / deletions (het and hom), tumor-only somatic SNVs, sequencing errors, `N`
bases, soft clips, reverse-strand mates.  Inputs the reference cannot process (SURVEY.md
Appendix B: shared names across datasets, unpaired flags, H/N ops, supplementary records) are
never generated here; hand-written known-answer cases cover the quirks instead.

The full-size benchmark data are generated on the device (csrc/ga_synth.cu); this module is for
sizes a CPU oracle finishes in seconds.
"""
from __future__ import annotations

import numpy as np

BASES = "ACGT"


def _rand_bases(rng, n):
    return "".join(BASES[i] for i in rng.integers(0, 4, size=n))


def make_reference(rng, length, n_run_prob=0.0, lower_prob=0.0):
    ref = list(_rand_bases(rng, length))
    if n_run_prob > 0:
        p = 0
        while p < length:
            if rng.random() < n_run_prob:
                run = int(rng.integers(1, 30))
                for k in range(p, min(length, p + run)):
                    ref[k] = "N"
                p += run
            p += int(rng.integers(50, 400))
    if lower_prob > 0:
        for k in range(length):
            if rng.random() < lower_prob:
                ref[k] = ref[k].lower()
    return "".join(ref)


def make_germline(rng, ref, snp_rate=1e-3, indel_rate=1e-4, max_indel=10, het_frac=0.6):
    """dict pos -> (kind, payload, hapmask); kind in 'S','I','D'.  hapmask bit h set = present on hap h."""
    var = {}
    L = len(ref)
    n_snp = rng.binomial(L, snp_rate)
    for p in rng.integers(1, L - 1, size=n_snp):
        p = int(p)
        r = ref[p].upper()
        if r not in BASES:
            continue
        alt = BASES[(BASES.index(r) + 1 + int(rng.integers(0, 3))) % 4]
        hap = int(rng.integers(1, 3)) if rng.random() < het_frac else 3
        var[p] = ("S", alt, hap)
    n_indel = rng.binomial(L, indel_rate)
    for p in rng.integers(max_indel + 2, L - max_indel - 2, size=n_indel):
        p = int(p)
        if any((p + d) in var for d in range(-max_indel - 1, max_indel + 2)):
            continue
        ln = int(rng.integers(1, max_indel + 1))
        hap = int(rng.integers(1, 3)) if rng.random() < het_frac else 3
        if rng.random() < 0.5:
            var[p] = ("I", _rand_bases(rng, ln), hap)
        else:
            var[p] = ("D", ln, hap)
    return var


def _merge_cigar(ops):
    out = []
    for op, n in ops:
        if n == 0:
            continue
        if out and out[-1][0] == op:
            out[-1][1] += n
        else:
            out.append([op, n])
    return out


def build_read(rng, ref, start, L, hap, germline, somatic, is_tumor, vaf, err_rate, n_rate):
    """Walk the reference from `start`, applying the haplotype's variants; returns (pos, cigar ops, seq)."""
    seq = []
    ops = []
    p = start
    G = len(ref)
    while len(seq) < L and p < G:
        v = germline.get(p)
        if v is not None and (v[2] >> hap) & 1:
            kind, payload, _ = v
            if kind == "I" and seq:
                room = L - len(seq)
                ins = payload
                if room <= len(ins):
                    # read ends inside the insertion: aligner would soft-clip it
                    seq.extend(ins[:room])
                    ops.append(("S", room))
                    break
                seq.extend(ins)
                ops.append(("I", len(ins)))
                # fall through to emit ref base p
            elif kind == "D" and seq:
                if p + payload < G:
                    ops.append(("D", payload))
                    p += payload
                    continue
            elif kind == "S":
                seq.append(payload)
                ops.append(("M", 1))
                p += 1
                continue
        b = ref[p].upper()
        if is_tumor and p in somatic and rng.random() < vaf:
            b = somatic[p]
        seq.append(b)
        ops.append(("M", 1))
        p += 1
    # a trailing D (deletion followed by nothing) cannot happen: D is only emitted when bases follow
    while ops and ops[-1][0] == "D":
        ops.pop()
    # sequencing errors on aligned / inserted bases alike
    for i in range(len(seq)):
        u = rng.random()
        if u < n_rate:
            seq[i] = "N"
        elif u < n_rate + err_rate:
            seq[i] = BASES[(BASES.index(seq[i]) + 1 + int(rng.integers(0, 3))) % 4] if seq[i] in BASES else seq[i]
    return start, _merge_cigar(ops), "".join(seq)


def soft_clip(rng, pos, ops, seq, max_clip):
    """Turn the head or tail of the first/last M op into a soft clip with random bases."""
    ops = [list(o) for o in ops]
    seq = list(seq)
    if rng.random() < 0.5:
        if ops[0][0] == "M" and ops[0][1] > 2:
            k = int(rng.integers(1, min(max_clip, ops[0][1] - 1) + 1))
            ops[0][1] -= k
            ops.insert(0, ["S", k])
            pos += k
            seq[:k] = list(_rand_bases(rng, k))
    else:
        if ops[-1][0] == "M" and ops[-1][1] > 2:
            k = int(rng.integers(1, min(max_clip, ops[-1][1] - 1) + 1))
            ops[-1][1] -= k
            ops.append(["S", k])
            seq[len(seq) - k:] = list(_rand_bases(rng, k))
    return pos, _merge_cigar([(o, n) for o, n in ops]), "".join(seq)


def cigar_string(ops):
    return "".join(f"{n}{op}" for op, n in ops)


def make_case(seed, contig_len=6000, n_pairs=(150, 150), read_len=100, somatic_positions=None,
              snp_rate=2e-3, indel_rate=6e-4, err_rate=2e-3, n_rate=3e-4, clip_frac=0.15,
              vaf=0.4, window_half=1000, ref_n_runs=0.0, ref_lower=0.0, qual_range=(2, 41),
              name="case", keep_somatic=True, max_indel=8):
    """Returns a dict describing one contig with its sessions (windows) and T/N reads."""
    rng = np.random.default_rng(seed)
    ref = make_reference(rng, contig_len, ref_n_runs, ref_lower)
    germ = make_germline(rng, ref, snp_rate, indel_rate, max_indel)
    if somatic_positions is None:
        somatic_positions = [contig_len // 2]
    somatic = {}
    for p in somatic_positions:
        r = ref[p].upper()
        if r not in BASES:
            r = "A"
        somatic[p] = BASES[(BASES.index(r) + 1 + int(rng.integers(0, 3))) % 4]
        germ.pop(p, None)
    reads = []
    for ds, npairs in enumerate(n_pairs):
        prefix = "T" if ds == 0 else "N"
        lst = []
        for i in range(npairs):
            insert = int(np.clip(rng.normal(2.8 * read_len, read_len / 3), read_len + 10, 5 * read_len))
            fs = int(rng.integers(0, max(1, contig_len - insert - 2 * max_indel - 2)))
            hap = int(rng.integers(0, 2))
            first_is_r1 = rng.random() < 0.5
            for mate in range(2):
                start = fs if mate == 0 else fs + insert - read_len
                pos, ops, seq = build_read(rng, ref, start, read_len, hap, germ, somatic, ds == 0,
                                           vaf, err_rate, n_rate)
                if rng.random() < clip_frac:
                    pos, ops, seq = soft_clip(rng, pos, ops, seq, max(2, read_len // 3))
                flag = 0x1 | 0x2
                reverse = (mate == 1)
                flag |= 0x10 if reverse else 0x20
                is_r1 = (mate == 0) == first_is_r1
                flag |= 0x40 if is_r1 else 0x80
                qual = rng.integers(qual_range[0], qual_range[1], size=len(seq)).tolist()
                lst.append({"name": f"{prefix}{i}", "flag": flag, "pos": pos, "cigar": cigar_string(ops),
                            "seq": seq, "qual": qual, "dataset": ds})
        lst.sort(key=lambda r: r["pos"])     # stable: file (coordinate) order
        reads.extend(lst)
    windows = []
    for p in sorted(somatic):
        keep = None
        if keep_somatic:
            keep = {"type": "SNV", "pos": p, "end": p, "length": 1, "allele": somatic[p]}
        windows.append({"first": p + 1 - window_half, "last": p + 1 + window_half + 1, "keep": keep})
    return {"name": name, "contig": "c", "reference": ref, "windows": windows, "reads": reads}
