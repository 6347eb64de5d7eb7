"""CPU restatement of the reference's FASTQ rendering - TEST INFRASTRUCTURE ONLY (see oracle/ga_oracle.c).

Follows AnonymizedRead.get_anonymized_fastq_record (anonymizer_methods.py:205-243), generate_anonymized_read
(anonymizer_methods.py:57-58), the complement table (anonymizer_methods.py:22) and the writer's newline
(short_read_tumor_normal_anonymizer.py:157-158).  Pinned by tests/test_oracle_golden.py against the "fastq" strings the
reference itself produced (tests/golden/session_cases.json)."""

CODE2ASC = "=ACMGRSVTWYHKDBN"
# anonymizer_methods.py:22 knows A, C, G, T, N only; the IUPAC complement of the other BAM codes (= the code with
# its 4 bits reversed) is this repository's documented extension
_COMPLEMENT = {CODE2ASC[c]: CODE2ASC[int(f"{c:04b}"[::-1], 2)] for c in range(16)}


def reverse_complement(seq: str) -> str:
    return "".join(_COMPLEMENT[b] for b in reversed(seq))


def render(name: str, flag: int, seq_alignment: str, qual_printed) -> str:
    """One record as written to the .fastq file, newline included.

    seq_alignment: final sequence in alignment (BAM) orientation, as the reference keeps it internally
    (anonymizer_methods.py:163).  qual_printed: final qualities in printed order.  The reference stores qualities in
    original-read orientation (get_forward_qualities, anonymizer_methods.py:95) and reverses them again when it prints
    a reverse read (anonymizer_methods.py:213), so the printed order is the BAM order."""
    reverse = bool(flag & 0x10)
    forward_quals = list(reversed(qual_printed)) if reverse else list(qual_printed)     # anonymizer_methods.py:95
    seq = seq_alignment
    quals = forward_quals
    if reverse:                                                                           # anonymizer_methods.py:205-214
        seq = reverse_complement(seq)
        quals = list(reversed(quals))
    pair = 1 if flag & 0x40 else 2                                                        # anonymizer_methods.py:218
    qual_s = "".join(chr(int(q) + 33) for q in quals)                                     # anonymizer_methods.py:232
    return f"@{name}/{pair}\n{seq}\n+\n{qual_s}\n"                                       # anonymizer_methods.py:57-58 + writer
