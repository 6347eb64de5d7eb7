"""CPU-side checks: the C-ABI library loads and exports every symbol include/ga_b200.h declares, the ctypes
mirrors match the C structs, the packer validates its input, the plugin's host objects behave like the
reference's, and the masking path refuses to run without a GPU."""
import ctypes as C
import os
import pickle
import re
import subprocess

import numpy as np
import pytest

from genomeanonymizer_b200 import _abi, _lib
from genomeanonymizer_b200 import batch as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ga_b200.h")


SYNTH_HEADER = os.path.join(ROOT, "include", "ga_synth.h")


def declared_functions(header=HEADER):
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ga_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)                                    # loads without a GPU; no compute call is made
    names = declared_functions()
    assert len(names) >= 17
    wire = declared_functions(os.path.join(ROOT, "include", "ga_wire.h"))
    assert sorted(_lib.WIRE_EXPORTS) == wire and not [n for n in wire if not hasattr(L, n)]
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names                         # the loader binds exactly the declared surface


def test_synth_library_is_separate_and_exports_its_header():
    """The input generator lives in libga_synth.so (include/ga_synth.h): a process that only needs input data never
    maps the masking engine, and the engine library carries no generator symbol."""
    L = C.CDLL(_lib.SYNTH_LIB_PATH)
    names = declared_functions(SYNTH_HEADER)
    assert len(names) == 9 and sorted(_lib.SYNTH_EXPORTS) == names
    assert not [n for n in names if not hasattr(L, n)]
    E = C.CDLL(_lib.LIB_PATH)
    assert not [n for n in names if hasattr(E, n)]


def test_abi_version_and_status_strings():
    L = _lib.lib()
    assert L.ga_abi_version() == _abi.GA_ABI_VERSION
    assert L.ga_status_string(0) == b"ok"
    assert b"capacity" in L.ga_status_string(_abi.GA_ERR_CAPACITY)


def test_ctypes_mirrors_match_the_c_structs(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "ga_synth.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(ga_reads),'
                   ' sizeof(ga_sessions), sizeof(ga_totals), sizeof(ga_result), sizeof(ga_synth_params), sizeof(ga_synth_plan));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirrors = [_abi.GaReads, _abi.GaSessions, _abi.GaTotals, _abi.GaResult, _abi.GaSynthParams, _abi.GaSynthPlan]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_engine_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from genomeanonymizer_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)
    h = C.c_void_p()
    assert _lib.lib().ga_engine_create(0, C.byref(h)) == _abi.GA_ERR_NO_DEVICE     # the C ABI says so too


def test_packer_rejects_inconsistent_reads():
    ok = {"name": "r", "flag": 0x43, "pos": 5, "cigar": "4M", "seq": "ACGT", "qual": [1, 2, 3, 4], "dataset": 0}
    B.pack_reads([ok])
    with pytest.raises(ValueError):
        B.pack_reads([dict(ok, qual=[1, 2, 3])])                 # AM.py:200-201 length check
    with pytest.raises(ValueError):
        B.pack_reads([dict(ok, cigar="5M")])
    with pytest.raises(ValueError):
        B.pack_reads([dict(ok, pos=9), dict(ok, pos=3)])         # must be coordinate sorted
    with pytest.raises(ValueError):
        B.pack_sessions([{"first": 10, "last": 20, "keep": None}, {"first": 5, "last": 9, "keep": None}])


def test_nibble_layout_low_nibble_first():
    b = B.pack_reads([{"name": "r", "flag": 0x43, "pos": 0, "cigar": "5M", "seq": "ACGTN", "qual": [9] * 5, "dataset": 0}])
    assert list(b.seq4[:3]) == [0x21, 0x84, 0x0F] and b.seq4.shape[0] == 16
    assert B.decode_bases(b.sequence_codes(0)) == "ACGTN"
    assert b.len_flag[0] == (0x43 << 16) | 5 and b.max_ref_span == 5


def test_anonymized_read_renders_like_the_reference():
    from genomeanonymizer_b200.anonymizer_methods import AnonymizedRead, anonymized_read_pair_is_writeable
    fwd = AnonymizedRead("q", True, False, False, 0, np.frombuffer(b"ACGTN", dtype=np.uint8).copy(), [1, 2, 3, 4, 5])
    assert fwd.get_anonymized_fastq_record() == "@q/1\nACGTN\n+\n\"#$%&"
    rev = AnonymizedRead("q", False, True, True, 1, np.frombuffer(b"AACGT", dtype=np.uint8).copy(), [5, 4, 3, 2, 1])
    assert rev.get_pair_idx() == 1
    # reverse reads: sequence reverse-complemented, the forward-orientation qualities reversed again (AM.py:213: Q1)
    assert rev.get_anonymized_fastq_record() == "@q/2\nACGTT\n+\n\"#$%&"
    assert anonymized_read_pair_is_writeable(fwd, rev) and not anonymized_read_pair_is_writeable(fwd, None)


def test_plugin_pickles_without_engine_state():
    from genomeanonymizer_b200.anonymizer_methods import B200GermlineAnonymizer, CompleteGermlineAnonymizer
    a = B200GermlineAnonymizer(device=3)
    b = pickle.loads(pickle.dumps(a))
    assert b.device == 3 and b._engine is None and b.anonymized_reads == {}
    assert CompleteGermlineAnonymizer is B200GermlineAnonymizer


def test_window_sharding_is_a_partition():
    from genomeanonymizer_b200.synthdev import shard_windows
    for total in (1, 7, 50_000):
        for world in (1, 2, 3, 8):
            spans = [shard_windows(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
