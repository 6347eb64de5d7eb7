"""ctypes mirrors of the structs in include/ga_b200.h (keep field order identical)."""
import ctypes as C

GA_ABI_VERSION = 1
GA_OK, GA_ERR_BAD_ARGUMENT, GA_ERR_OFFSET_RANGE, GA_ERR_LENGTH_MISMATCH = 0, 1, 2, 3
GA_ERR_CUDA, GA_ERR_CAPACITY, GA_ERR_UNSUPPORTED, GA_ERR_NO_DEVICE = 4, 5, 6, 7
GA_VT_NONE, GA_VT_SNV, GA_VT_DEL, GA_VT_INS = 0, 1, 2, 3
VT_BY_NAME = {"SNV": GA_VT_SNV, "DEL": GA_VT_DEL, "INS": GA_VT_INS}

CODE2ASC = "=ACMGRSVTWYHKDBN"          # BAM 4-bit base codes
CIGAR_OPS = "MIDNSHP=XB"               # BAM CIGAR op codes

_vp = C.c_void_p


class GaReads(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_tumor", C.c_int64),
                ("pos", _vp), ("len_flag", _vp), ("seq_off16", _vp), ("cigar_off", _vp), ("cigar", _vp),
                ("seq4", _vp), ("qual", _vp), ("seq4_bytes", C.c_int64),
                ("n_qual", C.c_int64), ("qual_reads", _vp), ("qual_off16", _vp),
                ("max_ref_span", C.c_int32), ("contig_id", C.c_int32)]


class GaSessions(C.Structure):
    _fields_ = [("n_sessions", C.c_int32),
                ("first", _vp), ("last", _vp), ("keep_type", _vp), ("keep_pos", _vp), ("keep_end", _vp),
                ("keep_len", _vp), ("keep_allele_off", _vp), ("keep_alleles", _vp)]


class GaTotals(C.Structure):
    _fields_ = [("n_modified", C.c_uint64), ("seq16_used", C.c_uint64), ("qual16_used", C.c_uint64),
                ("session_reads", C.c_uint64), ("session_bases", C.c_uint64), ("indel_records", C.c_uint64),
                ("masked", C.c_uint64 * 3), ("error", C.c_uint32), ("error_detail", C.c_uint32)]


class GaResult(C.Structure):
    _fields_ = [("cap_records", C.c_int64), ("cap_seq16", C.c_int64), ("cap_qual16", C.c_int64),
                ("mod_session", _vp), ("mod_read", _vp), ("mod_len", _vp), ("mod_seq_off16", _vp), ("mod_qual_off16", _vp),
                ("out_seq4", _vp), ("out_qual", _vp), ("sess_counts", _vp), ("totals", _vp)]


class GaFastqItems(C.Structure):
    _fields_ = [("n_items", C.c_int64), ("read", _vp), ("record", _vp), ("names", _vp), ("name_off", _vp)]


class GaWireDir(C.Structure):
    _fields_ = [("byte", C.c_uint64), ("read", C.c_uint32), ("unit", C.c_uint32), ("ops", C.c_uint32), ("pos", C.c_int32),
                ("reserved", C.c_uint32 * 2)]


class GaReadsWire(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_tumor", C.c_int64), ("n_blocks", C.c_int64), ("n_tumor_blocks", C.c_int64),
                ("blob", _vp), ("blob_bytes", C.c_int64), ("dir", _vp),
                ("qual", _vp), ("n_qual", C.c_int64), ("qual_reads", _vp), ("qual_off16", _vp), ("qual_units", C.c_int64),
                ("max_ref_span", C.c_int32), ("contig_id", C.c_int32)]


class GaDigestIds(C.Structure):
    _fields_ = [("session_base", C.c_int64), ("tumor_base", C.c_int64), ("normal_base", C.c_int64), ("n_tumor", C.c_int64),
                ("contig", C.c_int64)]


class GaSynthParams(C.Structure):
    _fields_ = [("contig_len", C.c_int64), ("seed", C.c_uint64), ("read_len", C.c_int32),
                ("total_windows", C.c_int32), ("window_begin", C.c_int32), ("n_windows", C.c_int32),
                ("window_half", C.c_int32), ("max_indel", C.c_int32), ("max_clip", C.c_int32), ("depth_var_pct", C.c_int32),
                ("cov_tumor", C.c_float), ("cov_normal", C.c_float),
                ("snp_rate", C.c_float), ("indel_rate", C.c_float), ("err_rate", C.c_float),
                ("n_rate", C.c_float), ("somatic_vaf", C.c_float), ("clip_frac", C.c_float)]


class GaSynthPlan(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_tumor", C.c_int64), ("seq4_bytes", C.c_int64),
                ("reads_per_window", C.c_int32 * 2), ("units_per_read", C.c_int32), ("window_stride", C.c_int32)]


TOTALS_BYTES = C.sizeof(GaTotals)


class GaError(RuntimeError):
    def __init__(self, status, msg=""):
        super().__init__(f"ga status {status}: {msg}")
        self.status = status


def raise_for_status(status, msg=""):
    """Map ABI status codes to the exception types the reference raises (SURVEY.md 8(b) Errors)."""
    if status == GA_OK:
        return
    if status in (GA_ERR_BAD_ARGUMENT, GA_ERR_LENGTH_MISMATCH):
        raise ValueError(f"ga status {status}: {msg}")
    if status == GA_ERR_OFFSET_RANGE:
        raise IndexError(f"ga status {status}: {msg}")
    raise GaError(status, msg)
