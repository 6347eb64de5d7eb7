"""VariantType order is the one the reference's statistics header implies
(short_read_tumor_normal_anonymizer.py:200,218-219): SNV,DEL,INS,DUP,INV,CNV,TRA,SGL."""
from dataclasses import dataclass
from enum import Enum
from typing import Any, Optional


class VariantType(Enum):
    SNV = 1
    DEL = 2
    INS = 3
    DUP = 4
    INV = 5
    CNV = 6
    TRA = 7
    SGL = 8


@dataclass
class VariantRecord:
    contig: str
    pos: int          # 1-based
    end: int
    length: int
    ref: str
    alt: str
    variant_type: VariantType
    alt_sv_breakend: Optional[Any] = None
    info: Optional[dict] = None
