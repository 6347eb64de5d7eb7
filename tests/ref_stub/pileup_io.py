"""Run-time de-cythonisation of the reference's pileup_io.pyx.

TEST INFRASTRUCTURE ONLY, usable only where /root/reference is mounted (this
container).  The .pyx bodies are plain Python once the two `cimport` lines and
the `cdef:` typed-declaration blocks are dropped and `cdef f(` becomes `def f(`;
this module does that transformation in memory and exec()s the result, so no
reference source is copied into this repository.
"""
import os
import re

_REF = os.environ.get("GA_REFERENCE_ROOT", "/root/reference")
_PYX = os.path.join(_REF, "src", "GenomeAnonymizer", "pileup_io.pyx")


def _decythonize(text):
    out = []
    lines = text.split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        s = ln.strip()
        if (s.startswith("cimport ") or (s.startswith("from ") and " cimport " in s)
                or s.startswith("from pysam.libc")):
            # multi-line cimport with trailing backslash
            while ln.rstrip().endswith("\\"):
                i += 1
                ln = lines[i]
            i += 1
            continue
        if s == "cdef:":
            indent = len(ln) - len(ln.lstrip())
            i += 1
            while i < len(lines) and (not lines[i].strip() or
                                      len(lines[i]) - len(lines[i].lstrip()) > indent):
                i += 1
            continue
        ln = re.sub(r"^(\s*)cdef (\w+\()", r"\1def \2", ln)
        out.append(ln)
        i += 1
    return "\n".join(out)


with open(_PYX) as _f:
    _src = _decythonize(_f.read())
exec(compile(_src, _PYX, "exec"), globals())
