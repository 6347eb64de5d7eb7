"""Builds libga_b200.so (hand-written CUDA for sm_100a + the C ABI) and libga_synth.so (synthetic input
generator, include/ga_synth.h) in-tree with nvcc.  Translation units are compiled side by side."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libga_b200.so")
SYNTH_LIB = os.path.join(HERE, "libga_synth.so")
SOURCES = ["ga_engine.cu", "ga_host_pipeline.cu", "ga_wire.cu", "ga_fastq.cu", "ga_genome_io.cpp", "ga_plan.cpp"]
SYNTH_SOURCES = ["ga_synth.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources(names=SOURCES):
    return [os.path.join(CSRC, s) for s in names if os.path.exists(os.path.join(CSRC, s))]


def _deps():
    inc = os.path.join(HERE, "..", "include")
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(inc, f) for f in os.listdir(inc)]


def needs_build():
    for lib in (LIB, SYNTH_LIB):
        if not os.path.exists(lib):
            return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(SYNTH_LIB))
    return any(os.path.getmtime(d) > t for d in _deps())


def _run(cmd, verbose):
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed: " + " ".join(cmd[-3:]))
    if verbose:
        sys.stderr.write(res.stderr)


def _compile(src, verbose, force):
    obj = os.path.join(OBJ, os.path.basename(src) + ".o")
    if not force and os.path.exists(obj) and all(os.path.getmtime(obj) >= os.path.getmtime(d) for d in _deps()):
        return obj
    nvcc = os.environ.get("NVCC", "nvcc")
    _run([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src], verbose)
    return obj


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    srcs, ssrcs = sources(SOURCES), sources(SYNTH_SOURCES)
    with ThreadPoolExecutor(max_workers=len(srcs) + len(ssrcs)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose, force), srcs + ssrcs))
    _run([nvcc] + NVCC_FLAGS + ["-shared", "-o", LIB] + objs[:len(srcs)] + ["-lz"], verbose)
    _run([nvcc] + NVCC_FLAGS + ["-shared", "-o", SYNTH_LIB] + objs[len(srcs):], verbose)
    return LIB


if __name__ == "__main__":
    build_library(force=True, verbose="-v" in sys.argv)
    print(LIB)
    print(SYNTH_LIB)
