"""Build libbtbb.so.1 (B200) in-tree with nvcc for sm_100a.

    python -m libbtbb_b200.build [--force] [--ptxas-verbose]

The shared object lands in libbtbb_b200/lib/libbtbb.so.1 (SONAME libbtbb.so.1, the
name upstream installs, lib/src/CMakeLists.txt:43-52) so it travels with the source
snapshot to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libbtbb.so.1")
SOURCES = ["capi.cu", "tables.cu", "find_ac.cu", "decode.cu", "sieve.cu", "synth.cu", "compat.cu", "host_pack.cpp", "pcap_out.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=default", "--shared",
    "-Xlinker", "-soname=libbtbb.so.1",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", f) for f in ("btbb_b200.h", "btbb.h")]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_verbose=False):
    if not force and not _stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if ptxas_verbose else [])
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libbtbb.so.1")
    if ptxas_verbose or verbose:
        sys.stderr.write(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True, ptxas_verbose="--ptxas-verbose" in sys.argv)
    print(LIB)
