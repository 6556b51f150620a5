"""Build libbtbb.so.1 (B200) in-tree with nvcc for sm_100a.

    python -m libbtbb_b200.build [--force] [--ptxas-verbose]
    python -m libbtbb_b200.build --compose-reference <upstream checkout>     # lib/full/libbtbb.so.1
    python -m libbtbb_b200.build --install <prefix> [--full]                 # upstream's install layout

The shared object lands in libbtbb_b200/lib/libbtbb.so.1 (SONAME libbtbb.so.1, the
name upstream installs, lib/src/CMakeLists.txt:43-52) so it travels with the source
snapshot to the GPU box.  Sources are compiled one object each (in parallel, rebuilt only
when the source or a header changed) and linked; libbtbb.a (upstream's optional static
library, lib/src/CMakeLists.txt:54-63) is archived from the same objects.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libbtbb.so.1")
STATIC = os.path.join(LIBDIR, "libbtbb.a")
SOURCES = ["capi.cu", "tables.cu", "find_ac.cu", "decode.cu", "decode_tables.cpp", "decode_host.cpp", "find_ac_host.cpp", "sieve.cu",
           "hops.cu", "synth.cu", "compat.cu", "host_pack.cpp", "pcap_out.cpp", "pcap_dev.cu", "sharded.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC,-fvisibility=default"]
LDFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xlinker", "-soname=libbtbb.so.1"]


def _headers_mtime():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps += [os.path.join(HERE, "..", "include", f) for f in ("btbb_b200.h", "btbb.h")]
    deps.append(os.path.abspath(__file__))
    return max(os.path.getmtime(d) for d in deps)


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build(force=False, verbose=False, ptxas_verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_t = _headers_mtime()
    jobs, objs = [], []
    for s in _sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s + ".o")
        objs.append(obj)
        if force or ptxas_verbose or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            cmd = [NVCC] + CFLAGS + (["-Xptxas", "-v"] if ptxas_verbose else []) + ["-c", src, "-o", obj]
            jobs.append((s, cmd))

    def run(job):
        name, cmd = job
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name, r in ex.map(run, jobs):
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f"nvcc failed compiling {name}")
                if ptxas_verbose or verbose:
                    sys.stderr.write(r.stdout + r.stderr)
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC] + LDFLAGS + objs + ["-o", LIB]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking libbtbb.so.1")
        if os.path.exists(STATIC):
            os.remove(STATIC)
        subprocess.run(["ar", "rcs", STATIC] + objs, check=False)
    _dev_files()
    return LIB


def _dev_files():
    """What upstream's install step adds for callers that build against the library: the unversioned
    libbtbb.so link (lib/src/CMakeLists.txt:43-52) and libbtbb.pc (lib/libbtbb.pc.in:6-10), here with
    this tree as the prefix -- `PKG_CONFIG_PATH=libbtbb_b200/lib/pkgconfig pkg-config --cflags --libs libbtbb`."""
    link = os.path.join(LIBDIR, "libbtbb.so")
    if not os.path.islink(link):
        if os.path.exists(link):
            os.remove(link)
        os.symlink("libbtbb.so.1", link)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")
    pc = ("prefix=%s\nlibdir=%s\nincludedir=${prefix}/include\n\n"
          "Name: Bluetooth Baseband Library\nDescription: C Utility Library (B200 build of the packet layer)\n"
          "Version: 0.1-b200\nCflags: -I${includedir}/\nLibs: -L${libdir} -lbtbb\n"
          "Libs.private: -L%s -lcudart -lstdc++ -lm -ldl -lpthread\n") % (os.path.abspath(os.path.join(HERE, "..")), LIBDIR, cuda_lib)
    os.makedirs(os.path.join(LIBDIR, "pkgconfig"), exist_ok=True)
    path = os.path.join(LIBDIR, "pkgconfig", "libbtbb.pc")
    if not os.path.exists(path) or open(path).read() != pc:
        with open(path, "w") as f:
            f.write(pc)


def install(prefix, full=False):
    """Upstream's install layout under `prefix` (lib/src/CMakeLists.txt:43-68, lib/libbtbb.pc.in): include/btbb.h
    (+ btbb_b200.h), lib/libbtbb.so.1, the libbtbb.so link, libbtbb.a, lib/pkgconfig/libbtbb.pc -- what
    pkg-config and upstream's cmake/modules/FindBTBB.cmake (LIBBTBB_DIR=<prefix>) look for.  full=True installs
    the composed library (every symbol of upstream's btbb.h) instead of the packet layer alone."""
    import shutil
    build()
    src_lib = FULL_LIB if full else LIB
    if not os.path.exists(src_lib):
        raise RuntimeError(src_lib + " has not been built")
    inc, lib = os.path.join(prefix, "include"), os.path.join(prefix, "lib")
    os.makedirs(inc, exist_ok=True)
    os.makedirs(os.path.join(lib, "pkgconfig"), exist_ok=True)
    for h in ("btbb.h", "btbb_b200.h"):
        shutil.copy(os.path.join(HERE, "..", "include", h), inc)
    shutil.copy(src_lib, os.path.join(lib, "libbtbb.so.1"))
    link = os.path.join(lib, "libbtbb.so")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink("libbtbb.so.1", link)
    if os.path.exists(STATIC) and not full:
        shutil.copy(STATIC, lib)
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")
    with open(os.path.join(lib, "pkgconfig", "libbtbb.pc"), "w") as f:
        f.write("prefix=%s\nexec_prefix=${prefix}\nlibdir=${prefix}/lib\nincludedir=${prefix}/include\n\n"
                "Name: Bluetooth Baseband Library\nDescription: C Utility Library (B200 build of the packet layer)\n"
                "Version: 0.1-b200\nCflags: -I${includedir}/\nLibs: -L${libdir} -lbtbb\n"
                "Libs.private: -L%s -lcudart -lstdc++ -lm -ldl -lpthread\n" % (os.path.abspath(prefix), cuda_lib))
    return prefix


FULL_LIB = os.path.join(LIBDIR, "full", "libbtbb.so.1")
REF_UNITS = ["bluetooth_piconet.c", "bluetooth_le_packet.c", "companies.c", "pcap.c", "pcapng.c", "pcapng-bt.c"]


def compose(reference_root="/root/reference", verbose=False):
    """The COMPLETE libbtbb.so.1 surface (every symbol of upstream's btbb.h): this library's packet
    layer plus upstream's own piconet / pcap / pcapng / LE translation units, compiled UNCHANGED
    where they lie (lib/src/CMakeLists.txt:26-32 minus bluetooth_packet.c) and linked into
    lib/full/libbtbb.so.1.  That is the file to hand to callers such as ubertooth-rx that also use
    btbb_piconet_* / btbb_process_packet / btbb_pcap* / lell_* (INTEGRATION.md).  No reference source is copied
    into this tree; without a checkout of upstream nothing is built and None is returned."""
    src = os.path.join(reference_root, "lib", "src")
    if not os.path.exists(os.path.join(src, "bluetooth_piconet.c")):
        return None
    build()
    os.makedirs(os.path.dirname(FULL_LIB), exist_ok=True)
    objs = []
    for u in REF_UNITS:
        o = os.path.join(OBJDIR, "upstream_" + u + ".o")
        cmd = ["gcc", "-O2", "-fPIC", "-w", "-std=gnu90", "-I" + src, "-c", os.path.join(src, u), "-o", o]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
        objs.append(o)
    ours = [os.path.join(OBJDIR, s + ".o") for s in _sources()]
    cmd = [NVCC] + LDFLAGS + ours + objs + ["-o", FULL_LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("linking lib/full/libbtbb.so.1 failed")
    return FULL_LIB


if __name__ == "__main__":
    if "--compose-reference" in sys.argv:
        i = sys.argv.index("--compose-reference")
        root = sys.argv[i + 1] if i + 1 < len(sys.argv) and not sys.argv[i + 1].startswith("-") else "/root/reference"
        print(compose(root, verbose="-v" in sys.argv))
        sys.exit(0)
    if "--install" in sys.argv:
        print(install(sys.argv[sys.argv.index("--install") + 1], full="--full" in sys.argv))
        sys.exit(0)
    build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_verbose="--ptxas-verbose" in sys.argv)
    print(LIB)
