"""The C-ABI shared object loads on a machine without a GPU and exports every symbol that
include/btbb_b200.h and include/btbb.h declare.  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

import util
from util import B


def declared(header):
    txt = open(os.path.join(util.ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(btbb_[a-z0-9_]+)\s*\(", txt)))


def test_exports_every_declared_symbol(product_lib):
    names = declared("btbb_b200.h") + declared("btbb.h") + B.CLASSIC_SYMBOLS
    assert len(names) > 60
    missing = [n for n in names if not hasattr(product_lib, n)]
    assert missing == []


def test_soname_and_record_layouts(product_lib):
    import subprocess
    out = subprocess.run(["readelf", "-d", B.LIB_PATH], capture_output=True, text=True).stdout
    assert "libbtbb.so.1" in out
    assert C.sizeof(B.Hit) == 16 and C.sizeof(B.Decoded) == 372 and C.sizeof(B.PktIn) == 24


def test_argument_validation_without_gpu(product_lib):
    h = C.c_void_p()
    assert product_lib.btbb_b200_create(0, 6, C.byref(h)) == -1      # same range as btbb_init (:282)
    assert product_lib.btbb_b200_create(0, -1, C.byref(h)) == -1
    assert b"max_ac_errors" in product_lib.btbb_b200_last_error()
    assert product_lib.btbb_gen_syncword(0x9e8b33) == 0x4e7a2cce331a3ae2   # pure host helper


def test_fails_loudly_without_cuda(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = product_lib.btbb_b200_create(0, 2, C.byref(h))
    assert rc == -2 and not h.value            # no CPU fallback: creation fails
    assert b"no CUDA device" in product_lib.btbb_b200_last_error()


def test_host_pack_bit_order_and_limit(product_lib):
    """host half of the packed transfer format: symbol i -> bit i & 31 of word i >> 5, nothing
    read at or past the limit (no GPU involved)."""
    import ctypes as C
    import numpy as np
    f = product_lib.bt_pack_range
    f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    f.restype = None
    rng = np.random.default_rng(3)
    for n in (1, 31, 32, 33, 127, 128, 129, 1000, 4096 + 17):
        s = rng.integers(0, 2, n, dtype=np.uint8)
        guard = np.concatenate([s, np.full(64, 0xFF, dtype=np.uint8)])     # poison past the limit
        for first in (0, 32, 96):
            if first >= n:
                continue
            nwords = (n - first + 31) // 32
            out = np.zeros(nwords, dtype=np.uint32)
            f(guard.ctypes.data, first, nwords, n, out.ctypes.data)
            pad = (-(n - first)) % 32
            want = np.packbits(np.concatenate([s[first:], np.zeros(pad, dtype=np.uint8)]), bitorder="little").view("<u4")
            assert (out == want).all(), (n, first)
    # the multi-stream form (several address ranges of a block advanced in lock step): same words
    g = product_lib.bt_pack_range_streams
    g.argtypes = f.argtypes + [C.c_int]
    g.restype = None
    for n in (4096 * 32, 4096 * 32 + 31, 16384 * 32, 16384 * 32 + 777, 5000 * 32 + 5):
        s = rng.integers(0, 2, n, dtype=np.uint8)
        guard = np.concatenate([s, np.full(64, 0xFF, dtype=np.uint8)])
        nwords = (n + 31) // 32
        want = np.packbits(np.concatenate([s, np.zeros((-n) % 32, dtype=np.uint8)]), bitorder="little").view("<u4")
        for streams in (1, 2, 3, 4, 8, 16):
            out = np.zeros(nwords, dtype=np.uint32)
            g(guard.ctypes.data, 0, nwords, n, out.ctypes.data, streams)
            assert (out == want).all(), (n, streams)


C_CALLER = r"""
/* what an existing libbtbb caller does: include the installed headers, link with -lbtbb */
#include <stdio.h>
#include <stdlib.h>
#include "btbb.h"
#include "btbb_b200.h"

int main(void)
{
	btbb_b200_ctx *ctx = NULL;
	btbb_b200_sieve pn;
	btbb_b200_hit hit;
	int rc = btbb_b200_create(0, 2, &ctx);
	/* sync word of LAP 0x9e8b33 (SURVEY.md 8a4) through the classic surface: a pure host helper */
	unsigned long long sw = (unsigned long long)btbb_gen_syncword(0x9e8b33);
	printf("%d %d %d %llx %s\n", rc, (int)sizeof(hit), (int)sizeof(pn), sw, btbb_get_release());
	if (ctx) btbb_b200_destroy(ctx);
	return 0;
}
"""


def test_c99_caller_compiles_links_and_runs(product_lib, tmp_path):
    """The public headers are plain C and a C program links against libbtbb.so.1 by its SONAME --
    the reference's own integration path (lib/libbtbb.pc.in: -lbtbb)."""
    import subprocess
    import torch
    src = tmp_path / "caller.c"
    src.write_text(C_CALLER)
    exe = tmp_path / "caller"
    libdir = os.path.dirname(B.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(util.ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-l:libbtbb.so.1", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    rc, hit_size, sieve_size, sw = int(out[0]), int(out[1]), int(out[2]), out[3]
    assert hit_size == 16 and sieve_size == 160 and sw == "4e7a2cce331a3ae2"
    assert rc == (0 if torch.cuda.is_available() else -2)


def test_group_by_lap_is_a_stable_grouping(product_lib):
    import numpy as np
    rng = np.random.default_rng(8)
    for n in (0, 1, 2, 17, 5000):
        hits = np.zeros(n, dtype=B.HIT_DTYPE)
        hits["offset"] = np.sort(rng.integers(0, 10**9, n))
        hits["lap"] = rng.choice(np.array([0x9E8B33, 0x123456, 0xFFFFFF, 0, 77], dtype=np.uint32), n)
        order, gs, laps = B.group_by_lap(hits)
        want = np.argsort(hits["lap"], kind="stable")
        assert np.array_equal(order, want)
        assert list(laps) == sorted(set(hits["lap"].tolist())) and gs[0] == 0 and gs[-1] == n if n else len(gs) == 1
        for g, lap in enumerate(laps):
            grp = hits[order[gs[g]:gs[g + 1]]]
            assert (grp["lap"] == lap).all() and (np.diff(grp["offset"]) >= 0).all()


def test_pkg_config_and_static_library(product_lib, tmp_path):
    """The rest of upstream's build contract (lib/libbtbb.pc.in:6-10, lib/src/CMakeLists.txt:43-63): a
    caller that asks pkg-config for its flags links against the unversioned libbtbb.so, and the same
    caller links statically against libbtbb.a with the private libraries the .pc file names."""
    import shutil
    import subprocess
    import torch
    from libbtbb_b200 import build
    if not shutil.which("pkg-config"):
        pytest.skip("no pkg-config here")
    build._dev_files()
    env = dict(os.environ, PKG_CONFIG_PATH=os.path.join(os.path.dirname(B.LIB_PATH), "pkgconfig"))
    pc = lambda *a: subprocess.run(["pkg-config", *a, "libbtbb"], env=env, capture_output=True, text=True, check=True).stdout.split()
    assert "-lbtbb" in pc("--libs")
    src = tmp_path / "caller.c"
    src.write_text(C_CALLER)
    libdir = os.path.dirname(B.LIB_PATH)
    want_rc = 0 if torch.cuda.is_available() else -2
    dyn = tmp_path / "caller_dyn"
    subprocess.run(["gcc", "-std=c99", str(src), *pc("--cflags", "--libs"), f"-Wl,-rpath,{libdir}", "-o", str(dyn)], check=True)
    out = subprocess.run([str(dyn)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == want_rc and out[3] == "4e7a2cce331a3ae2"
    assert os.path.exists(build.STATIC)
    private = [f for f in pc("--static", "--libs") if f != "-lbtbb"]
    rpaths = [f"-Wl,-rpath,{f[2:]}" for f in private if f.startswith("-L")]
    sta = tmp_path / "caller_static"
    subprocess.run(["gcc", "-std=c99", str(src), *pc("--cflags"), build.STATIC, *private, *rpaths, "-o", str(sta)], check=True)
    out = subprocess.run([str(sta)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == want_rc and out[3] == "4e7a2cce331a3ae2"
    ldd = subprocess.run(["ldd", str(sta)], capture_output=True, text=True).stdout
    assert "libbtbb" not in ldd


def test_install_layout_is_found_by_upstreams_cmake_module(product_lib, tmp_path):
    """`python -m libbtbb_b200.build --install <prefix>` lays the library out the way upstream's install step
    does; a CMake project that uses upstream's own FindBTBB.cmake (cmake/modules/FindBTBB.cmake:24-36, read
    from the reference checkout where there is one) finds it through LIBBTBB_DIR, builds and runs."""
    import shutil
    import subprocess
    import torch
    from libbtbb_b200 import build
    module_dir = "/root/reference/cmake/modules"
    if not shutil.which("cmake") or not os.path.exists(os.path.join(module_dir, "FindBTBB.cmake")):
        pytest.skip("needs cmake and an upstream checkout")
    prefix = tmp_path / "prefix"
    build.install(str(prefix))
    assert sorted(os.listdir(prefix / "include")) == ["btbb.h", "btbb_b200.h"]
    assert os.readlink(prefix / "lib" / "libbtbb.so") == "libbtbb.so.1" and (prefix / "lib" / "pkgconfig" / "libbtbb.pc").exists()
    proj = tmp_path / "proj"
    proj.mkdir()
    (proj / "caller.c").write_text(C_CALLER)
    (proj / "CMakeLists.txt").write_text(
        "cmake_minimum_required(VERSION 3.5)\nproject(caller C)\n"
        f"list(APPEND CMAKE_MODULE_PATH {module_dir})\n"
        "find_package(BTBB REQUIRED)\n"
        "include_directories(${LIBBTBB_INCLUDE_DIR})\n"
        "add_executable(caller caller.c)\n"
        "target_link_libraries(caller ${LIBBTBB_LIBRARIES})\n")
    env = dict(os.environ, LIBBTBB_DIR=str(prefix), PKG_CONFIG_PATH="")
    bdir = tmp_path / "b"
    subprocess.run(["cmake", "-S", str(proj), "-B", str(bdir)], env=env, check=True, capture_output=True)
    subprocess.run(["cmake", "--build", str(bdir)], env=env, check=True, capture_output=True)
    out = subprocess.run([str(bdir / "caller")], capture_output=True, text=True, check=True,
                         env=dict(os.environ, LD_LIBRARY_PATH=str(prefix / "lib"))).stdout.split()
    assert int(out[0]) == (0 if torch.cuda.is_available() else -2) and out[3] == "4e7a2cce331a3ae2"
