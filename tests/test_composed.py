"""The COMPLETE drop-in: lib/full/libbtbb.so.1 = this library's packet layer + upstream's own
bluetooth_piconet.c / pcap.c / pcapng*.c / bluetooth_le_packet.c compiled unchanged
(`python -m libbtbb_b200.build --compose-reference`, INTEGRATION.md option 2).  It exports every
symbol the reference library exports, and upstream's piconet layer and capture writers behave on
top of this packet layer exactly as on top of their own (same survey results, byte-identical pcap
and pcapng files).  Needs a checkout of upstream at build time; skipped where neither the composed
library nor the compiled reference (oracle/_ref) exists."""
import json
import os
import subprocess
import sys

import pytest

import util

FULL = os.path.join(util.ROOT, "libbtbb_b200", "lib", "full", "libbtbb.so.1")
DRIVER = os.path.join(util.ROOT, "tests", "composed_driver.py")
needs = pytest.mark.skipif(not (os.path.exists(FULL) and util.have_ref()), reason="composed library / compiled reference not built here")


def _exports(path):
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    return {l.split()[2] for l in out.splitlines() if len(l.split()) == 3 and l.split()[1] in "TBDR"}


@needs
def test_composed_library_exports_the_whole_reference_surface():
    ref = {s for s in _exports(util.REF_SO) if not s.startswith("ref_")}
    full = _exports(FULL)
    assert ref - full == set(), sorted(ref - full)
    for s in ("btbb_process_packet", "btbb_piconet_new", "btbb_pcap_append_packet", "btbb_pcapng_append_packet",
              "lell_allocate_and_decode", "btbb_find_ac", "btbb_b200_find_ac_dev"):
        assert s in full


@needs
def test_upstream_piconet_and_capture_writers_on_this_packet_layer():
    res = []
    for lib in (util.REF_SO, FULL):
        r = subprocess.run([sys.executable, DRIVER, lib], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(json.loads(r.stdout.strip().splitlines()[-1]))
    ref, full = res
    assert full["survey"] == ref["survey"] and len(ref["survey"]) == 5
    assert sum(1 for s in ref["survey"] if s[2]) >= 4          # UAPs were actually discovered
    assert full["pcap"] == ref["pcap"] and full["pcapng"] == ref["pcapng"]
    assert ref["pcap"][0] > 24 + 120 * 30
